// Stand-in, TEST INFRASTRUCTURE (oracle/_ref build only).
#include <google/protobuf/message.h>
