// Stand-in for <google/protobuf/message.h>, TEST INFRASTRUCTURE (oracle/_ref build only): the reference's
// layer_inl.h names these types in map save/load members that the hot path never instantiates.
#ifndef NBV_REF_SHIM_PROTOBUF_MESSAGE_
#define NBV_REF_SHIM_PROTOBUF_MESSAGE_
namespace google {
namespace protobuf {
class MessageLite {};
class Message : public MessageLite {};
}  // namespace protobuf
}  // namespace google
#endif
