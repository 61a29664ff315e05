// Stand-in for <glog/logging.h>, TEST INFRASTRUCTURE (oracle/_ref build only): CHECKs abort with a
// message, LOG/VLOG swallow their stream (the reference's per-step VLOG(5) lines are off by default).
#ifndef NBV_REF_SHIM_GLOG_
#define NBV_REF_SHIM_GLOG_
#include <cmath>
#include <cstdlib>
#include <iostream>
#include <sstream>

namespace nbv_ref_shim {
struct NullStream {
  template <class T>
  NullStream& operator<<(const T&) { return *this; }
  NullStream& operator<<(std::ostream& (*)(std::ostream&)) { return *this; }
};
struct FatalStream {
  std::ostringstream ss;
  FatalStream(const char* file, int line, const char* what) { ss << file << ":" << line << " CHECK failed: " << what << " "; }
  template <class T>
  FatalStream& operator<<(const T& v) { ss << v; return *this; }
  FatalStream& operator<<(std::ostream& (*)(std::ostream&)) { return *this; }
  [[noreturn]] ~FatalStream() { std::cerr << ss.str() << std::endl; std::abort(); }
};
struct Voidify { void operator&(const NullStream&) {} void operator&(const FatalStream&) {} };
template <class T>
T&& check_not_null(T&& p, const char* file, int line, const char* what) {
  if (p == nullptr) { std::cerr << file << ":" << line << " CHECK_NOTNULL(" << what << ")" << std::endl; std::abort(); }
  return static_cast<T&&>(p);
}
}  // namespace nbv_ref_shim

#define NBV_NULL_STREAM true ? (void)0 : ::nbv_ref_shim::Voidify() & ::nbv_ref_shim::NullStream()
#define CHECK(c) (c) ? (void)0 : ::nbv_ref_shim::Voidify() & ::nbv_ref_shim::FatalStream(__FILE__, __LINE__, #c)
#define CHECK_OP(a, op, b) CHECK((a)op(b))
#define CHECK_EQ(a, b) CHECK_OP(a, ==, b)
#define CHECK_NE(a, b) CHECK_OP(a, !=, b)
#define CHECK_LT(a, b) CHECK_OP(a, <, b)
#define CHECK_LE(a, b) CHECK_OP(a, <=, b)
#define CHECK_GT(a, b) CHECK_OP(a, >, b)
#define CHECK_GE(a, b) CHECK_OP(a, >=, b)
#define CHECK_NEAR(a, b, eps) CHECK(std::fabs((a) - (b)) <= (eps))
#define CHECK_NOTNULL(p) ::nbv_ref_shim::check_not_null((p), __FILE__, __LINE__, #p)
#define DCHECK(c) NBV_NULL_STREAM
#define DCHECK_EQ(a, b) NBV_NULL_STREAM
#define DCHECK_NE(a, b) NBV_NULL_STREAM
#define DCHECK_LT(a, b) NBV_NULL_STREAM
#define DCHECK_LE(a, b) NBV_NULL_STREAM
#define DCHECK_GT(a, b) NBV_NULL_STREAM
#define DCHECK_GE(a, b) NBV_NULL_STREAM
#define DCHECK_NOTNULL(p) (p)
#define LOG(sev) NBV_LOG_##sev
#define NBV_LOG_INFO NBV_NULL_STREAM
#define NBV_LOG_WARNING NBV_NULL_STREAM
#define NBV_LOG_ERROR NBV_NULL_STREAM
#define NBV_LOG_FATAL ::nbv_ref_shim::Voidify() & ::nbv_ref_shim::FatalStream(__FILE__, __LINE__, "LOG(FATAL)")
#define VLOG(n) NBV_NULL_STREAM
#define LOG_FIRST_N(sev, n) NBV_NULL_STREAM
#define LOG_EVERY_N(sev, n) NBV_NULL_STREAM
#define LOG_IF(sev, c) NBV_NULL_STREAM
#endif
