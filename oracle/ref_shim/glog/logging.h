// ORACLE — TEST INFRASTRUCTURE ONLY. Not part of the product path.
//
// Minimal stand-in for <glog/logging.h> (absent from this image) so that the reference's sources compile
// unchanged into oracle/_ref: LOG / VLOG / CHECK* with glog's stream syntax. Messages go to stderr;
// FATAL and failed CHECKs abort(), as glog does. Verbosity comes from --v (see gflags shim) or $GLOG_v.
#pragma once
#include <cmath>
#include <cstdlib>
#include <iostream>
#include <sstream>
#include <string>

namespace google {
enum LogSeverity { GLOG_INFO = 0, GLOG_WARNING = 1, GLOG_ERROR = 2, GLOG_FATAL = 3 };
int& refShimVerbosity();        // defined in ref_shim.cpp
int& refShimMinLogLevel();
inline void InitGoogleLogging(const char*) {}
class LogMessage {
 public:
  LogMessage(const char* file, int line, int sev, bool enabled = true) : sev_(sev), on_(enabled) {
    if (on_) ss_ << "IWEF"[sev] << " " << file << ":" << line << "] ";
  }
  ~LogMessage() {
    if (on_ && (sev_ >= refShimMinLogLevel() || sev_ == GLOG_FATAL)) {
      ss_ << "\n";
      std::cerr << ss_.str() << std::flush;
    }
    if (sev_ == GLOG_FATAL) abort();
  }
  std::ostream& stream() { return ss_; }
 private:
  std::ostringstream ss_;
  int sev_;
  bool on_;
};
struct LogMessageVoidify {
  void operator&(std::ostream&) {}
};
template <typename T>
T& CheckNotNull(const char* file, int line, const char* what, T& t) {
  if (t == nullptr) LogMessage(file, line, GLOG_FATAL).stream() << what;
  return t;
}
template <typename T>
T CheckNotNullRV(const char* file, int line, const char* what, T t) {
  if (t == nullptr) LogMessage(file, line, GLOG_FATAL).stream() << what;
  return t;
}
}  // namespace google

#define LOG(sev) google::LogMessage(__FILE__, __LINE__, google::GLOG_##sev).stream()
#define LOG_IF(sev, cond) !(cond) ? (void)0 : google::LogMessageVoidify() & LOG(sev)
#define VLOG_IS_ON(n) (google::refShimVerbosity() >= (n))
#define VLOG(n) !VLOG_IS_ON(n) ? (void)0 : google::LogMessageVoidify() & google::LogMessage(__FILE__, __LINE__, google::GLOG_INFO).stream()
#define CHECK(cond) (cond) ? (void)0 : google::LogMessageVoidify() & google::LogMessage(__FILE__, __LINE__, google::GLOG_FATAL).stream() << "Check failed: " #cond " "
#define REF_SHIM_CHECK_OP(a, b, op) CHECK((a) op (b))
#define CHECK_EQ(a, b) REF_SHIM_CHECK_OP(a, b, ==)
#define CHECK_NE(a, b) REF_SHIM_CHECK_OP(a, b, !=)
#define CHECK_LT(a, b) REF_SHIM_CHECK_OP(a, b, <)
#define CHECK_LE(a, b) REF_SHIM_CHECK_OP(a, b, <=)
#define CHECK_GT(a, b) REF_SHIM_CHECK_OP(a, b, >)
#define CHECK_GE(a, b) REF_SHIM_CHECK_OP(a, b, >=)
#define CHECK_NEAR(a, b, margin) CHECK(std::abs(double(a) - double(b)) <= double(margin))
#define CHECK_NOTNULL(p) google::CheckNotNullRV(__FILE__, __LINE__, "'" #p "' Must be non NULL", (p))
#define DCHECK(cond) CHECK(cond)
