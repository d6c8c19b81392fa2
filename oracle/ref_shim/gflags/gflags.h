// ORACLE — TEST INFRASTRUCTURE ONLY. Not part of the product path.
//
// Minimal stand-in for <gflags/gflags.h> (absent from this image): DEFINE_* register a flag in a process-wide
// table, ParseCommandLineNonHelpFlags accepts "--name value", "--name=value", "-name ...", and "--name" /
// "--noname" for bools, as gflags does. glog's own flags (--v, --logbuflevel, --log_dir, --stderrthreshold,
// --minloglevel, --logtostderr, --alsologtostderr) are accepted so the reference's command lines parse.
#pragma once
#include <cstdint>
#include <string>

namespace gflags {
struct FlagRegisterer {
  FlagRegisterer(const char* name, const char* type, const char* help, void* storage);
};
void ParseCommandLineNonHelpFlags(int* argc, char*** argv, bool removeFlags);
void ParseCommandLineFlags(int* argc, char*** argv, bool removeFlags);
void HandleCommandLineHelpFlags();
}  // namespace gflags

namespace fLB {
extern bool FLAGS_help;
extern bool FLAGS_helpshort;
}

#define REF_SHIM_DEFINE_FLAG(type, tname, name, val, help) \
  type FLAGS_##name = val;                                 \
  static gflags::FlagRegisterer refShimFlagReg_##name(#name, tname, help, &FLAGS_##name)
#define DEFINE_string(name, val, help) REF_SHIM_DEFINE_FLAG(std::string, "string", name, val, help)
#define DEFINE_bool(name, val, help) REF_SHIM_DEFINE_FLAG(bool, "bool", name, val, help)
#define DEFINE_int32(name, val, help) REF_SHIM_DEFINE_FLAG(int32_t, "int32", name, val, help)
#define DEFINE_int64(name, val, help) REF_SHIM_DEFINE_FLAG(int64_t, "int64", name, val, help)
#define DEFINE_double(name, val, help) REF_SHIM_DEFINE_FLAG(double, "double", name, val, help)
#define DECLARE_string(name) extern std::string FLAGS_##name
#define DECLARE_bool(name) extern bool FLAGS_##name
#define DECLARE_int32(name) extern int32_t FLAGS_##name
#define DECLARE_double(name) extern double FLAGS_##name
