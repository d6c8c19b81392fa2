// ORACLE — TEST INFRASTRUCTURE ONLY. Not part of the product path.
//
// C entry points over the reference's OWN footage container reader and raw-sample converter, compiled unchanged from /root/reference
// (camera_isp/BinaryFootageFile.cpp, camera_isp/RawConverter.cpp): what Unpacker.cpp:66-146 does before it hands a frame to the ISP.
#include <cstring>
#include <iostream>
#include <sstream>
#include <string>

#include "camera_isp/BinaryFootageFile.h"
#include "camera_isp/RawConverter.h"

namespace {
std::string gErr;
using namespace surround360;
}  // namespace

extern "C" {
const char* refunpack_last_error() { return gErr.c_str(); }

// RawConverter::convert8Frame / convert12Frame (RawConverter.cpp:14-59): packed footage samples -> w x h uint16
int refunpack_convert(const void* packed, int bits, int w, int h, unsigned short* out) {
  try {
    auto r = bits == 8 ? RawConverter::convert8Frame(packed, w, h) : RawConverter::convert12Frame(packed, w, h);
    std::memcpy(out, r->data(), (size_t)w * h * 2);
    return 0;
  } catch (const std::exception& e) { gErr = e.what(); return 1; }
}

// BinaryFootageFile: open() prints the header to stdout (BinaryFootageFile.cpp:121-133); it is swallowed here.
void* refunpack_open(const char* path, unsigned meta8[8], unsigned long long* frames, unsigned long long* frameSize) {
  try {
    auto* f = new BinaryFootageFile(path);
    std::streambuf* old = std::cout.rdbuf();
    std::ostringstream sink;
    std::cout.rdbuf(sink.rdbuf());
    try { f->open(); } catch (...) { std::cout.rdbuf(old); delete f; throw; }
    std::cout.rdbuf(old);
    std::memcpy(meta8, &f->getMetadata(), 32);
    *frames = f->getNumberOfFrames();
    *frameSize = f->getFrameSize();
    return f;
  } catch (const std::exception& e) { gErr = e.what(); return nullptr; }
}
// getFrame (BinaryFootageFile.cpp:143-154): copies the frame's bytes out of the mapping
int refunpack_frame(void* handle, unsigned frame, unsigned camera, void* out) {
  try {
    auto* f = static_cast<BinaryFootageFile*>(handle);
    std::memcpy(out, f->getFrame(frame, camera), f->getFrameSize());
    return 0;
  } catch (const std::exception& e) { gErr = e.what(); return 1; }
}
void refunpack_close(void* handle) { delete static_cast<BinaryFootageFile*>(handle); }
}
