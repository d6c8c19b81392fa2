// ORACLE — TEST INFRASTRUCTURE ONLY (see core.hpp). imread / imwrite for 8- and 16-bit PNG files (zlib).
#pragma once
#include "core.hpp"

namespace cv {
enum ImreadModes { IMREAD_UNCHANGED = -1, IMREAD_GRAYSCALE = 0, IMREAD_COLOR = 1, IMREAD_ANYDEPTH = 2, IMREAD_ANYCOLOR = 4 };
Mat imread(const String& filename, int flags = IMREAD_COLOR);
bool imwrite(const String& filename, const Mat& img, const std::vector<int>& params = std::vector<int>());

// Test hook (oracle/ref_capi.cpp): images registered under a path are served by imread instead of the file
// system, and imwrite keeps images written below a registered prefix in memory, so that a frame can be pushed
// through the reference's file-based renderStereoPanorama() without PNG encode/decode.
namespace refshim {
void memPut(const std::string& path, const Mat& m);
bool memGet(const std::string& path, Mat& m);
void memSinkPrefix(const std::string& prefix);
void memClear();
}  // namespace refshim
}  // namespace cv
#define CV_LOAD_IMAGE_UNCHANGED (-1)
#define CV_LOAD_IMAGE_GRAYSCALE 0
#define CV_LOAD_IMAGE_COLOR 1
