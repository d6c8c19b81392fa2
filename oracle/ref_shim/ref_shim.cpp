// ORACLE — TEST INFRASTRUCTURE ONLY. Not part of the product path.
//
// Out-of-line half of the third-party stand-ins under oracle/ref_shim/ (OpenCV, gflags, glog, folly) that let the
// reference's own sources compile unchanged into oracle/_ref. The OpenCV image primitives forward to oracle/ocv.h
// (bit-pinned against python cv2 4.13, tests/test_oracle_cv.py); this file only adapts cv::Mat <-> orc::Img and
// checks that the call is one of the argument combinations the reference's render path uses.
#include <dirent.h>
#include <fstream>
#include <map>
#include <mutex>
#include <sstream>

#include "opencv2/core.hpp"
#include "opencv2/imgcodecs.hpp"
#include "opencv2/imgproc.hpp"
#include "folly/dynamic.h"
#include "gflags/gflags.h"
#include "glog/logging.h"

#include "../ocv.h"
#include "../../surround360_b200/csrc/cli/png_io.h"   // plain zlib PNG codec shared with the command-line tool (host I/O only)

// =====================================================================================================================
// OpenCV
// =====================================================================================================================
namespace cv {
namespace {

orc::Img8 toImg8(const Mat& m) {
  if (m.depth() != CV_8U) refShimUnsupported("expected an 8-bit image");
  orc::Img8 o(m.cols, m.rows, m.channels());
  const size_t rb = (size_t)m.cols * m.channels();
  for (int y = 0; y < m.rows; ++y) std::memcpy(o.p(y), m.ptr(y), rb);
  return o;
}
orc::ImgF toImgF(const Mat& m) {
  if (m.depth() != CV_32F) refShimUnsupported("expected a float image");
  orc::ImgF o(m.cols, m.rows, m.channels());
  const size_t rb = (size_t)m.cols * m.channels() * sizeof(float);
  for (int y = 0; y < m.rows; ++y) std::memcpy(o.p(y), m.ptr(y), rb);
  return o;
}
// OutputArray::create + copy: reuses dst's buffer when shape and type already match (as cv::Mat::create does)
void store(const orc::Img8& s, Mat& dst) {
  dst.create(s.h, s.w, CV_MAKETYPE(CV_8U, s.c));
  const size_t rb = (size_t)s.w * s.c;
  for (int y = 0; y < s.h; ++y) std::memcpy(dst.ptr(y), s.p(y), rb);
}
void store(const orc::ImgF& s, Mat& dst) {
  dst.create(s.h, s.w, CV_MAKETYPE(CV_32F, s.c));
  const size_t rb = (size_t)s.w * s.c * sizeof(float);
  for (int y = 0; y < s.h; ++y) std::memcpy(dst.ptr(y), s.p(y), rb);
}
int orcBorder(int b) {
  switch (b) {
    case BORDER_CONSTANT: return orc::B_CONSTANT;
    case BORDER_REPLICATE: return orc::B_REPLICATE;
    case BORDER_WRAP: return orc::B_WRAP;
    case BORDER_REFLECT_101: return orc::B_REFLECT101;
  }
  refShimUnsupported("border mode");
}

}  // namespace

void split(const Mat& src, std::vector<Mat>& mv) {
  const int cn = src.channels();
  const size_t es = src.elemSize1();
  mv.assign(cn, Mat());
  for (int c = 0; c < cn; ++c) {
    mv[c].create(src.rows, src.cols, CV_MAKETYPE(src.depth(), 1));
    for (int y = 0; y < src.rows; ++y) {
      const uchar* s = src.ptr(y) + c * es;
      uchar* d = mv[c].ptr(y);
      for (int x = 0; x < src.cols; ++x) std::memcpy(d + x * es, s + (size_t)x * cn * es, es);
    }
  }
}
void merge(const std::vector<Mat>& mv, OutputArray dstA) {
  if (mv.empty()) refShimUnsupported("merge of nothing");
  const int cn = (int)mv.size();
  const size_t es = mv[0].elemSize1();
  for (const Mat& m : mv)
    if (m.channels() != 1 || m.size() != mv[0].size() || m.depth() != mv[0].depth()) refShimUnsupported("merge of unlike planes");
  Mat out(mv[0].rows, mv[0].cols, CV_MAKETYPE(mv[0].depth(), cn));
  for (int c = 0; c < cn; ++c)
    for (int y = 0; y < out.rows; ++y) {
      const uchar* s = mv[c].ptr(y);
      uchar* d = out.ptr(y) + c * es;
      for (int x = 0; x < out.cols; ++x) std::memcpy(d + (size_t)x * cn * es, s + x * es, es);
    }
  dstA.ref() = out;
}
void flip(const Mat& src, OutputArray dstA, int code) {
  Mat out(src.rows, src.cols, src.type());
  const size_t es = src.elemSize();
  for (int y = 0; y < src.rows; ++y) {
    const int sy = (code == 0 || code < 0) ? src.rows - 1 - y : y;
    const uchar* s = src.ptr(sy);
    uchar* d = out.ptr(y);
    if (code > 0 || code < 0) {
      for (int x = 0; x < src.cols; ++x) std::memcpy(d + x * es, s + (size_t)(src.cols - 1 - x) * es, es);
    } else {
      std::memcpy(d, s, (size_t)src.cols * es);
    }
  }
  // in place: OpenCV writes through the destination's existing buffer (CvUtil.cpp:120-127 flips shared Mats in place)
  Mat& dst = dstA.ref();
  if (dst.data && dst.rows == out.rows && dst.cols == out.cols && dst.type() == out.type()) {
    for (int y = 0; y < out.rows; ++y) std::memcpy(dst.ptr(y), out.ptr(y), (size_t)out.cols * es);
  } else {
    dst = out;
  }
}
void hconcat(const Mat& a, const Mat& b, OutputArray dstA) {
  if (a.rows != b.rows || a.type() != b.type()) refShimUnsupported("hconcat of unlike images");
  Mat out(a.rows, a.cols + b.cols, a.type());
  const size_t es = a.elemSize();
  for (int y = 0; y < a.rows; ++y) {
    std::memcpy(out.ptr(y), a.ptr(y), (size_t)a.cols * es);
    std::memcpy(out.ptr(y) + (size_t)a.cols * es, b.ptr(y), (size_t)b.cols * es);
  }
  dstA.ref() = out;
}
void vconcat(const Mat& a, const Mat& b, OutputArray dstA) {
  if (a.cols != b.cols || a.type() != b.type()) refShimUnsupported("vconcat of unlike images");
  Mat out(a.rows + b.rows, a.cols, a.type());
  const size_t rb = (size_t)a.cols * a.elemSize();
  for (int y = 0; y < a.rows; ++y) std::memcpy(out.ptr(y), a.ptr(y), rb);
  for (int y = 0; y < b.rows; ++y) std::memcpy(out.ptr(a.rows + y), b.ptr(y), rb);
  dstA.ref() = out;
}
void copyMakeBorder(const Mat& src, OutputArray dstA, int top, int bottom, int left, int right, int borderType, const Scalar& value) {
  if (borderType != BORDER_CONSTANT) refShimUnsupported("copyMakeBorder border type");
  Mat out(src.rows + top + bottom, src.cols + left + right, src.type(), value);
  const size_t es = src.elemSize();
  for (int y = 0; y < src.rows; ++y) std::memcpy(out.ptr(y + top) + (size_t)left * es, src.ptr(y), (size_t)src.cols * es);
  dstA.ref() = out;
}

void resize(const Mat& src, OutputArray dstA, Size dsize, double fx, double fy, int interpolation) {
  if (dsize.width <= 0 || dsize.height <= 0 || fx != 0 || fy != 0) refShimUnsupported("resize by factors");
  Mat& dst = dstA.ref();
  if (src.empty()) refShimUnsupported("resize of an empty image");
  if (src.depth() == CV_32F && interpolation == INTER_LINEAR) {
    orc::ImgF o;
    orc::resizeLinearF(toImgF(src), o, dsize.width, dsize.height);
    store(o, dst);
  } else if (src.depth() == CV_32F && interpolation == INTER_CUBIC) {
    orc::ImgF o;
    orc::resizeCubicF(toImgF(src), o, dsize.width, dsize.height);
    store(o, dst);
  } else if (src.depth() == CV_8U && interpolation == INTER_CUBIC && (src.channels() == 3 || src.channels() == 4)) {
    orc::Img8 o;
    orc::resizeCubic8(toImg8(src), o, dsize.width, dsize.height);
    store(o, dst);
  } else {
    refShimUnsupported("resize type / interpolation");
  }
}

void remap(const Mat& src, OutputArray dstA, const Mat& map1, const Mat& map2, int interpolation, int borderMode, const Scalar& borderValue) {
  if (map1.type() != CV_32FC2 || !map2.empty()) refShimUnsupported("remap map format");
  for (int c = 0; c < 4; ++c) if (borderValue.val[c] != 0) refShimUnsupported("remap border value");
  Mat& dst = dstA.ref();
  if (dst.data && dst.data == src.data) refShimUnsupported("remap in place");
  const orc::ImgF map = toImgF(map1);
  if (interpolation == INTER_CUBIC && src.depth() == CV_8U) {
    orc::Img8 o;
    orc::remapCubic8(toImg8(src), map, o, orcBorder(borderMode));
    store(o, dst);
  } else if (interpolation == INTER_CUBIC && src.depth() == CV_32F && borderMode == BORDER_CONSTANT) {
    orc::ImgF o;
    orc::remapCubicF(toImgF(src), map, o);
    store(o, dst);
  } else if (interpolation == INTER_NEAREST && src.depth() == CV_8U && borderMode == BORDER_WRAP) {
    // remapNearest: the float map is rounded half-to-even to short coordinates, then wrapped (imgwarp.cpp)
    Mat out(map1.rows, map1.cols, src.type());
    const size_t es = src.elemSize();
    for (int y = 0; y < out.rows; ++y)
      for (int x = 0; x < out.cols; ++x) {
        const int sx = orc::borderInterp(orc::sat_short(orc::cvRound(map.p(y, x)[0])), src.cols, orc::B_WRAP);
        const int sy = orc::borderInterp(orc::sat_short(orc::cvRound(map.p(y, x)[1])), src.rows, orc::B_WRAP);
        std::memcpy(out.ptr(y) + x * es, src.ptr(sy) + sx * es, es);
      }
    dst = out;
  } else {
    refShimUnsupported("remap type / interpolation / border");
  }
}

void GaussianBlur(const Mat& src, OutputArray dstA, Size ksize, double sigmaX, double sigmaY, int borderType) {
  if (ksize.width != ksize.height || (sigmaY != 0 && sigmaY != sigmaX) || borderType != BORDER_DEFAULT) refShimUnsupported("GaussianBlur parameters");
  Mat& dst = dstA.ref();
  if (src.depth() == CV_32F) {
    orc::ImgF o;
    orc::gaussianBlurF(toImgF(src), o, ksize.width, sigmaX);
    store(o, dst);
  } else if (src.type() == CV_8UC1) {
    orc::Img8 o;
    orc::gaussianBlur8Fixed(toImg8(src), o, orc::gaussKernelFixed8(ksize.width, sigmaX));
    store(o, dst);
  } else {
    refShimUnsupported("GaussianBlur type");
  }
}

void Sobel(const Mat& src, OutputArray dstA, int ddepth, int dx, int dy, int ksize, double scale, double delta, int borderType) {
  if (src.type() != CV_32FC1 || ddepth != -1 || ksize != 1 || scale != 1 || delta != 0 || borderType != BORDER_REPLICATE || dx + dy != 1)
    refShimUnsupported("Sobel parameters");
  orc::ImgF o;
  orc::sobel1(toImgF(src), o, dx == 1);
  store(o, dstA.ref());
}

void medianBlur(const Mat& src, OutputArray dstA, int ksize) {
  if (src.depth() != CV_32F || ksize != 5) refShimUnsupported("medianBlur parameters");
  orc::ImgF o;
  orc::median5(toImgF(src), o);
  store(o, dstA.ref());
}

Mat getStructuringElement(int shape, Size ksize, Point anchor) {
  if (shape != MORPH_CROSS) refShimUnsupported("structuring element shape");
  if (anchor.x < 0) anchor = Point(ksize.width / 2, ksize.height / 2);
  Mat k = Mat::zeros(ksize.height, ksize.width, CV_8UC1);
  for (int y = 0; y < ksize.height; ++y)
    for (int x = 0; x < ksize.width; ++x) k.at<uchar>(y, x) = (x == anchor.x || y == anchor.y) ? 1 : 0;
  return k;
}
void erode(const Mat& src, OutputArray dstA, const Mat& kernel) {
  // only the centred odd cross of featherAlphaChannel (CvUtil.cpp:144-149)
  const int n = kernel.rows, r = n / 2;
  if (src.type() != CV_8UC1 || kernel.rows != kernel.cols || (n & 1) == 0) refShimUnsupported("erode parameters");
  for (int y = 0; y < n; ++y)
    for (int x = 0; x < n; ++x)
      if ((kernel.at<uchar>(y, x) != 0) != (x == r || y == r)) refShimUnsupported("erode kernel is not a centred cross");
  orc::Img8 o;
  orc::erodeCross8(toImg8(src), o, r);
  store(o, dstA.ref());
}

void cvtColor(const Mat& src, OutputArray dstA, int code, int) {
  if (src.depth() == CV_32F && code == COLOR_BGR2BGRA && src.channels() == 3) {   // alpha = ColorChannel<float>::max() = 1.0f
    Mat out(src.rows, src.cols, CV_32FC4);
    for (int y = 0; y < src.rows; ++y) {
      const float* s = src.ptr<float>(y); float* d = out.ptr<float>(y);
      for (int x = 0; x < src.cols; ++x) { d[4 * x] = s[3 * x]; d[4 * x + 1] = s[3 * x + 1]; d[4 * x + 2] = s[3 * x + 2]; d[4 * x + 3] = 1.0f; }
    }
    dstA.ref() = out;
    return;
  }
  if (src.depth() != CV_8U) refShimUnsupported("cvtColor depth");
  Mat out;
  if (code == COLOR_BGR2BGRA && src.channels() == 3) {
    out.create(src.rows, src.cols, CV_8UC4);
    for (int y = 0; y < src.rows; ++y) {
      const uchar* s = src.ptr(y); uchar* d = out.ptr(y);
      for (int x = 0; x < src.cols; ++x) { d[4 * x] = s[3 * x]; d[4 * x + 1] = s[3 * x + 1]; d[4 * x + 2] = s[3 * x + 2]; d[4 * x + 3] = 255; }
    }
  } else if (code == COLOR_BGRA2BGR && src.channels() == 4) {
    out.create(src.rows, src.cols, CV_8UC3);
    for (int y = 0; y < src.rows; ++y) {
      const uchar* s = src.ptr(y); uchar* d = out.ptr(y);
      for (int x = 0; x < src.cols; ++x) { d[3 * x] = s[4 * x]; d[3 * x + 1] = s[4 * x + 1]; d[3 * x + 2] = s[4 * x + 2]; }
    }
  } else if ((code == COLOR_BGRA2GRAY && src.channels() == 4) || (code == COLOR_BGR2GRAY && src.channels() == 3)) {
    const int cn = src.channels();
    out.create(src.rows, src.cols, CV_8UC1);
    for (int y = 0; y < src.rows; ++y) {
      const uchar* s = src.ptr(y); uchar* d = out.ptr(y);
      for (int x = 0; x < src.cols; ++x) d[x] = orc::gray8(s[cn * x], s[cn * x + 1], s[cn * x + 2]);
    }
  } else {
    refShimUnsupported("cvtColor code");
  }
  dstA.ref() = out;
}

namespace refshim {
namespace {
std::mutex gMemMutex;
std::map<std::string, Mat> gMemImages;
std::vector<std::string> gMemSinks;
}  // namespace
static std::string normPath(const std::string& p) {   // "a//b" == "a/b", as for the file system
  std::string o;
  for (char ch : p) if (!(ch == '/' && !o.empty() && o.back() == '/')) o.push_back(ch);
  return o;
}
void memPut(const std::string& path, const Mat& m) { std::lock_guard<std::mutex> l(gMemMutex); gMemImages[normPath(path)] = m; }
bool memGet(const std::string& path, Mat& m) {
  std::lock_guard<std::mutex> l(gMemMutex);
  auto it = gMemImages.find(normPath(path));
  if (it == gMemImages.end()) return false;
  m = it->second;
  return true;
}
void memSinkPrefix(const std::string& prefix) { std::lock_guard<std::mutex> l(gMemMutex); gMemSinks.push_back(prefix); }
void memClear() { std::lock_guard<std::mutex> l(gMemMutex); gMemImages.clear(); gMemSinks.clear(); }
}  // namespace refshim

Mat imread(const String& filename, int flags) {
  {
    Mat m;
    if (refshim::memGet(filename, m)) {
      if (flags >= 0 && m.channels() == 4) { Mat bgr; cvtColor(m, bgr, COLOR_BGRA2BGR); return bgr; }   // IMREAD_COLOR: always 3 channels
      return m.clone();
    }
  }
  try {
    const bool keepAlpha = flags < 0;
    pngio::Image im = pngio::readColor(filename, keepAlpha);
    Mat m(im.h, im.w, CV_MAKETYPE(CV_8U, im.channels));
    std::memcpy(m.data, im.data.data(), im.data.size());
    return m;
  } catch (const std::exception&) {
    return Mat();   // cv::imread returns an empty Mat on failure
  }
}
bool imwrite(const String& filename, const Mat& img, const std::vector<int>&) {
  {
    bool sink = false;
    {
      std::lock_guard<std::mutex> l(refshim::gMemMutex);
      for (const std::string& p : refshim::gMemSinks) if (filename.compare(0, p.size(), p) == 0) sink = true;
    }
    if (sink) { refshim::memPut(filename, img.clone()); return true; }
  }
  if (img.depth() != CV_8U || (img.channels() != 3 && img.channels() != 4)) refShimUnsupported("imwrite type");
  try {
    const orc::Img8 t = toImg8(img);
    pngio::write(filename, t.d.data(), t.w, t.h, t.c);
    return true;
  } catch (const std::exception&) {
    return false;
  }
}

}  // namespace cv

// =====================================================================================================================
// glog / gflags
// =====================================================================================================================
namespace google {
int& refShimVerbosity() {
  static int v = [] { const char* e = std::getenv("GLOG_v"); return e ? std::atoi(e) : 0; }();
  return v;
}
int& refShimMinLogLevel() {
  static int v = [] { const char* e = std::getenv("GLOG_minloglevel"); return e ? std::atoi(e) : 0; }();
  return v;
}
}  // namespace google

namespace fLB {
bool FLAGS_help = false;
bool FLAGS_helpshort = false;
}

namespace gflags {
namespace {
struct FlagInfo { std::string type, help; void* storage; };
std::map<std::string, FlagInfo>& registry() {
  static std::map<std::string, FlagInfo> r;
  return r;
}
bool parseBool(const std::string& v, bool* out) {
  if (v == "true" || v == "1" || v == "t" || v == "yes" || v == "y") { *out = true; return true; }
  if (v == "false" || v == "0" || v == "f" || v == "no" || v == "n") { *out = false; return true; }
  return false;
}
[[noreturn]] void die(const std::string& msg) {
  std::fprintf(stderr, "ERROR: %s\n", msg.c_str());
  std::exit(1);
}
void assign(const std::string& name, const FlagInfo& f, const std::string& value) {
  try {
    if (f.type == "string") *(std::string*)f.storage = value;
    else if (f.type == "bool") { if (!parseBool(value, (bool*)f.storage)) die("illegal value '" + value + "' specified for bool flag '" + name + "'"); }
    else if (f.type == "int32") *(int32_t*)f.storage = (int32_t)std::stol(value);
    else if (f.type == "int64") *(int64_t*)f.storage = (int64_t)std::stoll(value);
    else if (f.type == "double") *(double*)f.storage = std::stod(value);
  } catch (const std::exception&) {
    die("illegal value '" + value + "' specified for " + f.type + " flag '" + name + "'");
  }
}
}  // namespace

FlagRegisterer::FlagRegisterer(const char* name, const char* type, const char* help, void* storage) {
  registry()[name] = FlagInfo{type, help, storage};
}

void ParseCommandLineNonHelpFlags(int* argc, char*** argv, bool removeFlags) {
  // glog's own flags, accepted and (where meaningful) honoured
  static int32_t glogV = 0, logbuflevel = 0, stderrthreshold = 2, minloglevel = 0;
  static std::string logDir;
  static bool logtostderr = false, alsologtostderr = false;
  static bool once = [] {
    registry()["v"] = FlagInfo{"int32", "verbosity", &glogV};
    registry()["logbuflevel"] = FlagInfo{"int32", "", &logbuflevel};
    registry()["stderrthreshold"] = FlagInfo{"int32", "", &stderrthreshold};
    registry()["minloglevel"] = FlagInfo{"int32", "", &minloglevel};
    registry()["log_dir"] = FlagInfo{"string", "", &logDir};
    registry()["logtostderr"] = FlagInfo{"bool", "", &logtostderr};
    registry()["alsologtostderr"] = FlagInfo{"bool", "", &alsologtostderr};
    registry()["help"] = FlagInfo{"bool", "", &fLB::FLAGS_help};
    registry()["helpshort"] = FlagInfo{"bool", "", &fLB::FLAGS_helpshort};
    return true;
  }();
  (void)once;
  std::vector<char*> rest;
  rest.push_back((*argv)[0]);
  for (int i = 1; i < *argc; ++i) {
    std::string a = (*argv)[i];
    if (a == "--") { for (int j = i + 1; j < *argc; ++j) rest.push_back((*argv)[j]); break; }
    if (a.size() < 2 || a[0] != '-') { rest.push_back((*argv)[i]); continue; }
    a = a.substr(a[1] == '-' ? 2 : 1);
    std::string name = a, value;
    bool hasValue = false;
    const size_t eq = a.find('=');
    if (eq != std::string::npos) { name = a.substr(0, eq); value = a.substr(eq + 1); hasValue = true; }
    auto it = registry().find(name);
    if (it == registry().end() && !hasValue && name.compare(0, 2, "no") == 0) {
      auto neg = registry().find(name.substr(2));
      if (neg != registry().end() && neg->second.type == "bool") { *(bool*)neg->second.storage = false; continue; }
    }
    if (it == registry().end()) die("unknown command line flag '" + name + "'");
    if (!hasValue) {
      if (it->second.type == "bool") { *(bool*)it->second.storage = true; continue; }
      if (i + 1 >= *argc) die("flag '--" + name + "' is missing its argument");
      value = (*argv)[++i];
    }
    assign(name, it->second, value);
  }
  google::refShimVerbosity() = glogV ? glogV : google::refShimVerbosity();
  google::refShimMinLogLevel() = minloglevel;
  if (removeFlags) {
    for (size_t i = 0; i < rest.size(); ++i) (*argv)[i] = rest[i];
    *argc = (int)rest.size();
  }
}
void ParseCommandLineFlags(int* argc, char*** argv, bool removeFlags) {
  ParseCommandLineNonHelpFlags(argc, argv, removeFlags);
  HandleCommandLineHelpFlags();
}
void HandleCommandLineHelpFlags() {
  if (!fLB::FLAGS_help && !fLB::FLAGS_helpshort) return;
  for (const auto& kv : registry())
    if (!kv.second.help.empty()) std::printf("    -%s (%s) type: %s\n", kv.first.c_str(), kv.second.help.c_str(), kv.second.type.c_str());
  std::exit(1);
}
}  // namespace gflags

// =====================================================================================================================
// folly::dynamic / json
// =====================================================================================================================
namespace folly {
constexpr dynamic::ObjectFn dynamic::object;

namespace {
struct Parser {
  const std::string& s;
  size_t i = 0;
  explicit Parser(const std::string& s_) : s(s_) {}
  [[noreturn]] void fail(const char* what) { throw std::runtime_error(std::string("json parse error: ") + what + " near offset " + std::to_string(i)); }
  void ws() { while (i < s.size() && (s[i] == ' ' || s[i] == '\n' || s[i] == '\t' || s[i] == '\r')) ++i; }
  std::string str() {
    std::string o;
    ++i;
    while (i < s.size() && s[i] != '"') {
      if (s[i] == '\\' && i + 1 < s.size()) {
        ++i;
        switch (s[i]) {
          case 'n': o.push_back('\n'); break;
          case 't': o.push_back('\t'); break;
          case 'r': o.push_back('\r'); break;
          case 'b': o.push_back('\b'); break;
          case 'f': o.push_back('\f'); break;
          default: o.push_back(s[i]);
        }
        ++i;
      } else {
        o.push_back(s[i++]);
      }
    }
    if (i >= s.size()) fail("unterminated string");
    ++i;
    return o;
  }
};
}  // namespace

static dynamic parseValue(Parser& p);
dynamic parseJson(const std::string& text) {
  Parser p(text);
  dynamic d = parseValue(p);
  p.ws();
  if (p.i != text.size()) p.fail("trailing characters");
  return d;
}
static dynamic parseValue(Parser& p) {
  p.ws();
  if (p.i >= p.s.size()) p.fail("unexpected end");
  const char ch = p.s[p.i];
  if (ch == '{') {
    dynamic::ObjectMaker m;
    ++p.i;
    p.ws();
    if (p.i < p.s.size() && p.s[p.i] == '}') { ++p.i; return dynamic(std::move(m)); }
    for (;;) {
      p.ws();
      if (p.i >= p.s.size() || p.s[p.i] != '"') p.fail("expected a key");
      const std::string k = p.str();
      p.ws();
      if (p.i >= p.s.size() || p.s[p.i++] != ':') p.fail("expected ':'");
      m.o[k] = std::make_shared<dynamic>(parseValue(p));
      p.ws();
      if (p.i < p.s.size() && p.s[p.i] == ',') { ++p.i; continue; }
      if (p.i < p.s.size() && p.s[p.i] == '}') { ++p.i; break; }
      p.fail("expected ',' or '}'");
    }
    return dynamic(std::move(m));
  }
  if (ch == '[') {
    dynamic a = dynamic::array();
    ++p.i;
    p.ws();
    if (p.i < p.s.size() && p.s[p.i] == ']') { ++p.i; return a; }
    for (;;) {
      a.push_back(parseValue(p));
      p.ws();
      if (p.i < p.s.size() && p.s[p.i] == ',') { ++p.i; continue; }
      if (p.i < p.s.size() && p.s[p.i] == ']') { ++p.i; break; }
      p.fail("expected ',' or ']'");
    }
    return a;
  }
  if (ch == '"') return dynamic(p.str());
  if (!p.s.compare(p.i, 4, "true")) { p.i += 4; return dynamic(true); }
  if (!p.s.compare(p.i, 5, "false")) { p.i += 5; return dynamic(false); }
  if (!p.s.compare(p.i, 4, "null")) { p.i += 4; return dynamic(nullptr); }
  // number: an integer literal becomes INT64, anything else DOUBLE (strtod: correctly rounded, like folly's parser)
  const char* b = p.s.c_str() + p.i;
  char* e = nullptr;
  const double dv = std::strtod(b, &e);
  if (e == b) p.fail("bad token");
  bool isInt = true;
  for (const char* q = b; q < e; ++q) if (!(std::isdigit((unsigned char)*q) || *q == '-' || *q == '+')) isInt = false;
  p.i += (size_t)(e - b);
  if (isInt && std::fabs(dv) < 9.0e18) return dynamic((long long)std::strtoll(b, nullptr, 10));
  return dynamic(dv);
}

std::string toJsonImpl(const dynamic& d, int indent, bool pretty) {
  std::ostringstream os;
  const std::string pad(pretty ? (size_t)indent * 2 : 0, ' '), pad2(pretty ? (size_t)(indent + 1) * 2 : 0, ' ');
  const char* nl = pretty ? "\n" : "";
  switch (d.type_) {
    case dynamic::NULLT: os << "null"; break;
    case dynamic::BOOL: os << (d.i_ ? "true" : "false"); break;
    case dynamic::INT64: os << d.i_; break;
    case dynamic::DOUBLE: { char buf[40]; std::snprintf(buf, sizeof buf, "%.17g", d.d_); os << buf; break; }
    case dynamic::STRING: os << '"' << d.s_ << '"'; break;
    case dynamic::ARRAY: {
      os << "[" << nl;
      for (size_t i = 0; i < d.a_.size(); ++i) os << pad2 << toJsonImpl(d.a_[i], indent + 1, pretty) << (i + 1 < d.a_.size() ? "," : "") << nl;
      os << pad << "]";
      break;
    }
    case dynamic::OBJECT: {
      os << "{" << nl;
      size_t i = 0;
      for (const auto& kv : d.o_) os << pad2 << '"' << kv.first << "\"" << (pretty ? " : " : ":") << toJsonImpl(*kv.second, indent + 1, pretty) << (++i < d.o_.size() ? "," : "") << nl;
      os << pad << "}";
      break;
    }
  }
  return os.str();
}
std::string toJson(const dynamic& d) { return toJsonImpl(d, 0, false); }
std::string toPrettyJson(const dynamic& d) { return toJsonImpl(d, 0, true); }
std::string dynamic::asString() const {
  if (type_ == STRING) return s_;
  return toJson(*this);
}
std::ostream& operator<<(std::ostream& os, const dynamic& d) { return os << toJson(d); }

bool readFile(const char* path, std::string& out) {
  std::ifstream f(path, std::ios::binary);
  if (!f) return false;
  std::ostringstream ss;
  ss << f.rdbuf();
  out = ss.str();
  return true;
}
bool writeFile(const std::string& data, const char* path) {
  std::ofstream f(path, std::ios::binary);
  if (!f) return false;
  f << data;
  return (bool)f;
}
}  // namespace folly
