// ORACLE — TEST INFRASTRUCTURE ONLY. Not part of the product path.
//
// Minimal stand-in for the slice of the OpenCV 3 C++ API that the reference's render path uses, so that the
// reference's own sources (optical_flow/PixFlow.h, optical_flow/NovelView.cpp, util/CvUtil.cpp, util/Filter.h,
// render/ImageWarper.cpp, render/PoleRemoval.cpp, test/TestRenderStereoPanorama.cpp, ...) compile UNCHANGED from
// /root/reference into oracle/_ref (recipe: oracle/ref_build.sh). OpenCV's C++ headers and libraries are absent
// from this image.
//
// Two kinds of content:
//  * header-level semantics of OpenCV's small types, restated from the published headers (core/types.hpp,
//    core/matx.hpp, core/mat.inl.hpp, core/cvstd.hpp): Point_/Vec arithmetic through saturate_cast, norm() in
//    double, Mat reference counting / ROI views / create() reuse rule, `Mat op= scalar` as convertTo(alpha),
//    and cvstd.hpp's `using std::exp` & co. inside namespace cv (it decides which libm overload the reference's
//    unqualified exp()/sqrt()/abs() calls pick);
//  * the image primitives (resize, remap, GaussianBlur, Sobel, medianBlur, erode, cvtColor, ...) forward to
//    oracle/ocv.h, which is pinned bit-exactly against python cv2 4.13 by tests/test_oracle_cv.py.
// Only the argument combinations the reference uses on this path are implemented; anything else aborts loudly.
#pragma once
#include <algorithm>
#include <cassert>
#include <ostream>
#include <algorithm>
#include <cfloat>
#include <climits>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <string>
#include <vector>

typedef unsigned char uchar;
typedef unsigned short ushort;
typedef int64_t int64;
typedef uint64_t uint64;

#define CV_8U 0
#define CV_8S 1
#define CV_16U 2
#define CV_16S 3
#define CV_32S 4
#define CV_32F 5
#define CV_64F 6
#define CV_CN_SHIFT 3
#define CV_MAT_DEPTH_MASK 7
#define CV_MAT_DEPTH(t) ((t) & 7)
#define CV_MAT_CN(t) ((((t) >> CV_CN_SHIFT) & 511) + 1)
#define CV_MAKETYPE(depth, cn) (CV_MAT_DEPTH(depth) + (((cn) - 1) << CV_CN_SHIFT))
#define CV_8UC1 CV_MAKETYPE(CV_8U, 1)
#define CV_8UC2 CV_MAKETYPE(CV_8U, 2)
#define CV_8UC3 CV_MAKETYPE(CV_8U, 3)
#define CV_8UC4 CV_MAKETYPE(CV_8U, 4)
#define CV_16UC1 CV_MAKETYPE(CV_16U, 1)
#define CV_16UC3 CV_MAKETYPE(CV_16U, 3)
#define CV_32FC1 CV_MAKETYPE(CV_32F, 1)
#define CV_32FC2 CV_MAKETYPE(CV_32F, 2)
#define CV_32FC3 CV_MAKETYPE(CV_32F, 3)
#define CV_32FC4 CV_MAKETYPE(CV_32F, 4)
#define CV_64FC1 CV_MAKETYPE(CV_64F, 1)

namespace cv {

[[noreturn]] inline void refShimUnsupported(const char* what) {
  std::fprintf(stderr, "oracle/ref_shim: unsupported OpenCV usage: %s\n", what);
  std::abort();
}

// ---- core/cvstd.hpp -------------------------------------------------------------------------------------------
static inline uchar abs(uchar a) { return a; }
static inline ushort abs(ushort a) { return a; }
static inline unsigned abs(unsigned a) { return a; }
static inline uint64 abs(uint64 a) { return a; }
using std::min;
using std::max;
using std::abs;
using std::swap;
using std::sqrt;
using std::exp;
using std::pow;
using std::log;
typedef std::string String;

// ---- core/fast_math.hpp, core/saturate.hpp ----------------------------------------------------------------------
static inline int cvRound(double v) {   // SSE2 cvtsd2si: round half to even, INT_MIN when out of range
  if (!(v > -2147483648.5 && v < 2147483647.5)) return INT32_MIN;
  return (int)std::nearbyint(v);
}
static inline int cvFloor(double v) { int i = (int)v; return i - (i > v); }
static inline int cvCeil(double v) { int i = (int)v; return i + (i < v); }
template <typename T> static inline T saturate_cast(uchar v) { return T(v); }
template <typename T> static inline T saturate_cast(ushort v) { return T(v); }
template <typename T> static inline T saturate_cast(short v) { return T(v); }
template <typename T> static inline T saturate_cast(int v) { return T(v); }
template <typename T> static inline T saturate_cast(unsigned v) { return T(v); }
template <typename T> static inline T saturate_cast(float v) { return T(v); }
template <typename T> static inline T saturate_cast(double v) { return T(v); }
template <> inline uchar saturate_cast<uchar>(int v) { return (uchar)((unsigned)v <= 255 ? v : v > 0 ? 255 : 0); }
template <> inline uchar saturate_cast<uchar>(short v) { return saturate_cast<uchar>((int)v); }
template <> inline uchar saturate_cast<uchar>(ushort v) { return (uchar)std::min<unsigned>(v, 255u); }
template <> inline uchar saturate_cast<uchar>(unsigned v) { return (uchar)std::min<unsigned>(v, 255u); }
template <> inline uchar saturate_cast<uchar>(float v) { return saturate_cast<uchar>(cvRound(v)); }
template <> inline uchar saturate_cast<uchar>(double v) { return saturate_cast<uchar>(cvRound(v)); }
template <> inline short saturate_cast<short>(int v) { return (short)((unsigned)(v + 32768) <= 65535u ? v : v > 0 ? 32767 : -32768); }
template <> inline short saturate_cast<short>(float v) { return saturate_cast<short>(cvRound(v)); }
template <> inline short saturate_cast<short>(double v) { return saturate_cast<short>(cvRound(v)); }
template <> inline ushort saturate_cast<ushort>(int v) { return (ushort)((unsigned)v <= 65535u ? v : v > 0 ? 65535 : 0); }
template <> inline ushort saturate_cast<ushort>(float v) { return saturate_cast<ushort>(cvRound(v)); }
template <> inline ushort saturate_cast<ushort>(double v) { return saturate_cast<ushort>(cvRound(v)); }
template <> inline int saturate_cast<int>(float v) { return cvRound(v); }
template <> inline int saturate_cast<int>(double v) { return cvRound(v); }

// ---- core/types.hpp: Point_, Point3_, Size_, Rect_, Scalar_ ----------------------------------------------------
template <typename T>
struct Point_ {
  typedef T value_type;
  T x, y;
  Point_() : x(0), y(0) {}
  Point_(T x_, T y_) : x(x_), y(y_) {}
  template <typename T2>
  operator Point_<T2>() const { return Point_<T2>(saturate_cast<T2>(x), saturate_cast<T2>(y)); }
  T dot(const Point_& pt) const { return saturate_cast<T>(x * pt.x + y * pt.y); }
  double ddot(const Point_& pt) const { return (double)x * pt.x + (double)y * pt.y; }
  double cross(const Point_& pt) const { return (double)x * pt.y - (double)y * pt.x; }
};
template <typename T> static inline Point_<T>& operator+=(Point_<T>& a, const Point_<T>& b) { a.x += b.x; a.y += b.y; return a; }
template <typename T> static inline Point_<T>& operator-=(Point_<T>& a, const Point_<T>& b) { a.x -= b.x; a.y -= b.y; return a; }
template <typename T> static inline Point_<T>& operator*=(Point_<T>& a, int b) { a.x = saturate_cast<T>(a.x * b); a.y = saturate_cast<T>(a.y * b); return a; }
template <typename T> static inline Point_<T>& operator*=(Point_<T>& a, float b) { a.x = saturate_cast<T>(a.x * b); a.y = saturate_cast<T>(a.y * b); return a; }
template <typename T> static inline Point_<T>& operator*=(Point_<T>& a, double b) { a.x = saturate_cast<T>(a.x * b); a.y = saturate_cast<T>(a.y * b); return a; }
template <typename T> static inline Point_<T>& operator/=(Point_<T>& a, int b) { a.x = saturate_cast<T>(a.x / b); a.y = saturate_cast<T>(a.y / b); return a; }
template <typename T> static inline Point_<T>& operator/=(Point_<T>& a, float b) { a.x = saturate_cast<T>(a.x / b); a.y = saturate_cast<T>(a.y / b); return a; }
template <typename T> static inline Point_<T>& operator/=(Point_<T>& a, double b) { a.x = saturate_cast<T>(a.x / b); a.y = saturate_cast<T>(a.y / b); return a; }
template <typename T> static inline double norm(const Point_<T>& pt) { return std::sqrt((double)pt.x * pt.x + (double)pt.y * pt.y); }
template <typename T> static inline bool operator==(const Point_<T>& a, const Point_<T>& b) { return a.x == b.x && a.y == b.y; }
template <typename T> static inline bool operator!=(const Point_<T>& a, const Point_<T>& b) { return a.x != b.x || a.y != b.y; }
template <typename T> static inline Point_<T> operator+(const Point_<T>& a, const Point_<T>& b) { return Point_<T>(saturate_cast<T>(a.x + b.x), saturate_cast<T>(a.y + b.y)); }
template <typename T> static inline Point_<T> operator-(const Point_<T>& a, const Point_<T>& b) { return Point_<T>(saturate_cast<T>(a.x - b.x), saturate_cast<T>(a.y - b.y)); }
template <typename T> static inline Point_<T> operator-(const Point_<T>& a) { return Point_<T>(saturate_cast<T>(-a.x), saturate_cast<T>(-a.y)); }
template <typename T> static inline Point_<T> operator*(const Point_<T>& a, int b) { return Point_<T>(saturate_cast<T>(a.x * b), saturate_cast<T>(a.y * b)); }
template <typename T> static inline Point_<T> operator*(int a, const Point_<T>& b) { return Point_<T>(saturate_cast<T>(b.x * a), saturate_cast<T>(b.y * a)); }
template <typename T> static inline Point_<T> operator*(const Point_<T>& a, float b) { return Point_<T>(saturate_cast<T>(a.x * b), saturate_cast<T>(a.y * b)); }
template <typename T> static inline Point_<T> operator*(float a, const Point_<T>& b) { return Point_<T>(saturate_cast<T>(b.x * a), saturate_cast<T>(b.y * a)); }
template <typename T> static inline Point_<T> operator*(const Point_<T>& a, double b) { return Point_<T>(saturate_cast<T>(a.x * b), saturate_cast<T>(a.y * b)); }
template <typename T> static inline Point_<T> operator*(double a, const Point_<T>& b) { return Point_<T>(saturate_cast<T>(b.x * a), saturate_cast<T>(b.y * a)); }
template <typename T> static inline Point_<T> operator/(const Point_<T>& a, int b) { Point_<T> t(a); t /= b; return t; }
template <typename T> static inline Point_<T> operator/(const Point_<T>& a, float b) { Point_<T> t(a); t /= b; return t; }
template <typename T> static inline Point_<T> operator/(const Point_<T>& a, double b) { Point_<T> t(a); t /= b; return t; }
typedef Point_<int> Point2i;
typedef Point_<float> Point2f;
typedef Point_<double> Point2d;
typedef Point2i Point;

template <typename T, int cn> struct Vec;
template <typename T>
struct Point3_ {
  T x, y, z;
  Point3_() : x(0), y(0), z(0) {}
  Point3_(T x_, T y_, T z_) : x(x_), y(y_), z(z_) {}
  Point3_(const Vec<T, 3>& v);          // types.hpp: Point3_(const Vec<_Tp, 3>& v)
  operator Vec<T, 3>() const;           //            operator Vec<_Tp, 3>() const
};
typedef Point3_<int> Point3i;
typedef Point3_<float> Point3f;
typedef Point3_<double> Point3d;

template <typename T>
struct Size_ {
  T width, height;
  Size_() : width(0), height(0) {}
  Size_(T w, T h) : width(w), height(h) {}
  T area() const { return width * height; }
};
template <typename T> static inline bool operator==(const Size_<T>& a, const Size_<T>& b) { return a.width == b.width && a.height == b.height; }
template <typename T> static inline bool operator!=(const Size_<T>& a, const Size_<T>& b) { return !(a == b); }
template <typename T> static inline Size_<T> operator/(const Size_<T>& a, T b) { return Size_<T>(a.width / b, a.height / b); }
template <typename T> static inline Size_<T> operator*(const Size_<T>& a, T b) { return Size_<T>(a.width * b, a.height * b); }
typedef Size_<int> Size2i;
typedef Size_<float> Size2f;
typedef Size2i Size;

template <typename T>
struct Rect_ {
  T x, y, width, height;
  Rect_() : x(0), y(0), width(0), height(0) {}
  Rect_(T x_, T y_, T w, T h) : x(x_), y(y_), width(w), height(h) {}
};
typedef Rect_<int> Rect2i;
typedef Rect2i Rect;

// ---- core/matx.hpp: Vec (as used: element access, conversion, scaling, sums) ---------------------------------------
template <typename T, int cn>
struct Vec {
  typedef T value_type;
  enum { channels = cn };
  T val[cn];
  Vec() { for (int i = 0; i < cn; ++i) val[i] = T(0); }
  Vec(T v0) { for (int i = 0; i < cn; ++i) val[i] = T(0); val[0] = v0; }
  Vec(T v0, T v1) { static_assert(cn >= 2, "channels"); for (int i = 2; i < cn; ++i) val[i] = T(0); val[0] = v0; val[1] = v1; }
  Vec(T v0, T v1, T v2) { static_assert(cn >= 3, "channels"); for (int i = 3; i < cn; ++i) val[i] = T(0); val[0] = v0; val[1] = v1; val[2] = v2; }
  Vec(T v0, T v1, T v2, T v3) { static_assert(cn >= 4, "channels"); for (int i = 4; i < cn; ++i) val[i] = T(0); val[0] = v0; val[1] = v1; val[2] = v2; val[3] = v3; }
  template <typename T2>
  operator Vec<T2, cn>() const {
    Vec<T2, cn> v;
    for (int i = 0; i < cn; ++i) v.val[i] = saturate_cast<T2>(val[i]);
    return v;
  }
  const T& operator[](int i) const { return val[i]; }
  T& operator[](int i) { return val[i]; }
  const T& operator()(int i) const { return val[i]; }
  T& operator()(int i) { return val[i]; }
};
template <typename T, int cn> static inline bool operator==(const Vec<T, cn>& a, const Vec<T, cn>& b) { for (int i = 0; i < cn; ++i) if (a.val[i] != b.val[i]) return false; return true; }
template <typename T, int cn> static inline bool operator!=(const Vec<T, cn>& a, const Vec<T, cn>& b) { return !(a == b); }
template <typename T, int cn> static inline Vec<T, cn> operator+(const Vec<T, cn>& a, const Vec<T, cn>& b) { Vec<T, cn> r; for (int i = 0; i < cn; ++i) r.val[i] = saturate_cast<T>(a.val[i] + b.val[i]); return r; }
template <typename T, int cn> static inline Vec<T, cn> operator-(const Vec<T, cn>& a, const Vec<T, cn>& b) { Vec<T, cn> r; for (int i = 0; i < cn; ++i) r.val[i] = saturate_cast<T>(a.val[i] - b.val[i]); return r; }
template <typename T, int cn> static inline Vec<T, cn>& operator+=(Vec<T, cn>& a, const Vec<T, cn>& b) { for (int i = 0; i < cn; ++i) a.val[i] = saturate_cast<T>(a.val[i] + b.val[i]); return a; }
template <typename T, int cn> static inline Vec<T, cn>& operator-=(Vec<T, cn>& a, const Vec<T, cn>& b) { for (int i = 0; i < cn; ++i) a.val[i] = saturate_cast<T>(a.val[i] - b.val[i]); return a; }
// Matx_ScaleOp: val[i] = saturate_cast<T>(a.val[i] * alpha); division multiplies by 1.f/alpha (float) or 1./alpha (int, double)
#define REF_SHIM_VEC_SCALE(S, RCP)                                                                                                                      \
  template <typename T, int cn> static inline Vec<T, cn> operator*(const Vec<T, cn>& a, S s) { Vec<T, cn> r; for (int i = 0; i < cn; ++i) r.val[i] = saturate_cast<T>(a.val[i] * s); return r; } \
  template <typename T, int cn> static inline Vec<T, cn> operator*(S s, const Vec<T, cn>& a) { Vec<T, cn> r; for (int i = 0; i < cn; ++i) r.val[i] = saturate_cast<T>(a.val[i] * s); return r; } \
  template <typename T, int cn> static inline Vec<T, cn> operator/(const Vec<T, cn>& a, S s) { Vec<T, cn> r; for (int i = 0; i < cn; ++i) r.val[i] = saturate_cast<T>(a.val[i] * (RCP / s)); return r; } \
  template <typename T, int cn> static inline Vec<T, cn>& operator*=(Vec<T, cn>& a, S s) { for (int i = 0; i < cn; ++i) a.val[i] = saturate_cast<T>(a.val[i] * s); return a; }
REF_SHIM_VEC_SCALE(int, 1.)
REF_SHIM_VEC_SCALE(float, 1.f)
REF_SHIM_VEC_SCALE(double, 1.)
#undef REF_SHIM_VEC_SCALE
template <typename T, int cn> static inline Vec<T, cn> operator-(const Vec<T, cn>& a) { Vec<T, cn> r; for (int i = 0; i < cn; ++i) r.val[i] = saturate_cast<T>(-a.val[i]); return r; }
template <typename T, int cn> static inline double norm(const Vec<T, cn>& v) {   // normL2Sqr<T, double> then sqrt
  double s = 0;
  for (int i = 0; i < cn; ++i) s += (double)v.val[i] * (double)v.val[i];
  return std::sqrt(s);
}
template <typename T> inline Point3_<T>::Point3_(const Vec<T, 3>& v) : x(v.val[0]), y(v.val[1]), z(v.val[2]) {}
template <typename T> inline Point3_<T>::operator Vec<T, 3>() const { return Vec<T, 3>(x, y, z); }
typedef Vec<short, 3> Vec3s;
typedef Vec<uchar, 2> Vec2b;
typedef Vec<uchar, 3> Vec3b;
typedef Vec<uchar, 4> Vec4b;
typedef Vec<ushort, 3> Vec3w;
typedef Vec<int, 2> Vec2i;
typedef Vec<float, 2> Vec2f;
typedef Vec<float, 3> Vec3f;
typedef Vec<float, 4> Vec4f;
typedef Vec<double, 3> Vec3d;

template <typename T>
struct Scalar_ : Vec<T, 4> {
  Scalar_() {}
  Scalar_(T v0) : Vec<T, 4>(v0) {}
  Scalar_(T v0, T v1, T v2 = 0, T v3 = 0) : Vec<T, 4>(v0, v1, v2, v3) {}
  static Scalar_ all(T v) { return Scalar_(v, v, v, v); }
};
typedef Scalar_<double> Scalar;

// ---- core/mat.hpp: Mat -------------------------------------------------------------------------------------------
class Mat;
namespace refshim {
inline size_t depthBytes(int depth) {
  static const size_t b[8] = {1, 1, 2, 2, 4, 4, 8, 0};
  return b[depth & 7];
}
}  // namespace refshim

class Mat {
 public:
  int flags = 0;   // CV_MAKETYPE(depth, cn)
  int dims = 0;    // 0 when empty (the reference tests `prevFlow.dims > 0`, PixFlow.h:90,101)
  int rows = 0, cols = 0;
  uchar* data = nullptr;
  size_t step = 0;   // bytes per row

  Mat() {}
  Mat(int rows_, int cols_, int type) { create(rows_, cols_, type); }
  Mat(Size sz, int type) { create(sz.height, sz.width, type); }
  Mat(int rows_, int cols_, int type, const Scalar& s) { create(rows_, cols_, type); setTo(s); }
  Mat(Size sz, int type, const Scalar& s) { create(sz.height, sz.width, type); setTo(s); }
  // user-allocated data (mat.inl.hpp): no ownership, continuous rows
  Mat(int rows_, int cols_, int type, void* data_) : flags(type & 0xfff), dims(2), rows(rows_), cols(cols_), data((uchar*)data_) { step = (size_t)cols * elemSize(); }
  Mat(const Mat& m, const Rect& roi) : flags(m.flags), dims(2), rows(roi.height), cols(roi.width), step(m.step), buf_(m.buf_) {
    if (roi.x < 0 || roi.y < 0 || roi.width < 0 || roi.height < 0 || roi.x + roi.width > m.cols || roi.y + roi.height > m.rows)
      refShimUnsupported("Mat ROI outside the image");
    data = m.data + (size_t)roi.y * m.step + (size_t)roi.x * m.elemSize();
    if (rows == 0 || cols == 0) { rows = cols = 0; }
  }
  // copies share the pixel buffer (reference counted), as cv::Mat does
  Mat(const Mat&) = default;
  Mat& operator=(const Mat&) = default;
  Mat(Mat&& m) noexcept { moveFrom(m); }
  Mat& operator=(Mat&& m) noexcept { if (this != &m) moveFrom(m); return *this; }

  // Mat::create: keeps the buffer when shape and type already match (mat.inl.hpp / matrix.cpp)
  void create(int rows_, int cols_, int type) {
    type &= 0xfff;
    if (dims == 2 && rows == rows_ && cols == cols_ && this->type() == type && data) return;
    if (rows_ < 0 || cols_ < 0) refShimUnsupported("negative Mat size");
    release();
    flags = type;
    dims = 2;
    rows = rows_;
    cols = cols_;
    step = (size_t)cols * elemSize();
    const size_t total = step * (size_t)rows;
    if (total > 0) {
      buf_ = std::shared_ptr<uchar>((uchar*)std::malloc(total + 64), std::free);
      if (!buf_) refShimUnsupported("out of memory");
      data = buf_.get();
    }
  }
  void create(Size sz, int type) { create(sz.height, sz.width, type); }
  void release() { buf_.reset(); data = nullptr; rows = cols = 0; dims = 0; step = 0; }

  int type() const { return flags & 0xfff; }
  int depth() const { return CV_MAT_DEPTH(flags); }
  int channels() const { return CV_MAT_CN(flags); }
  size_t elemSize() const { return refshim::depthBytes(depth()) * channels(); }
  size_t elemSize1() const { return refshim::depthBytes(depth()); }
  bool empty() const { return data == nullptr || rows == 0 || cols == 0; }
  size_t total() const { return (size_t)rows * cols; }
  Size size() const { return Size(cols, rows); }
  bool isContinuous() const { return rows <= 1 || step == (size_t)cols * elemSize(); }

  template <typename T> T* ptr(int y = 0) { return (T*)(data + (size_t)y * step); }
  template <typename T> const T* ptr(int y = 0) const { return (const T*)(data + (size_t)y * step); }
  uchar* ptr(int y = 0) { return data + (size_t)y * step; }
  const uchar* ptr(int y = 0) const { return data + (size_t)y * step; }
  template <typename T> T& at(int y, int x) { return ((T*)(data + (size_t)y * step))[x]; }
  template <typename T> const T& at(int y, int x) const { return ((const T*)(data + (size_t)y * step))[x]; }
  // at(i0): element i0 of a single-row or single-column matrix (mat.inl.hpp)
  template <typename T> T& at(int i0) {
    if (isContinuous() || rows == 1) return ((T*)data)[i0];
    if (cols == 1) return *(T*)(data + step * (size_t)i0);
    const int i = i0 / cols, j = i0 - i * cols;
    return ((T*)(data + step * (size_t)i))[j];
  }
  template <typename T> const T& at(int i0) const { return const_cast<Mat*>(this)->at<T>(i0); }

  Mat operator()(const Rect& roi) const { return Mat(*this, roi); }
  Mat clone() const { Mat m; copyTo(m); return m; }
  void copyTo(Mat& dst) const {
    if (empty()) { dst.release(); return; }
    if (dst.data == data && dst.rows == rows && dst.cols == cols && dst.step == step) return;
    Mat out;
    out.create(rows, cols, type());
    const size_t rb = (size_t)cols * elemSize();
    for (int y = 0; y < rows; ++y) std::memcpy(out.ptr(y), ptr(y), rb);
    dst = std::move(out);
  }
  Mat& setTo(const Scalar& s) {
    const int cn = channels();
    for (int y = 0; y < rows; ++y)
      for (int x = 0; x < cols; ++x)
        for (int c = 0; c < cn; ++c) {
          const double v = c < 4 ? s.val[c] : 0.0;
          switch (depth()) {
            case CV_8U: ptr<uchar>(y)[x * cn + c] = saturate_cast<uchar>(v); break;
            case CV_16U: ptr<ushort>(y)[x * cn + c] = saturate_cast<ushort>(v); break;
            case CV_32F: ptr<float>(y)[x * cn + c] = (float)v; break;
            case CV_64F: ptr<double>(y)[x * cn + c] = v; break;
            default: refShimUnsupported("setTo depth");
          }
        }
    return *this;
  }
  static Mat zeros(int rows_, int cols_, int type) {
    Mat m(rows_, cols_, type);
    for (int y = 0; y < m.rows; ++y) std::memset(m.ptr(y), 0, (size_t)m.cols * m.elemSize());
    return m;
  }
  static Mat zeros(Size sz, int type) { return zeros(sz.height, sz.width, type); }
  static Mat eye(int rows_, int cols_, int type) {   // single-channel identity (setIdentity)
    Mat m = zeros(rows_, cols_, type);
    for (int i = 0; i < std::min(rows_, cols_); ++i) {
      if (m.type() == CV_32FC1) m.at<float>(i, i) = 1.0f;
      else if (m.type() == CV_64FC1) m.at<double>(i, i) = 1.0;
      else refShimUnsupported("Mat::eye type");
    }
    return m;
  }

  // convertTo (convert.cpp): dst = saturate_cast<rtype>(src * alpha + beta); the 32F kernels take alpha and beta
  // as float (cvtScale_<.., float>) and skip the scaling when alpha == 1 and beta == 0 (plain cvt).
  void convertTo(Mat& dst, int rtype, double alpha = 1, double beta = 0) const;
  Mat mul(const Mat& m, double scale = 1) const;

 private:
  void moveFrom(Mat& m) {
    flags = m.flags; dims = m.dims; rows = m.rows; cols = m.cols; data = m.data; step = m.step;
    buf_ = std::move(m.buf_);
    m.data = nullptr; m.rows = m.cols = 0; m.dims = 0; m.step = 0;
  }
  std::shared_ptr<uchar> buf_;
};

inline void Mat::convertTo(Mat& dst, int rtype, double alpha, double beta) const {
  if (empty()) { dst.release(); return; }
  const int sdepth = depth(), cn = channels();
  const int ddepth = rtype < 0 ? sdepth : CV_MAT_DEPTH(rtype);
  const bool noScale = alpha == 1.0 && beta == 0.0;
  Mat out;
  out.create(rows, cols, CV_MAKETYPE(ddepth, cn));
  const float a = (float)alpha, b = (float)beta;
  for (int y = 0; y < rows; ++y) {
    const int n = cols * cn;
    if (sdepth == CV_8U && ddepth == CV_32F) {
      const uchar* s = ptr<uchar>(y); float* d = out.ptr<float>(y);
      for (int i = 0; i < n; ++i) d[i] = noScale ? (float)s[i] : (float)s[i] * a + b;
    } else if (sdepth == CV_32F && ddepth == CV_32F) {
      const float* s = ptr<float>(y); float* d = out.ptr<float>(y);
      for (int i = 0; i < n; ++i) d[i] = noScale ? s[i] : s[i] * a + b;
    } else if (sdepth == CV_32F && ddepth == CV_8U) {
      const float* s = ptr<float>(y); uchar* d = out.ptr<uchar>(y);
      for (int i = 0; i < n; ++i) d[i] = saturate_cast<uchar>(noScale ? s[i] : s[i] * a + b);
    } else if (sdepth == CV_8U && ddepth == CV_8U) {
      const uchar* s = ptr<uchar>(y); uchar* d = out.ptr<uchar>(y);
      for (int i = 0; i < n; ++i) d[i] = noScale ? s[i] : saturate_cast<uchar>((float)s[i] * a + b);
    } else {
      refShimUnsupported("convertTo depth combination");
    }
  }
  dst = std::move(out);
}
// `m *= s`, `m /= s`, `m * s` (mat.inl.hpp / matop.cpp): convertTo(m, -1, s) resp. convertTo(m, -1, 1./s)
static inline Mat& operator*=(Mat& a, double s) { a.convertTo(a, -1, s); return a; }
static inline Mat& operator/=(Mat& a, double s) { a.convertTo(a, -1, 1. / s); return a; }
static inline Mat operator*(const Mat& a, double s) { Mat r; a.convertTo(r, -1, s); return r; }
static inline Mat operator*(double s, const Mat& a) { Mat r; a.convertTo(r, -1, s); return r; }
static inline Mat operator/(const Mat& a, double s) { Mat r; a.convertTo(r, -1, 1. / s); return r; }
inline Mat Mat::mul(const Mat& m, double scale) const {
  if (type() != CV_32FC1 || m.type() != CV_32FC1 || m.size() != size()) refShimUnsupported("Mat::mul types");
  Mat r(rows, cols, type());
  const float sc = (float)scale;
  for (int y = 0; y < rows; ++y)
    for (int x = 0; x < cols; ++x) r.at<float>(y, x) = scale == 1.0 ? at<float>(y, x) * m.at<float>(y, x) : sc * at<float>(y, x) * m.at<float>(y, x);
  return r;
}

// Matrix product of small single-precision matrices (`A * B`, `A *= B` -> cv::gemm(A, B, 1, noArray(), 0), matmul.simd.hpp): for
// 2 <= len <= 4 gemm's fixed-size branch forms every element as the float sum a0*b0 + a1*b1 + a2*b2 (left to right) and stores
// (float)(t * alpha + c * beta) with alpha = 1.0, beta = 0.0 in double -- i.e. t itself. Pinned against cv2.gemm in tests/.
static inline Mat operator*(const Mat& a, const Mat& b) {
  if (a.type() != CV_32FC1 || b.type() != CV_32FC1 || a.cols != b.rows || a.cols < 2 || a.cols > 4 || b.cols != a.cols || a.rows != a.cols)
    refShimUnsupported("Mat * Mat: only square 2x2..4x4 CV_32F products are restated");
  Mat d(a.rows, b.cols, CV_32FC1);
  for (int i = 0; i < a.rows; ++i)
    for (int j = 0; j < b.cols; ++j) {
      float t = a.at<float>(i, 0) * b.at<float>(0, j);
      for (int k = 1; k < a.cols; ++k) t = t + a.at<float>(i, k) * b.at<float>(k, j);
      d.at<float>(i, j) = (float)((double)t * 1.0 + 0.0 * 0.0);
    }
  return d;
}
static inline Mat& operator*=(Mat& a, const Mat& b) { Mat d = a * b; a = d; return a; }
static inline void transpose(const Mat& src, Mat& dst) {
  Mat out(src.cols, src.rows, src.type());
  const size_t es = src.elemSize();
  for (int y = 0; y < src.rows; ++y)
    for (int x = 0; x < src.cols; ++x) std::memcpy(out.ptr(x) + (size_t)y * es, src.ptr(y) + (size_t)x * es, es);
  dst = out;
}
// cv::dct / cv::idct are not restated: the frequency-domain demosaic (camera_isp/CameraIsp.h:1176-1196) is outside the built path
static inline void dct(const Mat&, Mat&) { refShimUnsupported("cv::dct"); }
static inline void idct(const Mat&, Mat&) { refShimUnsupported("cv::idct"); }
// stream output of the types the reference logs (format irrelevant: verbose logging only)
template <typename T> static inline std::ostream& operator<<(std::ostream& o, const Point3_<T>& p) { return o << "[" << p.x << ", " << p.y << ", " << p.z << "]"; }
template <typename T> static inline std::ostream& operator<<(std::ostream& o, const std::vector<Point3_<T>>& v) { o << "["; for (const auto& p : v) o << p << ";"; return o << "]"; }
static inline std::ostream& operator<<(std::ostream& o, const Mat& m) { return o << "Mat(" << m.rows << "x" << m.cols << ")"; }

// Output arguments: OpenCV's OutputArray also binds to a `const Mat` lvalue (the reference flips `const vector<Mat>`
// elements in place, CvUtil.cpp:120-127).
struct OutputArray {
  Mat* m;
  OutputArray(Mat& m_) : m(&m_) {}
  OutputArray(const Mat& m_) : m(const_cast<Mat*>(&m_)) {}
  Mat& ref() const { return *m; }
};

namespace detail {}   // `using namespace cv::detail;` (PixFlow.h:29)

// ---- core.hpp free functions ----------------------------------------------------------------------------------------
enum BorderTypes { BORDER_CONSTANT = 0, BORDER_REPLICATE = 1, BORDER_REFLECT = 2, BORDER_WRAP = 3, BORDER_REFLECT_101 = 4, BORDER_REFLECT101 = 4, BORDER_DEFAULT = 4 };
void split(const Mat& src, std::vector<Mat>& mv);
void merge(const std::vector<Mat>& mv, OutputArray dst);
void flip(const Mat& src, OutputArray dst, int flipCode);
void hconcat(const Mat& a, const Mat& b, OutputArray dst);
void vconcat(const Mat& a, const Mat& b, OutputArray dst);
void copyMakeBorder(const Mat& src, OutputArray dst, int top, int bottom, int left, int right, int borderType, const Scalar& value = Scalar());

}  // namespace cv
