// ORACLE — TEST INFRASTRUCTURE ONLY. Not part of the product path.
//
// CPU restatement (plain C++, no OpenCV) of the third-party image primitives the
// Surround360 render path calls. The reference delegates these to OpenCV (pinned at
// Itseez/opencv f109c01, WITH_IPP=OFF, surround360_render/README.md:144-152), which is
// absent from /root/reference; each function restates the published OpenCV algorithm
// and cites the reference call site it serves. Pinning: tests/test_oracle_cv.py checks
// every primitive against python cv2 4.13 (IPP off) in this container and commits the
// resulting vectors under tests/golden/.
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
// legs may include, link or call anything in this directory.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <vector>

namespace orc {

template <class T>
struct Img {
  int w = 0, h = 0, c = 0;
  std::vector<T> d;
  Img() {}
  Img(int w_, int h_, int c_) : w(w_), h(h_), c(c_), d((size_t)w_ * h_ * c_) {}
  Img(int w_, int h_, int c_, T v) : w(w_), h(h_), c(c_), d((size_t)w_ * h_ * c_, v) {}
  T* p(int y, int x = 0) { return d.data() + ((size_t)y * w + x) * c; }
  const T* p(int y, int x = 0) const { return d.data() + ((size_t)y * w + x) * c; }
  bool empty() const { return d.empty(); }
  size_t step() const { return (size_t)w * c; }
};
using Img8 = Img<uint8_t>;
using ImgF = Img<float>;

// ---- scalar helpers with OpenCV semantics -------------------------------------------
static inline int cvFloor(double v) {
  int i = (int)v;
  return i - (i > v);
}
// cvRound = SSE cvtsd2si: round-half-even, INT_MIN on overflow / NaN
static inline int cvRound(double v) {
  if (!(v > -2147483648.5 && v < 2147483647.5)) return std::numeric_limits<int>::min();
  return (int)std::nearbyint(v);
}
static inline short sat_short(int v) { return (short)(v < -32768 ? -32768 : v > 32767 ? 32767 : v); }
static inline uint8_t sat_u8(int v) { return (uint8_t)(v < 0 ? 0 : v > 255 ? 255 : v); }
static inline int clipi(int x, int a, int b) { return x >= a ? (x < b ? x : b - 1) : a; }  // [a,b)

// cv::borderInterpolate
enum Border { B_CONSTANT = 0, B_REPLICATE = 1, B_REFLECT101 = 4, B_WRAP = 3 };
static inline int borderInterp(int p, int len, int type) {
  if ((unsigned)p < (unsigned)len) return p;
  if (type == B_REPLICATE) return p < 0 ? 0 : len - 1;
  if (type == B_REFLECT101) {
    if (len == 1) return 0;
    do {
      if (p < 0) p = -p;
      else p = len - 1 - (p - len) - 1;
    } while ((unsigned)p >= (unsigned)len);
    return p;
  }
  if (type == B_WRAP) {
    if (p < 0) p -= ((p - len + 1) / len) * len;
    if (p >= len) p %= len;
    return p;
  }
  return -1;  // constant
}

// cv interpolateCubic, A = -0.75 (imgwarp/resize). float arithmetic, no contraction.
static inline void cubicCoeffs(float x, float* k) {
  const float A = -0.75f;
  k[0] = ((A * (x + 1) - 5 * A) * (x + 1) + 8 * A) * (x + 1) - 4 * A;
  k[1] = ((A + 2) * x - (A + 3)) * x * x + 1;
  k[2] = ((A + 2) * (1 - x) - (A + 3)) * (1 - x) * (1 - x) + 1;
  k[3] = 1.f - k[0] - k[1] - k[2];
}

// ---- cv::getGaussianKernel(n, sigma, CV_32F) -----------------------------------------
static inline std::vector<float> gaussKernel(int n, double sigma) {
  std::vector<double> t(n);
  double scale2X = -0.5 / (sigma * sigma), sum = 0;
  for (int i = 0; i < n; ++i) {
    double x = i - (n - 1) * 0.5;
    t[i] = std::exp(scale2X * x * x);
    sum += t[i];
  }
  sum = 1. / sum;
  std::vector<float> k(n);
  for (int i = 0; i < n; ++i) k[i] = (float)(t[i] * sum);
  return k;
}

// ---- cv::GaussianBlur on CV_32FC{1,2}, BORDER_REFLECT_101 ----------------------------
// sepFilter2D float path: row pass then column pass, symmetric kernel.
// Summation order restated from OpenCV filter.simd.hpp: row ksize<=5 -> centre first then
// pairs ascending; row ksize>5 -> left-to-right taps; column -> centre first then pairs.
// Reference call sites: PixFlow.h:137-138,363-366,379-383,439-443,178-182.
static inline void gaussianBlurF(const ImgF& src, ImgF& dst, int ksize, double sigma) {
  const int w = src.w, h = src.h, cn = src.c, r = ksize / 2;
  std::vector<float> k = gaussKernel(ksize, sigma);
  ImgF tmp(w, h, cn);
  std::vector<int> xi(w + 2 * r);
  for (int i = 0; i < w + 2 * r; ++i) xi[i] = borderInterp(i - r, w, B_REFLECT101);
  for (int y = 0; y < h; ++y) {
    const float* S = src.p(y);
    float* D = tmp.p(y);
    for (int x = 0; x < w; ++x)
      for (int ch = 0; ch < cn; ++ch) {
        float s;
        if (ksize <= 5) {
          s = S[xi[x + r] * cn + ch] * k[r];
          for (int j = 1; j <= r; ++j)
            s = s + (S[xi[x + r - j] * cn + ch] + S[xi[x + r + j] * cn + ch]) * k[r + j];
        } else {
          s = k[0] * S[xi[x] * cn + ch];
          for (int j = 1; j < ksize; ++j) s = s + k[j] * S[xi[x + j] * cn + ch];
        }
        D[x * cn + ch] = s;
      }
  }
  ImgF out(w, h, cn);
  for (int y = 0; y < h; ++y) {
    float* D = out.p(y);
    const float* S0 = tmp.p(y);
    for (int i = 0; i < w * cn; ++i) {
      float s = k[r] * S0[i];
      for (int j = 1; j <= r; ++j) {
        const float* Sa = tmp.p(borderInterp(y + j, h, B_REFLECT101));
        const float* Sb = tmp.p(borderInterp(y - j, h, B_REFLECT101));
        s = s + k[r + j] * (Sa[i] + Sb[i]);
      }
      D[i] = s;
    }
  }
  dst = std::move(out);
}

// ---- cv::Sobel(ksize=1), BORDER_REPLICATE: central difference (PixFlow.h:356-359) -----
static inline void sobel1(const ImgF& src, ImgF& dst, bool dx) {
  ImgF out(src.w, src.h, 1);
  for (int y = 0; y < src.h; ++y)
    for (int x = 0; x < src.w; ++x) {
      if (dx) {
        int x0 = std::max(x - 1, 0), x1 = std::min(x + 1, src.w - 1);
        out.p(y)[x] = src.p(y)[x1] - src.p(y)[x0];
      } else {
        int y0 = std::max(y - 1, 0), y1 = std::min(y + 1, src.h - 1);
        out.p(y)[x] = src.p(y1)[x] - src.p(y0)[x];
      }
    }
  dst = std::move(out);
}

// ---- cv::medianBlur(CV_32FC2, 5): per-channel 5x5 median, replicate border -----------
// (PixFlow.h:398,411)
static inline void median5(const ImgF& src, ImgF& dst) {
  const int w = src.w, h = src.h, cn = src.c;
  ImgF out(w, h, cn);
  float v[25];
  for (int y = 0; y < h; ++y)
    for (int x = 0; x < w; ++x)
      for (int ch = 0; ch < cn; ++ch) {
        int n = 0;
        for (int dy = -2; dy <= 2; ++dy) {
          const float* S = src.p(std::min(std::max(y + dy, 0), h - 1));
          for (int dx2 = -2; dx2 <= 2; ++dx2) v[n++] = S[std::min(std::max(x + dx2, 0), w - 1) * cn + ch];
        }
        std::nth_element(v, v + 12, v + 25);
        out.p(y, x)[ch] = v[12];
      }
  dst = std::move(out);
}

// ---- cv::resize coefficient tables (resize.cpp) ---------------------------------------
struct ResizeTab {
  std::vector<int> ofs;      // base source index (per dst coord)
  std::vector<float> coef;   // ksize coefficients per dst coord
  int lo, hi;                // xmin,xmax: coords outside [lo,hi) touch the border
};
static inline ResizeTab resizeTab(int ssize, int dsize, int ksize, bool isX) {
  ResizeTab t;
  t.ofs.resize(dsize);
  t.coef.resize((size_t)dsize * ksize);
  double inv_scale = (double)dsize / ssize;
  double scale = 1. / inv_scale;
  int ksize2 = ksize / 2;
  t.lo = 0;
  t.hi = dsize;
  for (int d = 0; d < dsize; ++d) {
    float f = (float)((d + 0.5) * scale - 0.5);
    int s = cvFloor(f);
    f -= s;
    if (isX) {
      if (s < ksize2 - 1) {
        t.lo = d + 1;
        if (s < 0 && ksize == 2) f = 0, s = 0;
      }
      if (s + ksize2 >= ssize) {
        t.hi = std::min(t.hi, d);
        if (s >= ssize - 1 && ksize == 2) f = 0, s = ssize - 1;
      }
    }
    t.ofs[d] = s;
    float* cb = &t.coef[(size_t)d * ksize];
    if (ksize == 2) {
      cb[0] = 1.f - f;
      cb[1] = f;
    } else {
      cubicCoeffs(f, cb);
    }
  }
  return t;
}

// ---- cv::resize INTER_LINEAR on CV_32FC{1,2} (PixFlow.h:487,176) ----------------------
static inline void resizeLinearF(const ImgF& src, ImgF& dst, int dw, int dh) {
  const int cn = src.c;
  ResizeTab tx = resizeTab(src.w, dw, 2, true), ty = resizeTab(src.h, dh, 2, false);
  ImgF out(dw, dh, cn);
  std::vector<float> r0((size_t)dw * cn), r1((size_t)dw * cn);
  auto hrow = [&](int sy, std::vector<float>& R) {
    const float* S = src.p(sy);
    for (int dx = 0; dx < dw; ++dx) {
      int sx = tx.ofs[dx];
      const float* a = &tx.coef[(size_t)dx * 2];
      for (int ch = 0; ch < cn; ++ch) {
        if (dx < tx.hi) R[dx * cn + ch] = S[sx * cn + ch] * a[0] + S[(sx + 1) * cn + ch] * a[1];
        else R[dx * cn + ch] = S[sx * cn + ch];
      }
    }
  };
  for (int dy = 0; dy < dh; ++dy) {
    int sy0 = clipi(ty.ofs[dy], 0, src.h), sy1 = clipi(ty.ofs[dy] + 1, 0, src.h);
    hrow(sy0, r0);
    hrow(sy1, r1);
    float b0 = ty.coef[(size_t)dy * 2], b1 = ty.coef[(size_t)dy * 2 + 1];
    float* D = out.p(dy);
    for (int i = 0; i < dw * cn; ++i) D[i] = r0[i] * b0 + r1[i] * b1;
  }
  dst = std::move(out);
}

// ---- cv::resize INTER_CUBIC on CV_32FC2 (PixFlow.h:170; prev-flow :103) ---------------
// Vertical pass: the SIMD body (first dw*cn & ~3 elements of a row, VResizeCubicVec_32f)
// associates S0*b0 + (S1*b1 + (S2*b2 + S3*b3)); the scalar tail associates left-to-right.
static inline void resizeCubicF(const ImgF& src, ImgF& dst, int dw, int dh) {
  const int cn = src.c;
  ResizeTab tx = resizeTab(src.w, dw, 4, true), ty = resizeTab(src.h, dh, 4, false);
  ImgF out(dw, dh, cn);
  std::vector<float> R[4];
  for (auto& r : R) r.resize((size_t)dw * cn);
  auto hrow = [&](int sy, std::vector<float>& Rr) {
    const float* S = src.p(sy);
    for (int dx = 0; dx < dw; ++dx) {
      const float* a = &tx.coef[(size_t)dx * 4];
      for (int ch = 0; ch < cn; ++ch) {
        float v = 0;
        for (int j = 0; j < 4; ++j) {
          int sx = std::min(std::max(tx.ofs[dx] - 1 + j, 0), src.w - 1);
          v = v + S[sx * cn + ch] * a[j];
        }
        Rr[dx * cn + ch] = v;
      }
    }
  };
  for (int dy = 0; dy < dh; ++dy) {
    for (int k = 0; k < 4; ++k) hrow(clipi(ty.ofs[dy] - 1 + k, 0, src.h), R[k]);
    const float* b = &ty.coef[(size_t)dy * 4];
    float* D = out.p(dy);
    const int simdEnd = (dw * cn) & ~3;
    for (int i = 0; i < dw * cn; ++i)
      D[i] = i < simdEnd ? R[0][i] * b[0] + (R[1][i] * b[1] + (R[2][i] * b[2] + R[3][i] * b[3]))
                         : R[0][i] * b[0] + R[1][i] * b[1] + R[2][i] * b[2] + R[3][i] * b[3];
  }
  dst = std::move(out);
}

// ---- cv::resize INTER_CUBIC on CV_8UC{3,4} (PixFlow.h:98-99; TRSP:943-956) ------------
// 11-bit fixed-point coefficients, integer horizontal pass. Vertical pass: OpenCV's
// VResizeCubicVec_32s8u computes in float (beta * 2^-22, association S0*b0+(S1*b1+(S2*b2+
// S3*b3)), round-half-even) for the first (dw*cn & ~7) elements of each row and in exact
// integers ((sum + 2^21) >> 22) for the tail.
static inline void resizeCubic8(const Img8& src, Img8& dst, int dw, int dh) {
  const int cn = src.c;
  ResizeTab tx = resizeTab(src.w, dw, 4, true), ty = resizeTab(src.h, dh, 4, false);
  std::vector<short> ia((size_t)dw * 4), ib((size_t)dh * 4);
  auto toShort = [](float v) { return sat_short(cvRound(v * 2048.f)); };
  for (size_t i = 0; i < ia.size(); ++i) ia[i] = toShort(tx.coef[i]);
  for (size_t i = 0; i < ib.size(); ++i) ib[i] = toShort(ty.coef[i]);
  Img8 out(dw, dh, cn);
  std::vector<int> R[4];
  for (auto& r : R) r.resize((size_t)dw * cn);
  auto hrow = [&](int sy, std::vector<int>& Rr) {
    const uint8_t* S = src.p(sy);
    for (int dx = 0; dx < dw; ++dx) {
      const short* a = &ia[(size_t)dx * 4];
      for (int ch = 0; ch < cn; ++ch) {
        int v = 0;
        for (int j = 0; j < 4; ++j) {
          int sx = std::min(std::max(tx.ofs[dx] - 1 + j, 0), src.w - 1);
          v += S[sx * cn + ch] * a[j];
        }
        Rr[dx * cn + ch] = v;
      }
    }
  };
  const float scale = 1.f / (2048 * 2048);
  const int simdEnd = (dw * cn) & ~7;
  for (int dy = 0; dy < dh; ++dy) {
    for (int k = 0; k < 4; ++k) hrow(clipi(ty.ofs[dy] - 1 + k, 0, src.h), R[k]);
    const short* b = &ib[(size_t)dy * 4];
    const float fb0 = b[0] * scale, fb1 = b[1] * scale, fb2 = b[2] * scale, fb3 = b[3] * scale;
    uint8_t* D = out.p(dy);
    for (int i = 0; i < dw * cn; ++i) {
      if (i < simdEnd) {
        float f = (float)R[0][i] * fb0 + ((float)R[1][i] * fb1 + ((float)R[2][i] * fb2 + (float)R[3][i] * fb3));
        D[i] = sat_u8(cvRound(f));
      } else {
        int s = R[0][i] * b[0] + R[1][i] * b[1] + R[2][i] * b[2] + R[3][i] * b[3];
        D[i] = sat_u8((s + (1 << 21)) >> 22);
      }
    }
  }
  dst = std::move(out);
}

// ---- cv::remap interpolation tables (imgwarp.cpp initInterTab2D) ----------------------
struct RemapTabs {
  float f[1024][16];
  short i[1024][16];
  RemapTabs() {
    float t1[32][4];
    for (int k = 0; k < 32; ++k) cubicCoeffs(k * (1.f / 32), t1[k]);
    for (int yi = 0; yi < 32; ++yi)
      for (int xi = 0; xi < 32; ++xi) {
        float* tf = f[yi * 32 + xi];
        short* ti = i[yi * 32 + xi];
        int isum = 0;
        for (int k1 = 0; k1 < 4; ++k1)
          for (int k2 = 0; k2 < 4; ++k2) {
            float v = t1[yi][k1] * t1[xi][k2];
            tf[k1 * 4 + k2] = v;
            isum += ti[k1 * 4 + k2] = sat_short(cvRound(v * 32768.f));
          }
        if (isum != 32768) {
          int diff = isum - 32768;
          int ks2 = 2, Mk1 = ks2, Mk2 = ks2, mk1 = ks2, mk2 = ks2;
          for (int k1 = ks2; k1 < ks2 + 2; ++k1)
            for (int k2 = ks2; k2 < ks2 + 2; ++k2) {
              if (ti[k1 * 4 + k2] < ti[mk1 * 4 + mk2]) mk1 = k1, mk2 = k2;
              else if (ti[k1 * 4 + k2] > ti[Mk1 * 4 + Mk2]) Mk1 = k1, Mk2 = k2;
            }
          if (diff < 0) ti[Mk1 * 4 + Mk2] = (short)(ti[Mk1 * 4 + Mk2] - diff);
          else ti[mk1 * 4 + mk2] = (short)(ti[mk1 * 4 + mk2] - diff);
        }
      }
  }
};
static inline const RemapTabs& remapTabs() {
  static RemapTabs t;
  return t;
}

// map coordinate -> (integer base, table index); cv::remap's 1/32-px quantisation
static inline void quantMap(float mx, float my, int& sx, int& sy, int& tab) {
  int ix = cvRound(mx * 32.f), iy = cvRound(my * 32.f);
  tab = (iy & 31) * 32 + (ix & 31);
  sx = sat_short(ix >> 5);
  sy = sat_short(iy >> 5);
}

// ---- cv::remap INTER_CUBIC, CV_8UC{3,4}, BORDER_CONSTANT(0) or BORDER_WRAP ------------
// map: dw x dh x 2 floats (x,y). Reference call sites: ImageWarper.cpp:173,131-137;
// NovelView.cpp:206; TRSP:497-503.
static inline void remapCubic8(const Img8& src, const ImgF& map, Img8& dst, int border) {
  const RemapTabs& T = remapTabs();
  const int cn = src.c, dw = map.w, dh = map.h;
  Img8 out(dw, dh, cn);
  const int w1 = std::max(src.w - 3, 0), h1 = std::max(src.h - 3, 0);
  for (int y = 0; y < dh; ++y)
    for (int x = 0; x < dw; ++x) {
      int sx, sy, tab;
      quantMap(map.p(y, x)[0], map.p(y, x)[1], sx, sy, tab);
      sx -= 1;
      sy -= 1;
      const short* w = T.i[tab];
      uint8_t* D = out.p(y, x);
      if ((unsigned)sx < (unsigned)w1 && (unsigned)sy < (unsigned)h1) {
        for (int k = 0; k < cn; ++k) {
          int sum = 0;
          for (int r = 0; r < 4; ++r) {
            const uint8_t* S = src.p(sy + r, sx) + k;
            sum += S[0] * w[r * 4] + S[cn] * w[r * 4 + 1] + S[cn * 2] * w[r * 4 + 2] + S[cn * 3] * w[r * 4 + 3];
          }
          D[k] = sat_u8((sum + (1 << 14)) >> 15);
        }
      } else {
        if (border == B_CONSTANT && (sx >= src.w || sx + 4 <= 0 || sy >= src.h || sy + 4 <= 0)) {
          for (int k = 0; k < cn; ++k) D[k] = 0;
          continue;
        }
        int xs[4], ys[4];
        for (int i = 0; i < 4; ++i) {
          xs[i] = borderInterp(sx + i, src.w, border);
          ys[i] = borderInterp(sy + i, src.h, border);
        }
        for (int k = 0; k < cn; ++k) {
          int sum = 0;
          for (int r = 0; r < 4; ++r) {
            if (ys[r] < 0) continue;
            const uint8_t* S = src.p(ys[r]) + k;
            for (int j = 0; j < 4; ++j)
              if (xs[j] >= 0) sum += S[xs[j] * cn] * w[r * 4 + j];
          }
          D[k] = sat_u8((sum + (1 << 14)) >> 15);
        }
      }
    }
  dst = std::move(out);
}

// ---- cv::remap INTER_CUBIC, CV_32FC2, BORDER_CONSTANT(0) (NovelView.cpp:191) ----------
// float weight table; interior = 4 row sums, border = sequential in-range taps (imgwarp.cpp remapBicubic).
static inline void remapCubicF(const ImgF& src, const ImgF& map, ImgF& dst) {
  const RemapTabs& T = remapTabs();
  const int cn = src.c, dw = map.w, dh = map.h;
  ImgF out(dw, dh, cn);
  for (int y = 0; y < dh; ++y)
    for (int x = 0; x < dw; ++x) {
      int sx, sy, tab;
      quantMap(map.p(y, x)[0], map.p(y, x)[1], sx, sy, tab);
      sx -= 1;
      sy -= 1;
      const float* w = T.f[tab];
      float* D = out.p(y, x);
      if (sx >= src.w || sx + 4 <= 0 || sy >= src.h || sy + 4 <= 0) {
        for (int k = 0; k < cn; ++k) D[k] = 0.f;
        continue;
      }
      if ((unsigned)sx < (unsigned)std::max(src.w - 3, 0) && (unsigned)sy < (unsigned)std::max(src.h - 3, 0)) {
        // interior: four row sums (each left-to-right), accumulated top to bottom
        for (int k = 0; k < cn; ++k) {
          const float* S = src.p(sy, sx) + k;
          const size_t st = src.step();
          float sum = S[0] * w[0] + S[cn] * w[1] + S[cn * 2] * w[2] + S[cn * 3] * w[3];
          S += st;
          sum += S[0] * w[4] + S[cn] * w[5] + S[cn * 2] * w[6] + S[cn * 3] * w[7];
          S += st;
          sum += S[0] * w[8] + S[cn] * w[9] + S[cn * 2] * w[10] + S[cn * 3] * w[11];
          S += st;
          sum += S[0] * w[12] + S[cn] * w[13] + S[cn * 2] * w[14] + S[cn * 3] * w[15];
          D[k] = sum;
        }
        continue;
      }
      // border: sequential accumulation of the in-range taps
      for (int k = 0; k < cn; ++k) {
        float sum = 0.f;
        for (int r = 0; r < 4; ++r) {
          int yy = sy + r;
          if ((unsigned)yy >= (unsigned)src.h) continue;
          const float* S = src.p(yy) + k;
          for (int j = 0; j < 4; ++j) {
            int xx = sx + j;
            if ((unsigned)xx < (unsigned)src.w) sum = sum + S[xx * cn] * w[r * 4 + j];
          }
        }
        D[k] = sum;
      }
    }
  dst = std::move(out);
}

// ---- cv::cvtColor BGRA2GRAY / BGR2GRAY 8U (PixFlow.h:125-126) -------------------------
// cv2 4.13: 15-bit coefficients. The 2016 pin used 14-bit (B*1868+G*9617+R*4899+2^13)>>14;
// selectable for documentation of the drift.
static inline uint8_t gray8(int b, int g, int r, bool legacy14 = false) {
  if (legacy14) return (uint8_t)((b * 1868 + g * 9617 + r * 4899 + (1 << 13)) >> 14);
  return (uint8_t)((b * 3735 + g * 19235 + r * 9798 + (1 << 14)) >> 15);
}

// ---- cv::erode with MORPH_CROSS (2r+1)^2, anchor centre (CvUtil.cpp:144-149) ----------
// default border for erode = +inf (never lowers the min)
static inline void erodeCross8(const Img8& src, Img8& dst, int r) {
  Img8 out(src.w, src.h, 1);
  for (int y = 0; y < src.h; ++y)
    for (int x = 0; x < src.w; ++x) {
      int m = 255;
      for (int d = -r; d <= r; ++d) {
        int xx = x + d, yy = y + d;
        if (xx >= 0 && xx < src.w) m = std::min<int>(m, src.p(y)[xx]);
        if (yy >= 0 && yy < src.h) m = std::min<int>(m, src.p(yy)[x]);
      }
      out.p(y)[x] = (uint8_t)m;
    }
  dst = std::move(out);
}

// ---- cv::GaussianBlur on CV_8UC1 (CvUtil.cpp:152) --------------------------------------
// cv2 4.13 uses a bit-exact 8.8 fixed-point kernel (ufixedpoint16), BORDER_REFLECT_101:
// horizontal pass sums u8*coef in 16-bit fixed point, vertical pass sums in 32-bit and
// rounds (v + 2^15) >> 16. `fk` = the 8.8 kernel (sums to 256), supplied by the caller
// (tests pin it against cv2; see gaussKernelFixed8).
static inline void gaussianBlur8Fixed(const Img8& src, Img8& dst, const std::vector<int>& fk) {
  const int n = (int)fk.size(), r = n / 2, w = src.w, h = src.h;
  std::vector<uint16_t> tmp((size_t)w * h);
  for (int y = 0; y < h; ++y)
    for (int x = 0; x < w; ++x) {
      uint32_t s = 0;
      for (int j = 0; j < n; ++j) s += src.p(y)[borderInterp(x + j - r, w, B_REFLECT101)] * (uint32_t)fk[j];
      tmp[(size_t)y * w + x] = (uint16_t)std::min<uint32_t>(s, 65535u);
    }
  Img8 out(w, h, 1);
  for (int y = 0; y < h; ++y)
    for (int x = 0; x < w; ++x) {
      uint32_t s = 0;
      for (int j = 0; j < n; ++j)
        s += tmp[(size_t)borderInterp(y + j - r, h, B_REFLECT101) * w + x] * (uint32_t)fk[j];
      out.p(y)[x] = (uint8_t)std::min<uint32_t>((s + (1u << 15)) >> 16, 255u);
    }
  dst = std::move(out);
}

// 8.8 fixed-point Gaussian kernel: round(k*256) with the residual pushed to the centre
// so that the taps sum to exactly 256. (cv2's own rounding is pinned in the tests.)
static inline std::vector<int> gaussKernelFixed8(int n, double sigma) {
  std::vector<double> t(n);
  double scale2X = -0.5 / (sigma * sigma), sum = 0;
  for (int i = 0; i < n; ++i) {
    double x = i - (n - 1) * 0.5;
    t[i] = std::exp(scale2X * x * x);
    sum += t[i];
  }
  std::vector<int> k(n);
  int isum = 0;
  // error-diffusion rounding from the edge to the centre, symmetric (OpenCV
  // getGaussianKernelFixedPoint_ED restated)
  double err = 0;
  for (int i = 0; i < n / 2; ++i) {
    double v = t[i] / sum * 256.0 + err;
    int q = (int)std::floor(v + 0.5);
    err = v - q;
    k[i] = k[n - 1 - i] = q;
    isum += 2 * q;
  }
  k[n / 2] = 256 - isum;
  return k;
}

// ---- cv::remap INTER_NEAREST + BORDER_WRAP with the map of CvUtil.cpp:93-115 ----------
// offsetHorizontalWrap: srcX = x - offset wrapped once into [0,cols); nearest =
// round-half-even (saturate_cast<short>), then wrap of the rounded index.
static inline int offsetWrapIndex(int x, int cols, float offset) {
  float srcX = float(x) - offset;
  if (srcX < 0) srcX += cols;
  if (srcX >= cols) srcX -= cols;
  int sx = sat_short(cvRound(srcX));
  return borderInterp(sx, cols, B_WRAP);
}

}  // namespace orc
