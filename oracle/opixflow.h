// ORACLE — TEST INFRASTRUCTURE ONLY (see ocv.h header).
//
// Restates the PixFlow optical-flow solver:
//   surround360_render/source/optical_flow/PixFlow.h:81-534
//   presets: surround360_render/source/optical_flow/OpticalFlowFactory.h:25-61
// Scalar float arithmetic, no FMA contraction (reference build: -O3 -mavx, no -mfma,
// surround360_render/CMakeLists.txt:33-35). cv::Mat `/=` and `*=` with a scalar are
// convertTo(alpha) = multiplication by float(alpha) (OpenCV mat.inl.hpp).
#pragma once
#include <stdexcept>
#include <string>
#include "ocv.h"

namespace orc {

enum DirectionHint { HINT_UNKNOWN = 0, HINT_RIGHT = 1, HINT_DOWN = 2, HINT_LEFT = 3, HINT_UP = 4 };

struct PixFlowParams {
  float pyrScaleFactor = 0.9f;
  float smoothnessCoef = 0.001f;
  float verticalRegularizationCoef = 0.01f;
  float horizontalRegularizationCoef = 0.01f;
  float gradientStepSize = 0.5f;
  float downscaleFactor = 0.5f;
  float directionalRegularizationCoef = 0.0f;
  int maxPercentage = 0;
};

// OpticalFlowFactory.h:23-64
static inline PixFlowParams flowParamsByName(const std::string& name) {
  PixFlowParams p;
  if (name == "pixflow_low") return p;
  if (name == "pixflow_search_20") { p.maxPercentage = 20; return p; }
  throw std::runtime_error("unrecognized flow algorithm name: " + name);
}

struct PixFlow {
  static constexpr int kPyrMinImageSize = 24;
  static constexpr int kPyrMaxLevels = 1000;
  static constexpr float kGradEpsilon = 0.001f;
  static constexpr float kUpdateAlphaThreshold = 0.9f;
  PixFlowParams P;
  bool legacyGray = false;
  explicit PixFlow(const PixFlowParams& p) : P(p) {}

  // per-level dump hook for stage-wise parity tests (level, flow after the level)
  std::vector<ImgF>* levelDump = nullptr;

  static void scaleF(ImgF& m, float s) { for (float& v : m.d) v = v * s; }

  // PixFlow.h:477-491
  std::vector<ImgF> buildPyramid(const ImgF& src) const {
    std::vector<ImgF> pyr;
    pyr.push_back(src);
    while ((int)pyr.size() < kPyrMaxLevels) {
      int nw = (int)(pyr.back().w * P.pyrScaleFactor + 0.5f);
      int nh = (int)(pyr.back().h * P.pyrScaleFactor + 0.5f);
      if (nh <= kPyrMinImageSize || nw <= kPyrMinImageSize) break;
      ImgF s;
      resizeLinearF(pyr.back(), s, nw, nh);
      pyr.push_back(std::move(s));
    }
    return pyr;
  }

  // PixFlow.h:457-475
  static inline float bilinearExtend(const ImgF& img, float x, float y) {
    x = std::min(img.w - 2.0f, std::max(0.0f, x));
    y = std::min(img.h - 2.0f, std::max(0.0f, y));
    const int x0 = int(x), y0 = int(y);
    const float xR = x - float(x0), yR = y - float(y0);
    const float* p = img.p(y0);
    const float f00 = p[x0], f01 = p[x0 + img.w], f10 = p[x0 + 1], f11 = p[x0 + img.w + 1];
    const float a1 = f00, a2 = f10 - f00, a3 = f01 - f00, a4 = f00 + f11 - f10 - f01;
    return a1 + a2 * xR + a3 * yR + a4 * xR * yR;
  }

  struct Level {
    const ImgF *a0, *a1, *I0x, *I0y, *I1x, *I1y, *blurred;
    ImgF* flow;
    int cols, rows;
  };

  // PixFlow.h:493-534 (UseDirectionalRegularization = false in both presets)
  inline float errorFunction(const Level& L, int x, int y, float fx, float fy) const {
    const float matchX = x + fx, matchY = y + fy;
    const float i0x = L.I0x->p(y)[x], i0y = L.I0y->p(y)[x];
    const float i1x = bilinearExtend(*L.I1x, matchX, matchY);
    const float i1y = bilinearExtend(*L.I1y, matchX, matchY);
    const float dfx = L.blurred->p(y, x)[0] - fx, dfy = L.blurred->p(y, x)[1] - fy;
    const float smoothness = sqrtf(dfx * dfx + dfy * dfy);
    float err = sqrtf((i0x - i1x) * (i0x - i1x) + (i0y - i1y) * (i0y - i1y)) + smoothness * P.smoothnessCoef +
                P.verticalRegularizationCoef * fabsf(fy) / float(L.cols) +
                P.horizontalRegularizationCoef * fabsf(fx) / float(L.rows);
    return err;
  }

  // one pixel of a sweep: PixFlow.h:390-395 / 403-408 (dir = +1 forward, -1 backward)
  inline void sweepPixel(const Level& L, int x, int y, int dir) const {
    if (!(L.a0->p(y)[x] > kUpdateAlphaThreshold && L.a1->p(y)[x] > kUpdateAlphaThreshold)) return;
    float* f = L.flow->p(y, x);
    float currErr = errorFunction(L, x, y, f[0], f[1]);
    const int xn = x - dir, yn = y - dir;
    if (xn >= 0 && xn < L.cols) {
      const float* pf = L.flow->p(y, xn);
      const float px = pf[0], py = pf[1];
      const float e = errorFunction(L, x, y, px, py);
      if (e < currErr) { f[0] = px; f[1] = py; currErr = e; }
    }
    if (yn >= 0 && yn < L.rows) {
      const float* pf = L.flow->p(yn, x);
      const float px = pf[0], py = pf[1];
      const float e = errorFunction(L, x, y, px, py);
      if (e < currErr) { f[0] = px; f[1] = py; currErr = e; }
    }
    const float ex = errorFunction(L, x, y, f[0] + kGradEpsilon, f[1] + 0.0f);
    const float ey = errorFunction(L, x, y, f[0] + 0.0f, f[1] + kGradEpsilon);
    const float gx = (ex - currErr) / kGradEpsilon, gy = (ey - currErr) / kGradEpsilon;
    f[0] = f[0] - P.gradientStepSize * gx;
    f[1] = f[1] - P.gradientStepSize * gy;
  }

  // PixFlow.h:437-454
  void lowAlphaFlowDiffusion(const ImgF& a0, const ImgF& a1, ImgF& flow) const {
    ImgF blurred;
    gaussianBlurF(flow, blurred, 15, 8.0);
    for (int y = 0; y < flow.h; ++y)
      for (int x = 0; x < flow.w; ++x) {
        const float d = 1.0f - a0.p(y)[x] * a1.p(y)[x];
        float* f = flow.p(y, x);
        const float* b = blurred.p(y, x);
        f[0] = d * b[0] + (1.0f - d) * f[0];
        f[1] = d * b[1] + (1.0f - d) * f[1];
      }
  }

  // ---- PixFlow.h:219-342 (search initialisation, MaxPercentage > 0) ----
  int searchDistance() const { return (kPyrMinImageSize * P.maxPercentage + 50) / 100; }
  float patchError(const ImgF& i0, const ImgF& al0, int i0x, int i0y, const ImgF& i1, const ImgF& al1, int i1x,
                   int i1y) const {
    float sad = 0, alpha = 0;
    for (int dy = -2; dy <= 2; ++dy) {
      const int d0y = i0y + dy;
      if (0 <= d0y && d0y < i0.h) {
        const int d1y = std::min(std::max(i1y + dy, 0), i1.h - 1);
        for (int dx = -2; dx <= 2; ++dx) {
          const int d0x = i0x + dx;
          if (0 <= d0x && d0x < i0.w) {
            const int d1x = std::min(std::max(i1x + dx, 0), i1.w - 1);
            const float difference = i0.p(d0y)[d0x] - i1.p(d1y)[d1x];
            sad += std::abs(difference);
            alpha += al0.p(d0y)[d0x] * al1.p(d1y)[d1x];
          }
        }
      }
    }
    sad /= alpha;
    // cv::norm(Point2f) computes in double
    const float fx = float(i1x - i0x), fy = float(i1y - i0y);
    const float length = (float)std::sqrt((double)fx * fx + (double)fy * fy);
    sad *= 1 + length / searchDistance();
    return sad;
  }
  void adjustInitialFlow(const ImgF& I0, const ImgF& I1, const ImgF& a0, const ImgF& a1, ImgF& flow, int hint) const {
    float sumL = 0, sumR = 0;
    for (int y = 0; y < I0.h; ++y)
      for (int x = 0; x < I0.w; ++x) {
        float alpha = a0.p(y)[x] * a1.p(y)[x];
        sumL += alpha * I0.p(y)[x];
        sumR += alpha * I1.p(y)[x];
      }
    const float ratio = sumL / sumR;
    ImgF I1eq = I1;
    scaleF(I1eq, ratio);
    const int dist = searchDistance();
    const int ortho = (dist + 4) / 8, thick = 2 * ortho + 1;
    int bx, by, bw, bh;
    switch (hint) {
      case HINT_RIGHT: bx = 0; by = -ortho; bw = dist + 1; bh = thick; break;
      case HINT_DOWN: bx = -ortho; by = 0; bw = thick; bh = dist + 1; break;
      case HINT_LEFT: bx = -dist; by = -ortho; bw = dist + 1; bh = thick; break;
      case HINT_UP: bx = -ortho; by = -dist; bw = thick; bh = dist + 1; break;
      default: throw std::runtime_error("unexpected direction");
    }
    for (int i0y = 0; i0y < I0.h; ++i0y)
      for (int i0x = 0; i0x < I0.w; ++i0x)
        if (a0.p(i0y)[i0x] > kUpdateAlphaThreshold) {
          float errorBest = 0.8f * patchError(I0, a0, i0x, i0y, I1eq, a1, i0x, i0y);
          int bxBest = i0x, byBest = i0y;
          for (int dy = by; dy < by + bh; ++dy)
            for (int dx = bx; dx < bx + bw; ++dx) {
              int i1x = i0x + dx, i1y = i0y + dy;
              if (0 <= i1x && i1x < I1.w && 0 <= i1y && i1y < I1.h) {
                float e = patchError(I0, a0, i0x, i0y, I1eq, a1, i1x, i1y);
                if (errorBest > e) { errorBest = e; bxBest = i1x; byBest = i1y; }
              }
            }
          flow.p(i0y, i0x)[0] = float(bxBest - i0x);
          flow.p(i0y, i0x)[1] = float(byBest - i0y);
        }
  }

  // PixFlow.h:344-413
  void patchMatchPropagationAndSearch(const ImgF& I0, const ImgF& I1, const ImgF& a0, const ImgF& a1, ImgF& flow,
                                      int hint) const {
    ImgF I0x, I0y, I1x, I1y;
    sobel1(I0, I0x, true); sobel1(I0, I0y, false); sobel1(I1, I1x, true); sobel1(I1, I1y, false);
    gaussianBlurF(I0x, I0x, 3, 0.5); gaussianBlurF(I0y, I0y, 3, 0.5);
    gaussianBlurF(I1x, I1x, 3, 0.5); gaussianBlurF(I1y, I1y, 3, 0.5);
    if (flow.empty()) {
      flow = ImgF(I0.w, I0.h, 2, 0.f);
      if (P.maxPercentage > 0 && hint != HINT_UNKNOWN) adjustInitialFlow(I0, I1, a0, a1, flow, hint);
    }
    ImgF blurred;
    gaussianBlurF(flow, blurred, 15, 8.0);
    Level L{&a0, &a1, &I0x, &I0y, &I1x, &I1y, &blurred, &flow, I0.w, I0.h};
    for (int y = 0; y < L.rows; ++y)
      for (int x = 0; x < L.cols; ++x) sweepPixel(L, x, y, +1);
    median5(flow, flow);
    for (int y = L.rows - 1; y >= 0; --y)
      for (int x = L.cols - 1; x >= 0; --x) sweepPixel(L, x, y, -1);
    median5(flow, flow);
    lowAlphaFlowDiffusion(a0, a1, flow);
  }

  // BGRA8 -> (grey/255 blurred, alpha/255) at half size. PixFlow.h:96-99,121-138
  void prepare(const Img8& rgba, int dw, int dh, ImgF& I, ImgF& A, Img8* smallOut = nullptr) const {
    Img8 small;
    resizeCubic8(rgba, small, dw, dh);
    I = ImgF(dw, dh, 1);
    A = ImgF(dw, dh, 1);
    const float inv255 = (float)(1. / 255.0f);
    for (int i = 0; i < dw * dh; ++i) {
      const uint8_t* px = &small.d[(size_t)i * 4];
      I.d[i] = float(gray8(px[0], px[1], px[2], legacyGray)) * inv255;
      A.d[i] = float(px[3]) * inv255;
    }
    gaussianBlurF(I, I, 5, 0.25);
    if (smallOut) *smallOut = std::move(small);
  }

  // PixFlow.h:81-183. prev* may be empty (no temporal regularisation).
  void computeOpticalFlow(const Img8& rgba0, const Img8& rgba1, const ImgF& prevFlow, const Img8& prevI0,
                          const Img8& prevI1, ImgF& flow, int hint) const {
    const int ow = rgba0.w, oh = rgba0.h;
    const int dw = (int)(ow * P.downscaleFactor), dh = (int)(oh * P.downscaleFactor);
    ImgF I0, I1, A0, A1;
    Img8 small0, small1;
    prepare(rgba0, dw, dh, I0, A0, &small0);
    prepare(rgba1, dw, dh, I1, A1, &small1);
    const bool temporal = !prevFlow.empty();
    ImgF prevFlowDs, motion;
    if (temporal) {  // PixFlow.h:101-118
      resizeCubicF(prevFlow, prevFlowDs, dw, dh);
      scaleF(prevFlowDs, float(dh) / float(prevFlow.h));
      Img8 p1s;
      resizeCubic8(prevI1, p1s, dw, dh);  // prevI0 is resized by the reference but never read
      (void)prevI0;
      motion = ImgF(dw, dh, 1);
      for (int i = 0; i < dw * dh; ++i) {
        const uint8_t* a = &small1.d[(size_t)i * 4];
        const uint8_t* b = &p1s.d[(size_t)i * 4];
        motion.d[i] = (float)((fabs(a[0] - b[0]) + fabs(a[1] - b[1]) + fabs(a[2] - b[2])) / (255.0f * 3.0f));
      }
    }
    std::vector<ImgF> pI0 = buildPyramid(I0), pI1 = buildPyramid(I1), pA0 = buildPyramid(A0), pA1 = buildPyramid(A1);
    std::vector<ImgF> pPrev, pMotion;
    if (temporal) {
      pPrev = buildPyramid(prevFlowDs);
      pMotion = buildPyramid(motion);
      for (size_t l = 0; l < pPrev.size(); ++l) scaleF(pPrev[l], float(pPrev[l].h) / float(pPrev[0].h));
    }
    flow = ImgF();
    for (int level = (int)pI0.size() - 1; level >= 0; --level) {
      patchMatchPropagationAndSearch(pI0[level], pI1[level], pA0[level], pA1[level], flow, hint);
      if (temporal) {  // PixFlow.h:185-193
        for (int i = 0; i < flow.w * flow.h; ++i) {
          const float w = 1.0f - pMotion[level].d[i];
          float* f = &flow.d[(size_t)i * 2];
          const float* pf = &pPrev[level].d[(size_t)i * 2];
          f[0] = f[0] * (1.0f - w) + pf[0] * w;
          f[1] = f[1] * (1.0f - w) + pf[1] * w;
        }
      }
      if (levelDump) levelDump->push_back(flow);
      if (level > 0) {
        resizeCubicF(flow, flow, pI0[level - 1].w, pI0[level - 1].h);
        scaleF(flow, 1.0f / P.pyrScaleFactor);
      }
    }
    resizeLinearF(flow, flow, ow, oh);
    scaleF(flow, 1.0f / P.downscaleFactor);
    gaussianBlurF(flow, flow, 3, 1.0);
  }
};

}  // namespace orc
