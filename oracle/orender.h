// ORACLE — TEST INFRASTRUCTURE ONLY (see ocv.h header).
//
// Restates the per-frame stereo panorama pipeline of
//   surround360_render/source/test/TestRenderStereoPanorama.cpp:99-972   (TRSP)
//   surround360_render/source/optical_flow/NovelView.cpp:101-299
//   surround360_render/source/render/ImageWarper.cpp:26-174
//   surround360_render/source/util/CvUtil.cpp:69-260
//   surround360_render/source/util/Filter.h:39-137
// with the reference's threading shape (thread per side camera / per pair, 4 pole
// threads, 2 sharpen threads) so it can also serve as the timed CPU baseline.
#pragma once
#include <chrono>
#include <functional>
#include <thread>
#include "ocamera.h"
#include "opixflow.h"

namespace orc {

static inline float rampf(float x, float a, float b) { return std::max(0.0f, std::min(1.0f, (x - a) / (b - a))); }
static inline float lerpf(float x0, float x1, float a) { return x0 * (1.0f - a) + x1 * a; }

struct RenderParams {  // gflags of TRSP:44-70 that affect pixels
  double interpupilary_dist = 6.4;
  int side_alpha_feather_size = 100;
  int std_alpha_feather_size = 31;
  double sharpening = 0.0;
  bool enable_top = false, enable_bottom = false, enable_pole_removal = false;
  std::string poleremoval_flow_alg = "pixflow_low";
  std::string side_flow_alg = "pixflow_low", polar_flow_alg = "pixflow_low";
  double zero_parallax_dist = 10000.0;
  int eqr_width = 256, eqr_height = 128;
  int final_eqr_width = 3480, final_eqr_height = 960;
  int cubemap_width = 1536, cubemap_height = 1536;
  std::string cubemap_format = "video";
  bool want_cubemap = false;
  bool threads = true;
  bool keep_intermediates = false;
};

// ---------------- ImageWarper.cpp:143-174 ----------------
static inline void sphericalWarpMap(ImgF& warp, int w, int h, const Camera& cam, float leftAngle, float rightAngle,
                                    float topAngle, float bottomAngle) {
  warp = ImgF(w, h, 2);
  for (int x = 0; x < w; ++x) {
    const float xFrac = (x + 0.5f) / w;
    const float xAngle = (1 - xFrac) * leftAngle + xFrac * rightAngle;
    for (int y = 0; y < h; ++y) {
      const float yFrac = (y + 0.5f) / h;
      float yAngle = (1 - yFrac) * topAngle + yFrac * bottomAngle;
      const float ux = cosf(yAngle) * cosf(xAngle), uy = cosf(yAngle) * sinf(xAngle), uz = sinf(yAngle);
      const int mul = int(Camera::kNearInfinity);
      V2 px = cam.pixel(V3((double)ux * mul, (double)uy * mul, (double)uz * mul));
      warp.p(y, x)[0] = (float)(px.x - 0.5);
      warp.p(y, x)[1] = (float)(px.y - 0.5);
    }
  }
}
static inline Img8 bgr2bgra(const Img8& s) {
  Img8 d(s.w, s.h, 4);
  for (size_t i = 0; i < (size_t)s.w * s.h; ++i) {
    d.d[i * 4] = s.d[i * 3]; d.d[i * 4 + 1] = s.d[i * 3 + 1]; d.d[i * 4 + 2] = s.d[i * 3 + 2]; d.d[i * 4 + 3] = 255;
  }
  return d;
}
static inline Img8 bgra2bgr(const Img8& s) {
  Img8 d(s.w, s.h, 3);
  for (size_t i = 0; i < (size_t)s.w * s.h; ++i) {
    d.d[i * 3] = s.d[i * 4]; d.d[i * 3 + 1] = s.d[i * 4 + 1]; d.d[i * 3 + 2] = s.d[i * 4 + 2];
  }
  return d;
}
// dstChannels 3 or 4; src 3 or 4 channels
static inline Img8 bicubicRemapToSpherical(int dw, int dh, int dstChannels, const Img8& src, const Camera& cam,
                                           float l, float r, float t, float b) {
  ImgF warp;
  sphericalWarpMap(warp, dw, dh, cam, l, r, t, b);
  Img8 dst;
  if (src.c == 3 && dstChannels == 4) remapCubic8(bgr2bgra(src), warp, dst, B_CONSTANT);
  else remapCubic8(src, warp, dst, B_CONSTANT);
  return dst;
}

// ---------------- TRSP:99-135 ----------------
static inline Img8 projectSideToSpherical(int dw, int dh, const Img8& src, const Camera& cam, float l, float r,
                                          float t, float b, int feather) {
  Img8 tmp = src.c == 3 ? bgr2bgra(src) : src;
  for (int y = 0; y < feather; ++y) {
    const uint8_t alpha = (uint8_t)(255.0f * float(y + 0.5f) / float(feather));
    for (int x = 0; x < tmp.w; ++x) {
      tmp.p(y, x)[3] = alpha;
      tmp.p(tmp.h - 1 - y, x)[3] = alpha;
    }
  }
  return bicubicRemapToSpherical(dw, dh, 4, tmp, cam, l, r, t, b);
}

// ---------------- NovelView.cpp:174-224 ----------------
// warp columns: (x, t) per chunk column u; y coordinate of the warp is the row v.
struct LazyWarp {
  std::vector<float> x, t;  // per column
};
static inline void renderLazyNovelView(int width, int height, const LazyWarp& lw, const Img8& srcImage,
                                       const ImgF& flow, bool invertT, Img8& novelView, ImgF& flowMag) {
  ImgF warpFlow(width, height, 2);
  for (int y = 0; y < height; ++y)
    for (int x = 0; x < width; ++x) {
      warpFlow.p(y, x)[0] = lw.x[x];
      warpFlow.p(y, x)[1] = x < (int)lw.x.size() && lw.t[x] >= 0 ? float(y) : 0.f;
    }
  ImgF remappedFlow;
  remapCubicF(flow, warpFlow, remappedFlow);
  ImgF comp(width, height, 2);
  for (int y = 0; y < height; ++y)
    for (int x = 0; x < width; ++x) {
      const float lt = std::max(lw.t[x], 0.f);
      const float t = invertT ? (1.0f - lt) : lt;
      const float* fd = remappedFlow.p(y, x);
      comp.p(y, x)[0] = warpFlow.p(y, x)[0] + fd[0] * t;
      comp.p(y, x)[1] = warpFlow.p(y, x)[1] + fd[1] * t;
    }
  remapCubic8(srcImage, comp, novelView, B_CONSTANT);
  flowMag = ImgF(width, height, 1);
  for (int y = 0; y < height; ++y)
    for (int x = 0; x < width; ++x) {
      const float lt = std::max(lw.t[x], 0.f);
      const float t = invertT ? (1.0f - lt) : lt;
      const float* fd = remappedFlow.p(y, x);
      novelView.p(y, x)[3] = (uint8_t) int((1.0f - t) * novelView.p(y, x)[3]);
      flowMag.p(y)[x] = sqrtf(fd[0] * fd[0] + fd[1] * fd[1]);
    }
}

// ---------------- NovelView.cpp:101-154 ----------------
static inline Img8 combineLazyViews(const Img8& imageL, const Img8& imageR, const ImgF& magLm, const ImgF& magRm) {
  Img8 out(imageL.w, imageL.h, 4);
  for (int y = 0; y < imageL.h; ++y)
    for (int x = 0; x < imageL.w; ++x) {
      const uint8_t* cL = imageL.p(y, x);
      const uint8_t* cR = imageR.p(y, x);
      uint8_t* o = out.p(y, x);
      const uint8_t outAlpha = std::max(cL[3], cR[3]) / 255.0f > 0.1 ? 255 : 0;
      if (cL[3] == 0 && cR[3] == 0) { o[0] = o[1] = o[2] = 0; o[3] = outAlpha; }
      else if (cL[3] == 0) { o[0] = cR[0]; o[1] = cR[1]; o[2] = cR[2]; o[3] = outAlpha; }
      else if (cR[3] == 0) { o[0] = cL[0]; o[1] = cL[1]; o[2] = cL[2]; o[3] = outAlpha; }
      else {
        const float magL = magLm.p(y)[x] / float(imageL.w);
        const float magR = magRm.p(y)[x] / float(imageL.w);
        float blendL = float(cL[3]), blendR = float(cR[3]);
        const float norm = blendL + blendR;
        blendL /= norm;
        blendR /= norm;
        const float colorDiff =
            (std::abs(cL[0] - cR[0]) + std::abs(cL[1] - cR[1]) + std::abs(cL[2] - cR[2])) / 255.0f;
        const float deghostCoef = tanhf(colorDiff * 10.0f);
        const double expL = exp(10.0f * blendL * (1.0 + 20.0f * magL));
        const double expR = exp(10.0f * blendR * (1.0 + 20.0f * magR));
        const double sumExp = expL + expR + 0.00001;
        const float softmaxL = float(expL / sumExp), softmaxR = float(expR / sumExp);
        const float wL = lerpf(blendL, softmaxL, deghostCoef), wR = lerpf(blendR, softmaxR, deghostCoef);
        o[0] = (uint8_t)(float(cL[0]) * wL + float(cR[0]) * wR);
        o[1] = (uint8_t)(float(cL[1]) * wL + float(cR[1]) * wR);
        o[2] = (uint8_t)(float(cL[2]) * wL + float(cR[2]) * wR);
        o[3] = 255;
      }
    }
  return out;
}

// ---------------- CvUtil.cpp:224-260 ----------------
static inline Img8 flattenLayersDeghostPreferBase(const Img8& bottom, const Img8& top) {
  Img8 out(bottom.w, bottom.h, 4);
  for (size_t i = 0; i < (size_t)bottom.w * bottom.h; ++i) {
    const uint8_t* base = &bottom.d[i * 4];
    const uint8_t* tc = &top.d[i * 4];
    uint8_t* o = &out.d[i * 4];
    const float colorDiff = (std::abs(base[0] - tc[0]) + std::abs(base[1] - tc[1]) + std::abs(base[2] - tc[2])) / 255.0f;
    const float deghostCoef = tanhf(colorDiff * 5.0f);
    const float alphaR = tc[3] / 255.0f;
    const float alphaL = 1.0f - alphaR;
    // CvUtil.cpp:245-246: the argument is float and CvUtil.cpp is C++ with <math.h> + using
    // namespace std, so exp() resolves to the float overload (expf), widened to double after
    const double expL = expf(5.0f * alphaL * 2.0f);
    const double expR = expf(5.0f * alphaR);
    const double sumExp = expL + expR + 0.00001;
    const float softmaxL = float(expL / sumExp);
    const float softmaxR = 1.0f - softmaxL;
    const float wL = lerpf(alphaL, softmaxL, deghostCoef), wR = lerpf(alphaR, softmaxR, deghostCoef);
    o[0] = (uint8_t)(float(base[0]) * wL + float(tc[0]) * wR);
    o[1] = (uint8_t)(float(base[1]) * wL + float(tc[1]) * wR);
    o[2] = (uint8_t)(float(base[2]) * wL + float(tc[2]) * wR);
    o[3] = std::max(tc[3], base[3]);
  }
  return out;
}

// ---------------- CvUtil.cpp:140-157 ----------------
static inline Img8 featherAlphaChannel(const Img8& src, int erodeSize) {
  Img8 a(src.w, src.h, 1);
  for (size_t i = 0; i < (size_t)src.w * src.h; ++i) a.d[i] = src.d[i * 4 + 3];
  erodeCross8(a, a, erodeSize);
  gaussianBlur8Fixed(a, a, gaussKernelFixed8(erodeSize, erodeSize / 2.0f));
  Img8 d = src;
  for (size_t i = 0; i < (size_t)src.w * src.h; ++i) d.d[i * 4 + 3] = a.d[i];
  return d;
}

static inline Img8 flipBoth(const Img8& s) {  // cv::flip(-1)
  Img8 d(s.w, s.h, s.c);
  for (int y = 0; y < s.h; ++y)
    for (int x = 0; x < s.w; ++x) memcpy(d.p(s.h - 1 - y, s.w - 1 - x), s.p(y, x), s.c);
  return d;
}
static inline Img8 padRows(const Img8& s, int above, int below) {  // copyMakeBorder CONSTANT 0
  Img8 d(s.w, s.h + above + below, s.c, 0);
  memcpy(d.p(above), s.d.data(), s.d.size());
  return d;
}
static inline Img8 cropCols(const Img8& s, int x0, int w) {
  Img8 d(w, s.h, s.c);
  for (int y = 0; y < s.h; ++y) memcpy(d.p(y), s.p(y, x0), (size_t)w * s.c);
  return d;
}
static inline Img8 cropRows(const Img8& s, int h) {
  Img8 d(s.w, h, s.c);
  memcpy(d.d.data(), s.d.data(), d.d.size());
  return d;
}

// ---------------- CvUtil.cpp:93-115 ----------------
static inline Img8 offsetHorizontalWrap(const Img8& src, float offset) {
  Img8 d(src.w, src.h, src.c);
  std::vector<int> idx(src.w);
  for (int x = 0; x < src.w; ++x) idx[x] = offsetWrapIndex(x, src.w, offset);
  for (int y = 0; y < src.h; ++y)
    for (int x = 0; x < src.w; ++x) memcpy(d.p(y, x), src.p(y, idx[x]), src.c);
  return d;
}

// ---------------- Filter.h:39-137 (P = Vec3b, wrap horizontally, reflect vertically) ----
static inline int wrapi(int x, int r) { return x < 0 ? r + x : x >= r ? x - r : x; }
static inline int reflecti(int x, int r) { return x < 0 ? -x : x >= r ? 2 * r - x - 2 : x; }
static inline void iirLowPass(const Img8& in, float amount, Img8& lp) {
  const float alpha = powf(amount, 1.0f / 4.0f);
  const int width = in.w, height = in.h;
  lp = Img8(width, height, 3);
  std::vector<float> buffer((size_t)std::max(width, height) * 3);
  auto clampf = [](float v) { return v < 0.0f ? 0.0f : v > 255.0f ? 255.0f : v; };
  for (int i = 0; i < height; ++i) {
    float v[3] = {float(in.p(i, 0)[0]), float(in.p(i, 0)[1]), float(in.p(i, 0)[2])};
    for (int j = 1; j <= width; ++j) {
      const uint8_t* ip = in.p(i, wrapi(j, width));
      float* b = &buffer[(size_t)wrapi(j - 1, width) * 3];
      for (int c = 0; c < 3; ++c) { v[c] = lerpf(float(ip[c]), v[c], alpha); b[c] = v[c]; }
    }
    for (int j = width - 2; j >= -1; --j) {
      const float* ip = &buffer[(size_t)wrapi(j, width) * 3];
      for (int c = 0; c < 3; ++c) { v[c] = lerpf(ip[c], v[c], alpha); lp.p(i, j + 1)[c] = (uint8_t)clampf(v[c]); }
    }
  }
  for (int j = 0; j < width; ++j) {
    float v[3] = {float(lp.p(0, j)[0]), float(lp.p(0, j)[1]), float(lp.p(0, j)[2])};
    for (int i = 1; i <= height; ++i) {
      const uint8_t* ip = lp.p(reflecti(i, height), j);
      float* b = &buffer[(size_t)reflecti(i - 1, height) * 3];
      for (int c = 0; c < 3; ++c) { v[c] = lerpf(float(ip[c]), v[c], alpha); b[c] = v[c]; }
    }
    for (int i = height - 2; i >= -1; --i) {
      const float* ip = &buffer[(size_t)reflecti(i, height) * 3];
      for (int c = 0; c < 3; ++c) { v[c] = lerpf(ip[c], v[c], alpha); lp.p(i + 1, j)[c] = (uint8_t)clampf(v[c]); }
    }
  }
}
static inline void sharpenWithIirLowPass(Img8& img, const Img8& lp, float amount) {
  auto clampf = [](float v) { return v < 0.0f ? 0.0f : v > 255.0f ? 255.0f : v; };
  for (size_t i = 0; i < img.d.size(); ++i) {
    const float l = float(lp.d[i]);
    const float hp = float(img.d[i]) - l;
    const float ng = 1.0f - expf(-((hp * hp) * 100.0f));
    img.d[i] = (uint8_t)clampf(l + hp * ng * amount);
  }
}
static inline void sharpenEye(Img8& img, double sharpening) {  // TRSP:688-696
  Img8 lp;
  iirLowPass(img, 0.25f, lp);
  sharpenWithIirLowPass(img, lp, (float)(1.0f + sharpening));
}

// ---------------- ImageWarper.cpp:26-141 ----------------
enum CubeFace { FACE_BACK, FACE_LEFT, FACE_TOP, FACE_BOTTOM, FACE_FRONT, FACE_RIGHT };
static inline void cubemapCoord(float x, float y, int face, int srcW, int srcH, float fov, float& srcX, float& srcY) {
  float d[3] = {x, y, 0.5f}, o[3] = {x, y, 0.5f};
  switch (face) {
    case FACE_BACK: o[0] = d[0]; o[1] = d[2]; o[2] = -d[1]; break;
    case FACE_LEFT: o[0] = -d[2]; o[1] = d[0]; o[2] = -d[1]; break;
    case FACE_TOP: break;
    case FACE_BOTTOM: o[0] = d[0]; o[1] = -d[1]; o[2] = -d[2]; break;
    case FACE_FRONT: o[0] = -d[0]; o[1] = -d[2]; o[2] = -d[1]; break;
    case FACE_RIGHT: o[0] = d[2]; o[1] = -d[0]; o[2] = -d[1]; break;
  }
  const float r = sqrtf(o[0] * o[0] + o[1] * o[1]);
  const double nrm = std::sqrt((double)o[0] * o[0] + (double)o[1] * o[1] + (double)o[2] * o[2]);
  const float phi = acosf((float)(o[2] / nrm));
  float theta = r > 0.0f ? acosf(fabsf(o[0] / r)) : 0.0f;
  if (o[0] > 0 && o[1] > 0) {}
  else if (o[0] <= 0 && o[1] > 0) theta = (float)(M_PI - theta);
  else if (o[0] <= 0 && o[1] <= 0) theta = (float)(M_PI + theta);
  else theta = (float)(2 * M_PI - theta);
  const float phiP = phi < 0.0f ? 0.0f : phi > fov ? fov : phi;
  const float twoPi = float(2.0f * M_PI);
  const float thetaP = theta < 0.0f ? 0.0f : theta > twoPi ? twoPi : theta;
  srcX = (float)(float(srcW) * thetaP / (2.0f * M_PI));
  srcY = float(srcH) * phiP / fov;
}
static inline std::vector<Img8> sphericalToCubemap(const Img8& src, float fov, int fw, int fh) {
  static const int faces[6] = {FACE_RIGHT, FACE_LEFT, FACE_TOP, FACE_BOTTOM, FACE_BACK, FACE_FRONT};
  const float dy = 1.0f / float(fw), dx = 1.0f / float(fh);
  std::vector<Img8> out;
  for (int f = 0; f < 6; ++f) {
    ImgF warp(fw, fh, 2);
    for (int j = 0; j < fh; ++j)
      for (int i = 0; i < fw; ++i)
        cubemapCoord(float(i) * dy - 0.5f, float(j) * dx - 0.5f, faces[f], src.w, src.h, fov, warp.p(j, i)[0],
                     warp.p(j, i)[1]);
    Img8 face;
    remapCubic8(src, warp, face, B_WRAP);
    out.push_back(std::move(face));
  }
  return out;
}
static inline Img8 flipH(const Img8& s) {
  Img8 d(s.w, s.h, s.c);
  for (int y = 0; y < s.h; ++y)
    for (int x = 0; x < s.w; ++x) memcpy(d.p(y, s.w - 1 - x), s.p(y, x), s.c);
  return d;
}
static inline Img8 hstack(const std::vector<Img8>& v) {
  int w = 0;
  for (auto& m : v) w += m.w;
  Img8 d(w, v[0].h, v[0].c);
  int x0 = 0;
  for (auto& m : v) {
    for (int y = 0; y < m.h; ++y) memcpy(d.p(y, x0), m.p(y), (size_t)m.w * m.c);
    x0 += m.w;
  }
  return d;
}
static inline Img8 vstack(const std::vector<Img8>& v) {
  int h = 0;
  for (auto& m : v) h += m.h;
  Img8 d(v[0].w, h, v[0].c);
  int y0 = 0;
  for (auto& m : v) {
    memcpy(d.p(y0), m.d.data(), m.d.size());
    y0 += m.h;
  }
  return d;
}
static inline Img8 stackCubemapFaces(const std::string& fmt, const std::vector<Img8>& im) {  // CvUtil.cpp:117-138
  if (fmt == "video")
    return vstack({hstack({flipH(im[1]), flipH(im[0]), flipH(im[2])}), hstack({flipH(im[3]), flipH(im[4]), flipH(im[5])})});
  if (fmt == "photo") return vstack(im);
  throw std::runtime_error("unexpected cubemap format: " + fmt + ". valid formats are: video,photo");
}

// ---------------- frame-level state ----------------
struct SideGeometry {  // TRSP:155-162, 309-316, 340-348
  float hRadians, vRadians;
  int camW, camH, overlapW, numNovel;
  float fovHRad, slabDisp, shiftPx;
};
static inline SideGeometry sideGeometry(const Rig& rig, const RenderParams& P) {
  SideGeometry g;
  g.hRadians = 2 * approximateFov(rig.side, false);
  g.vRadians = 2 * approximateFov(rig.side, true);
  g.camH = (int)(P.eqr_height * g.vRadians / M_PI);
  g.camW = (int)(P.eqr_width * g.hRadians / (2 * M_PI));
  const int numCams = (int)rig.side.size();
  const float camFovDeg = (float)(2 * approximateFov(rig.side, false) * (180 / M_PI));
  g.fovHRad = (float)(camFovDeg * M_PI / 180.0f);
  const float overlapDeg = (float)((camFovDeg * float(numCams) - 360.0) / float(numCams));
  g.overlapW = (int)(float(g.camW) * (overlapDeg / camFovDeg));
  g.numNovel = g.camW - g.overlapW;
  const float v = atanf((float)(P.zero_parallax_dist / (P.interpupilary_dist / 2.0f)));
  const float psi = asinf((float)(sinf(v) * (P.interpupilary_dist / 2.0f) / rig.ringRadius()));
  g.slabDisp = psi * (float(g.camW) / g.fovHRad);
  const float theta = (float)(-M_PI / 2.0f + v + psi);
  g.shiftPx = (float)(float(P.eqr_width) * (theta / (2.0f * M_PI)));
  return g;
}

struct FrameResult {
  Img8 equirect;  // BGR, left eye over right eye
  Img8 cubemap;   // BGR, optional
  // intermediates (keep_intermediates)
  std::vector<Img8> projections;
  std::vector<ImgF> flowLtoR, flowRtoL;
  Img8 panoL, panoR;  // after offsetHorizontalWrap, before padding
  Img8 topSpherical, bottomSpherical, bottomMerged;
  ImgF poleFlow[4];   // top_left, top_right, bottom_left, bottom_right
  Img8 poleWarped[4];
  Img8 eyeL, eyeR;    // BGR before sharpening
  double tProject = 0, tFlow = 0, tNovel = 0, tPole = 0, tSharpen = 0, tTotal = 0;
};

// What the next frame's temporal regularisation reads back (flow .bin files and overlap / extended PNGs in the
// reference: TRSP:200-255, 413-452; lossless, so keeping them in memory is equivalent)
struct FrameState {
  bool valid = false;
  std::vector<Img8> ovL, ovR;
  std::vector<ImgF> fLR, fRL;
  Img8 extSide[4], extFish[4];
  ImgF poleFlow[4];
  bool poleValid[4] = {false, false, false, false};
};

static inline double nowSec() {
  return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}
static inline void runAll(std::vector<std::function<void()>>& jobs, bool threads) {
  if (!threads) { for (auto& j : jobs) j(); return; }
  std::vector<std::thread> th;
  for (auto& j : jobs) th.emplace_back(j);
  for (auto& t : th) t.join();
}

// TRSP:388-561
static inline void poleToSideFlow(const Rig& rig, const RenderParams& P, const Img8& sideSpherical,
                                  const Img8& fisheyeSpherical, Img8& warpedOut, ImgF* flowOut, FrameState* state = nullptr,
                                  int slot = 0) {
  const int cols = fisheyeSpherical.w, rows = fisheyeSpherical.h;
  Img8 cropped = featherAlphaChannel(cropRows(sideSpherical, rows), P.std_alpha_feather_size);
  const float kExtendFrac = 1.2f;
  const int extW = (int)(float(cols) * kExtendFrac);
  Img8 extSide(extW, rows, 4), extFish(extW, rows, 4);
  for (int y = 0; y < rows; ++y)
    for (int x = 0; x < extW; ++x) {
      memcpy(extSide.p(y, x), cropped.p(y, x % cols), 4);
      memcpy(extFish.p(y, x), fisheyeSpherical.p(y, x % cols), 4);
    }
  ImgF flow;
  PixFlow pf(flowParamsByName(P.polar_flow_alg));
  if (state && state->valid && state->poleValid[slot])  // TRSP:418-447
    pf.computeOpticalFlow(extSide, extFish, state->poleFlow[slot], state->extSide[slot], state->extFish[slot], flow, HINT_DOWN);
  else
    pf.computeOpticalFlow(extSide, extFish, ImgF(), Img8(), Img8(), flow, HINT_DOWN);
  if (state) { state->extSide[slot] = extSide; state->extFish[slot] = extFish; state->poleFlow[slot] = flow; state->poleValid[slot] = true; }
  if (flowOut) *flowOut = flow;

  float poleCameraRadius = (float)rig.findByDirection(V3(0, 0, -1)).getFov();
  float sideCameraRadius = approximateFov(rig.side, true);
  float poleCameraCropRadius =
      (float)(0.5f * (M_PI / 2 - sideCameraRadius) + 0.5f * (std::min(float(M_PI / 2), poleCameraRadius)));
  poleCameraCropRadius = (float)(poleCameraCropRadius * (180 / M_PI));
  poleCameraRadius = (float)(poleCameraRadius * (180 / M_PI));
  sideCameraRadius = (float)(sideCameraRadius * (180 / M_PI));
  const float phiFromPole = poleCameraCropRadius;
  const float phiFromSide = 90.0f - sideCameraRadius;
  const float phiMid = (phiFromPole + phiFromSide) / 2.0f;
  const float phiDiff = fabsf(phiFromPole - phiFromSide);
  const float phiRampStart = phiMid - 1.0f * phiDiff / 2.0f;
  const float phiRampEnd = phiMid + 1.0f * phiDiff / 2.0f;

  ImgF warp(extW, rows, 2);
  for (int y = 0; y < rows; ++y) {
    const float phi = poleCameraRadius * float(y + 0.5f) / float(rows);
    const float alpha = 1.0f - rampf(phi, phiRampStart, phiMid);
    for (int x = 0; x < extW; ++x) {
      warp.p(y, x)[0] = float(x) + (1.0f - alpha) * flow.p(y, x)[0];
      warp.p(y, x)[1] = float(y) + (1.0f - alpha) * flow.p(y, x)[1];
    }
  }
  Img8 warpedExt;
  remapCubic8(extFish, warp, warpedExt, B_CONSTANT);
  Img8 warped = cropCols(warpedExt, 0, cols);
  const int maxBlendX = (int)(float(cols) * (kExtendFrac - 1.0f));
  for (int y = 0; y < rows; ++y)
    for (int x = 0; x < maxBlendX; ++x) {
      uint8_t* s = warped.p(y, x);
      const uint8_t* wp = warpedExt.p(y, x + cols);
      const float alpha = 1.0f - rampf(float(x), float(maxBlendX) * 0.333f, float(maxBlendX) * 0.667f);
      for (int c = 0; c < 3; ++c) s[c] = (uint8_t)(float(wp[c]) * alpha + float(s[c]) * (1.0f - alpha));
    }
  for (int y = 0; y < rows; ++y) {
    const float phi = poleCameraRadius * float(y + 0.5f) / float(rows);
    const float alpha = 1.0f - rampf(phi, phiMid, phiRampEnd);
    for (int x = 0; x < cols; ++x) warped.p(y, x)[3] = (uint8_t)(warped.p(y, x)[3] * alpha);
  }
  warpedOut = padRows(warped, 0, sideSpherical.h - rows);
}

// ---------------- CvUtil.cpp:201-222 ----------------
static inline void circleAlphaCut(Img8& bgra, float radius) {
  for (int y = 0; y < bgra.h; ++y)
    for (int x = 0; x < bgra.w; ++x) {
      const float dx = float(x) - float(bgra.w) / 2.0f, dy = float(y) - float(bgra.h) / 2.0f;
      const float r = sqrtf(dx * dx + dy * dy);
      const float alpha = r < radius ? 1.0f : 0.0f;
      bgra.p(y, x)[3] = (unsigned char)(alpha * 255.0f);
    }
}
static inline void cutRedMaskOutOfAlphaChannel(Img8& bgra, const Img8& maskBGR) {
  for (int y = 0; y < maskBGR.h; ++y)
    for (int x = 0; x < maskBGR.w; ++x) {
      const uint8_t* m = maskBGR.p(y, x);
      if (m[0] == 0 && m[1] == 0 && m[2] == 255) bgra.p(y, x)[3] = 0;
    }
}
// ---------------- PoleRemoval.cpp:32-188 (prev_frame_data_dir NONE) ----------------
static inline Img8 combineBottomImagesWithPoleRemoval(const Img8& bottomBGR, const Img8& bottom2BGR, const Img8& mask, const Img8& mask2,
                                                      float radius, float radius2, bool flip180, const std::string& flowAlg,
                                                      int featherSize, ImgF* flowOut = nullptr) {
  Img8 b1 = bgr2bgra(bottomBGR), b2 = bgr2bgra(bottom2BGR);
  circleAlphaCut(b1, radius);
  circleAlphaCut(b2, radius2);
  cutRedMaskOutOfAlphaChannel(b1, mask);
  cutRedMaskOutOfAlphaChannel(b2, mask2);
  b1 = featherAlphaChannel(b1, featherSize);
  b2 = featherAlphaChannel(b2, featherSize);
  if (flip180) b2 = flipBoth(b2);
  ImgF flow;
  PixFlow pf(flowParamsByName(flowAlg));
  pf.computeOpticalFlow(b1, b2, ImgF(), Img8(), Img8(), flow, HINT_DOWN);
  if (flowOut) *flowOut = flow;
  ImgF warp(b1.w, b1.h, 2);
  for (int y = 0; y < b1.h; ++y)
    for (int x = 0; x < b1.w; ++x) {
      warp.p(y, x)[0] = float(x) + flow.p(y, x)[0];
      warp.p(y, x)[1] = float(y) + flow.p(y, x)[1];
    }
  Img8 warped2;
  remapCubic8(b2, warp, warped2, B_CONSTANT);
  for (int y = 0; y < b1.h; ++y)
    for (int x = 0; x < b1.w; ++x) {
      uint8_t* p1 = b1.p(y, x);
      const uint8_t* p2 = warped2.p(y, x);
      const float alpha = p1[3] / 255.0f, alpha2 = p2[3] / 255.0f;
      if (alpha < 1.0f && alpha2 > 0.0f) {
        const float a1 = alpha, a2 = 1.0f - alpha;
        const float bb = a1 * float(p1[0]) + a2 * float(p2[0]), gg = a1 * float(p1[1]) + a2 * float(p2[1]),
                    rr = a1 * float(p1[2]) + a2 * float(p2[2]);
        p1[0] = (uint8_t)bb; p1[1] = (uint8_t)gg; p1[2] = (uint8_t)rr; p1[3] = 255;
      }
    }
  circleAlphaCut(b1, radius);
  return featherAlphaChannel(b1, featherSize);
}

// TRSP:647-685 (top) / 564-644 (bottom, no pole removal)
static inline Img8 preparePoleSpherical(const Rig& rig, const RenderParams& P, const Img8& img, bool top) {
  const Camera& cam = rig.findByDirection(V3(0, 0, top ? 1 : -1));
  const int h = (int)(P.eqr_height * cam.getFov() / M_PI), w = P.eqr_width;
  // dst is created CV_8UC3, but cv::remap re-creates it with the source's type: a 4-channel source
  // (pole removal) yields a 4-channel spherical image with the remapped alpha (TRSP:607-623)
  Img8 sphAny;
  if (top)
    sphAny = bicubicRemapToSpherical(w, h, 3, img, cam, (float)(2.0f * M_PI), 0.f, (float)(M_PI / 2.0f),
                                     (float)(M_PI / 2.0f - cam.getFov()));
  else
    sphAny = bicubicRemapToSpherical(w, h, 3, img, cam, 0.f, (float)(2.0f * M_PI), (float)(-(M_PI / 2.0f)),
                                     (float)(-(M_PI / 2.0f - cam.getFov())));
  Img8 sph = sphAny.c == 4 ? sphAny : bgr2bgra(sphAny);
  const int yStart = h - 1 - P.std_alpha_feather_size;
  for (int y = std::max(yStart, 0); y < h; ++y) {
    const float alpha = 1.0f - float(y - yStart) / float(P.std_alpha_feather_size);
    const uint8_t a = (uint8_t)(255.0f * alpha);
    for (int x = 0; x < w; ++x) sph.p(y, x)[3] = top ? a : std::min(sph.p(y, x)[3], a);
  }
  return sph;
}

// images: rig.rig.size() BGR frames in rig order (index = position in the rig JSON)
static inline FrameResult renderFrame(const Rig& rig, const std::vector<Img8>& images, const RenderParams& P,
                                      const Img8* poleMask = nullptr, const Img8* poleMask2 = nullptr, FrameState* state = nullptr) {
  FrameResult R;
  const double t0 = nowSec();
  const int numCams = (int)rig.side.size();
  if (P.eqr_width % numCams != 0) throw std::runtime_error("eqr_width must be evenly divisible by the number of cameras");
  auto imageOf = [&](const std::string& id) -> const Img8& {
    for (size_t i = 0; i < rig.rig.size(); ++i)
      if (rig.rig[i].id == id) return images[i];
    throw std::runtime_error("no image for camera " + id);
  };
  const SideGeometry g = sideGeometry(rig, P);

  Img8 topSph, botSph;
  std::thread topThread, botThread;
  auto topJob = [&] { topSph = preparePoleSpherical(rig, P, imageOf(rig.findByDirection(V3(0, 0, 1)).id), true); };
  auto botJob = [&] {
    const Camera& cam = rig.findByDirection(V3(0, 0, -1));
    if (P.enable_pole_removal) {  // TRSP:569-597
      if (!poleMask || !poleMask2) throw std::runtime_error("missing or bad pole mask");
      const Camera& cam2 = rig.findLargestAxisDist();
      const bool flip180 = cam.up().dot(cam2.up()) < 0;
      Img8 merged = combineBottomImagesWithPoleRemoval(imageOf(cam.id), imageOf(cam2.id), *poleMask, *poleMask2,
                                                       Camera::approximateUsablePixelsRadius(cam), Camera::approximateUsablePixelsRadius(cam2),
                                                       flip180, P.poleremoval_flow_alg, P.std_alpha_feather_size);
      if (P.keep_intermediates) R.bottomMerged = merged;
      botSph = preparePoleSpherical(rig, P, merged, false);
    } else {
      botSph = preparePoleSpherical(rig, P, imageOf(cam.id), false);
    }
  };
  if (P.enable_bottom) { if (P.threads) botThread = std::thread(botJob); else botJob(); }
  if (P.enable_top) { if (P.threads) topThread = std::thread(topJob); else topJob(); }

  // --- side projections TRSP:138-175
  std::vector<Img8> proj(numCams);
  {
    std::vector<std::function<void()>> jobs;
    for (int i = 0; i < numCams; ++i)
      jobs.push_back([&, i] {
        const float direction = (float)(-float(i) / float(numCams) * 2.0f * M_PI);
        proj[i] = projectSideToSpherical(g.camW, g.camH, imageOf(rig.side[i].id), rig.side[i],
                                         direction + g.hRadians / 2, direction - g.hRadians / 2, g.vRadians / 2,
                                         -g.vRadians / 2, P.side_alpha_feather_size);
      });
    runAll(jobs, P.threads);
  }
  const double t1 = nowSec();
  R.tProject = t1 - t0;

  // --- side flow TRSP:318-337, NovelView.cpp:270-299
  std::vector<Img8> ovL(numCams), ovR(numCams);
  std::vector<ImgF> fLR(numCams), fRL(numCams);
  {
    std::vector<std::function<void()>> jobs;
    for (int i = 0; i < numCams; ++i)
      jobs.push_back([&, i] {
        const int r = (i + 1) % numCams;
        ovL[i] = cropCols(proj[i], proj[i].w - g.overlapW, g.overlapW);
        ovR[i] = cropCols(proj[r], 0, g.overlapW);
        PixFlow pf(flowParamsByName(P.side_flow_alg));
        if (state && state->valid) {  // TRSP:210-244, NovelView.cpp:282-297
          pf.computeOpticalFlow(ovL[i], ovR[i], state->fLR[i], state->ovL[i], state->ovR[i], fLR[i], HINT_LEFT);
          pf.computeOpticalFlow(ovR[i], ovL[i], state->fRL[i], state->ovR[i], state->ovL[i], fRL[i], HINT_RIGHT);
        } else {
          pf.computeOpticalFlow(ovL[i], ovR[i], ImgF(), Img8(), Img8(), fLR[i], HINT_LEFT);
          pf.computeOpticalFlow(ovR[i], ovL[i], ImgF(), Img8(), Img8(), fRL[i], HINT_RIGHT);
        }
      });
    runAll(jobs, P.threads);
  }
  const double t2 = nowSec();
  R.tFlow = t2 - t1;

  // --- novel view chunks TRSP:259-292, 350-384
  std::vector<Img8> chunkL(numCams), chunkR(numCams);
  {
    const int bufW = P.eqr_width / numCams;
    LazyWarp wl, wr;
    wl.x.assign(bufW, 0.f); wl.t.assign(bufW, -1.f);  // t<0 marks unfilled columns ((0,0,0) in the reference)
    wr = wl;
    for (int nv = 0; nv < g.numNovel && nv < bufW; ++nv) {
      const float shift = float(nv) / float(g.numNovel);
      const float slabShift = float(g.camW) * 0.5f - float(g.numNovel - nv);
      wl.x[nv] = slabShift + g.slabDisp; wl.t[nv] = shift;
      wr.x[nv] = slabShift - g.slabDisp; wr.t[nv] = shift;
    }
    std::vector<std::function<void()>> jobs;
    for (int i = 0; i < numCams; ++i)
      jobs.push_back([&, i] {
        Img8 a, b, c, d;
        ImgF ma, mb, mc, md;
        renderLazyNovelView(bufW, g.camH, wl, ovL[i], fRL[i], false, a, ma);
        renderLazyNovelView(bufW, g.camH, wl, ovR[i], fLR[i], true, b, mb);
        renderLazyNovelView(bufW, g.camH, wr, ovL[i], fRL[i], false, c, mc);
        renderLazyNovelView(bufW, g.camH, wr, ovR[i], fLR[i], true, d, md);
        chunkL[i] = combineLazyViews(a, b, ma, mb);
        chunkR[i] = combineLazyViews(c, d, mc, md);
      });
    runAll(jobs, P.threads);
  }
  Img8 sphL = offsetHorizontalWrap(hstack(chunkL), g.shiftPx);
  Img8 sphR = offsetHorizontalWrap(hstack(chunkR), -g.shiftPx);
  const double t3 = nowSec();
  R.tNovel = t3 - t2;
  if (P.keep_intermediates) {
    R.projections = proj; R.flowLtoR = fLR; R.flowRtoL = fRL; R.panoL = sphL; R.panoR = sphR;
  }

  // --- pad TRSP:701-713, 806-807
  auto padTo = [&](Img8& m) {
    const int above = (P.eqr_height - m.h) / 2, below = P.eqr_height - m.h - above;
    m = padRows(m, above, below);
  };
  padTo(sphL);
  padTo(sphR);

  // --- poles TRSP:811-885
  Img8 topWL, topWR, botWL, botWR, flipL, flipR;
  std::vector<std::function<void()>> poleJobs;
  if (P.enable_top) {
    if (topThread.joinable()) topThread.join();
    poleJobs.push_back([&] { poleToSideFlow(rig, P, sphL, topSph, topWL, P.keep_intermediates ? &R.poleFlow[0] : nullptr, state, 0); });
    poleJobs.push_back([&] { poleToSideFlow(rig, P, sphR, topSph, topWR, P.keep_intermediates ? &R.poleFlow[1] : nullptr, state, 1); });
  }
  if (P.enable_bottom) {
    if (botThread.joinable()) botThread.join();
    flipL = flipBoth(sphL);
    flipR = flipBoth(sphR);
    poleJobs.push_back([&] { poleToSideFlow(rig, P, flipL, botSph, botWL, P.keep_intermediates ? &R.poleFlow[2] : nullptr, state, 2); });
    poleJobs.push_back([&] { poleToSideFlow(rig, P, flipR, botSph, botWR, P.keep_intermediates ? &R.poleFlow[3] : nullptr, state, 3); });
  }
  runAll(poleJobs, P.threads);
  if (state) { state->ovL = ovL; state->ovR = ovR; state->fLR = fLR; state->fRL = fRL; state->valid = true; }
  if (P.enable_top) {
    sphL = flattenLayersDeghostPreferBase(sphL, topWL);
    sphR = flattenLayersDeghostPreferBase(sphR, topWR);
  }
  if (P.enable_bottom) {
    sphL = flipBoth(flattenLayersDeghostPreferBase(flipBoth(sphL), botWL));
    sphR = flipBoth(flattenLayersDeghostPreferBase(flipBoth(sphR), botWR));
  }
  const double t4 = nowSec();
  R.tPole = t4 - t3;
  if (P.keep_intermediates) {
    R.topSpherical = topSph; R.bottomSpherical = botSph;
    R.poleWarped[0] = topWL; R.poleWarped[1] = topWR; R.poleWarped[2] = botWL; R.poleWarped[3] = botWR;
  }

  Img8 eyeL = bgra2bgr(sphL), eyeR = bgra2bgr(sphR);
  if (P.keep_intermediates) { R.eyeL = eyeL; R.eyeR = eyeR; }
  if (P.sharpening > 0.0f) {
    std::vector<std::function<void()>> jobs = {[&] { sharpenEye(eyeL, P.sharpening); }, [&] { sharpenEye(eyeR, P.sharpening); }};
    runAll(jobs, P.threads);
  }
  const double t5 = nowSec();
  R.tSharpen = t5 - t4;

  if (P.want_cubemap && P.cubemap_width > 0 && P.cubemap_height > 0) {
    Img8 cl = stackCubemapFaces(P.cubemap_format, sphericalToCubemap(eyeL, (float)M_PI, P.cubemap_width, P.cubemap_height));
    Img8 cr = stackCubemapFaces(P.cubemap_format, sphericalToCubemap(eyeR, (float)M_PI, P.cubemap_width, P.cubemap_height));
    R.cubemap = vstack({cl, cr});
  }
  if (P.final_eqr_width != 0 && P.final_eqr_height != 0 && P.final_eqr_width != P.eqr_width &&
      P.final_eqr_height != P.eqr_height / 2) {
    resizeCubic8(eyeL, eyeL, P.final_eqr_width, P.final_eqr_height / 2);
    resizeCubic8(eyeR, eyeR, P.final_eqr_width, P.final_eqr_height / 2);
  }
  R.equirect = vstack({eyeL, eyeR});
  R.tTotal = nowSec() - t0;
  return R;
}

}  // namespace orc
