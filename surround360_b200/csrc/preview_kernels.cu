// TestHyperPreview on sm_100a: the reference's fast mono preview (test/TestHyperPreview.cpp:77-185) straight from raw
// Bayer frames of the top and the two bottom cameras:
//   simpleDemosaic (2x2 binning, :161-185) -> BGR2BGRA (alpha 1) -> topDownAlphaFade (bottom cameras) and radialAlphaFade
//   (util/CvUtil.cpp:310-334) -> cubic remap through a PRECOMPUTED equirect->camera warp (:114-139, projectEquirectToCam
//   render/ImageWarper.cpp:181-199) -> flattenLayersAlphaSoftmax (util/CvUtil.cpp:336-360) -> * 255/65535.
// The warp tables are frame-invariant: the host builds them once per (rig, equirect size) with the reference's arithmetic (float
// sinf/cosf of the C library, the double camera model) and they stay resident, exactly the reference's "precomputed projection".
// The blend's expf is the exact port of the host's (expf_bits.cuh): bit-identical to the reference build for gamma == 1 (its default).
// gamma != 1 goes through the device's powf: stated tolerance (tests/test_gpu_preview.py).
#include <cmath>
#include <vector>
#include "expf_bits.cuh"
#include "kernels.h"
#include "rig.h"

namespace s360 {

// demosaic + fades of one camera: raw w x h uint16 -> (w/2) x (h/2) float4 (B, G, R, alpha)
__global__ void __launch_bounds__(256)
k_preview_prep(const uint16_t* __restrict__ raw, int w, int h, float gamma, int topDown, float4* __restrict__ dst) {
  const int dw = w / 2, dh = h / 2;
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
  if (x >= dw || y >= dh) return;
  // static const float kMaxValue16 = 2^16-1; kMaxValue32 = 2^32-1;  -- `^` is XOR in C++: 2 ^ 15 = 13 and 2 ^ 31 = 29
  const float kMaxValue16 = 13.0f, kMaxValue32 = 29.0f;
  const uint16_t* r0 = raw + (size_t)(2 * y) * w + 2 * x;
  const uint16_t* r1 = r0 + w;
  const float g1 = (float)r0[0] / kMaxValue16, g2 = (float)r1[1] / kMaxValue16;
  float b = (float)r0[1] / kMaxValue16, r = (float)r1[0] / kMaxValue16;
  float g = (g1 + g2) / 2.0f;
  if (gamma != 1.0f) { r = powf(r, gamma); g = powf(g, gamma); b = powf(b, gamma); }   // powf(x, 1) == x
  float alpha = 1.0f;                                                                   // cvtColor BGR2BGRA on CV_32F
  if (topDown) alpha *= (float)y / float(dh);                                          // topDownAlphaFade
  {                                                                                     // radialAlphaFade
    const float dx = (float)x - float(dw) / 2.0f, dy = (float)y - float(dh) / 2.0f;
    const float rr = sqrtf(dx * dx + dy * dy) / float(min(dh, dw) / 2.0f);
    alpha *= fmaxf(0.0f, 1.0f - rr);
  }
  dst[(size_t)y * dw + x] = make_float4(b * kMaxValue32, g * kMaxValue32, r * kMaxValue32, alpha);
}

// cv::remap INTER_CUBIC BORDER_CONSTANT on CV_32FC4 (same two code paths as remapCubicF2 in remap.cuh)
__device__ __forceinline__ float4 remapCubicF4(const float4* __restrict__ src, int w, int h, float mx, float my, const float* __restrict__ tabF) {
  int sx, sy, tab;
  quantMap(mx, my, sx, sy, tab);
  sx -= 1;
  sy -= 1;
  if (sx >= w || sx + 4 <= 0 || sy >= h || sy + 4 <= 0) return make_float4(0.f, 0.f, 0.f, 0.f);
  const float* wt = tabF + tab * 16;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  if ((unsigned)sx < (unsigned)max(w - 3, 0) && (unsigned)sy < (unsigned)max(h - 3, 0)) {
    const float4* p = src + (size_t)sy * w + sx;
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const float4 a = __ldg(p), b = __ldg(p + 1), c = __ldg(p + 2), d = __ldg(p + 3);
      const float w0 = __ldg(wt + r * 4), w1 = __ldg(wt + r * 4 + 1), w2 = __ldg(wt + r * 4 + 2), w3 = __ldg(wt + r * 4 + 3);
      const float4 row = make_float4(a.x * w0 + b.x * w1 + c.x * w2 + d.x * w3, a.y * w0 + b.y * w1 + c.y * w2 + d.y * w3,
                                     a.z * w0 + b.z * w1 + c.z * w2 + d.z * w3, a.w * w0 + b.w * w1 + c.w * w2 + d.w * w3);
      if (r == 0) acc = row;
      else { acc.x = acc.x + row.x; acc.y = acc.y + row.y; acc.z = acc.z + row.z; acc.w = acc.w + row.w; }
      p += w;
    }
  } else {
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int yy = sy + r;
      if ((unsigned)yy >= (unsigned)h) continue;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int xx = sx + j;
        if ((unsigned)xx >= (unsigned)w) continue;
        const float4 v = __ldg(src + (size_t)yy * w + xx);
        const float wv = __ldg(wt + r * 4 + j);
        acc.x = acc.x + v.x * wv; acc.y = acc.y + v.y * wv; acc.z = acc.z + v.z * wv; acc.w = acc.w + v.w * wv;
      }
    }
  }
  return acc;
}

// projection of the three layers + flattenLayersAlphaSoftmax + the 16 -> 8 bit scale; out: eqrH x eqrW x 3 float (B, G, R)
__global__ void __launch_bounds__(256)
k_preview_render(const float4* __restrict__ layers, int lw, int lh, const float2* __restrict__ warp, int eqrW, int eqrH, float softmaxCoef,
                 float norm, const float* __restrict__ tabF, float* __restrict__ out) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
  if (x >= eqrW || y >= eqrH) return;
  const size_t px = (size_t)y * eqrW + x, plane = (size_t)eqrW * eqrH;
  float sumAlpha = 0.0f, sb = 0.0f, sg = 0.0f, sr = 0.0f;
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    const float2 m = __ldg(warp + i * plane + px);
    const float4 c = remapCubicF4(layers + (size_t)i * lw * lh, lw, lh, m.x, m.y, tabF);
    const float alpha = expfHostBits(softmaxCoef * c.w) - 1.0f;   // exp() on a float inside namespace cv is std::exp(float) = expf
    sb = sb + alpha * c.x; sg = sg + alpha * c.y; sr = sr + alpha * c.z;   // sumColor += alpha * Vec3f(...)
    sumAlpha += alpha;
  }
  const float inv = 1.f / sumAlpha;                                       // Vec / float multiplies by 1.f / alpha (core/matx.hpp)
  out[px * 3] = sb * inv * norm; out[px * 3 + 1] = sg * inv * norm; out[px * 3 + 2] = sr * inv * norm;
}

// ---- host: cameras and the precomputed warp ----
// Camera::createRescaledCamera (Camera.cpp:268-289)
static Camera rescaledCamera(const Camera& cam, float scale) {
  Camera s = cam;
  s.resolution[0] = (double)(int)(cam.resolution[0] * scale);
  s.resolution[1] = (double)(int)(cam.resolution[1] * scale);
  const float scaleX = (float)(s.resolution[0] / cam.resolution[0]), scaleY = (float)(s.resolution[1] / cam.resolution[1]);
  s.principal[0] *= scaleX; s.principal[1] *= scaleY;
  s.focal[0] *= scaleX; s.focal[1] *= scaleY;
  return s;
}
// Camera::sees (Camera.h:161-181)
static bool cameraSees(const Camera& c, const Vec3& rig) {
  if (c.fovThreshold != -1) {
    const Vec3 v = rig - c.position;
    if (c.fovThreshold == 0) { if (dot(c.backward(), v) >= 0) return false; }
    else {
      const double d = -dot(c.backward(), v);
      if (d * std::abs(d) <= c.fovThreshold * dot(v, v)) return false;
    }
  }
  double px, py;
  c.pixel(rig, px, py);
  return 0 <= px && px < c.resolution[0] && 0 <= py && py < c.resolution[1];
}
// PreviewRenderer::precomputeProjectionWarp (TestHyperPreview.cpp:114-127) for the three cameras, planes [cam][y][x]
void buildPreviewWarp(const Rig& rig, int eqrW, int eqrH, std::vector<float2>& out) {
  const int which[3] = {rig.findCameraByDirection({0, 0, 1}), rig.findCameraByDirection({0, 0, -1}), rig.findLargestDistCamAxisToRigCenter()};
  out.assign((size_t)3 * eqrW * eqrH, make_float2(-1.f, -1.f));
  for (int i = 0; i < 3; ++i) {
    const Camera cam = rescaledCamera(rig.cameras[which[i]], 0.5f);
    float2* plane = out.data() + (size_t)i * eqrW * eqrH;
    for (int y = 0; y < eqrH; ++y)
      for (int x = 0; x < eqrW; ++x) {
        const float theta = (float)(2.0f * M_PI * (1.0 - (x + 0.5f) / float(eqrW)));
        const float phi = (float)(M_PI * (y + 0.5f) / float(eqrH));
        // projectEquirectToCam(theta, phi, cam, kNearInfinity): float trig, float depth, double camera model
        const Vec3 dir{(double)(sinf(phi) * cosf(theta)), (double)(sinf(phi) * sinf(theta)), (double)cosf(phi)};
        const float depth = 1e6f;
        const Vec3 world{(double)depth * dir.x, (double)depth * dir.y, (double)depth * dir.z};
        if (cameraSees(cam, world)) {
          double px, py;
          cam.pixel(world, px, py);
          plane[(size_t)y * eqrW + x] = make_float2((float)px, (float)py);
        }
      }
  }
}

void launchPreview(cudaStream_t st, const uint16_t* const rawDev[3], int imgW, int imgH, float gamma, float4* layers, const float2* warpDev, int eqrW,
                   int eqrH, float softmaxCoef, const RemapTables& rt, float* outDev) {
  const int lw = imgW / 2, lh = imgH / 2;
  dim3 b(32, 8);
  for (int i = 0; i < 3; ++i) {
    S360_PROF(st, "k_preview_prep", (double)imgW * imgH * 2 + (double)lw * lh * 16);
    k_preview_prep<<<dim3(iDivUp(lw, 32), iDivUp(lh, 8)), b, 0, st>>>(rawDev[i], imgW, imgH, gamma, i > 0, layers + (size_t)i * lw * lh);
  }
  {
    S360_PROF(st, "k_preview_render", 3.0 * lw * lh * 16 + (double)eqrW * eqrH * (3 * 8 + 12));
    k_preview_render<<<dim3(iDivUp(eqrW, 32), iDivUp(eqrH, 8)), b, 0, st>>>(layers, lw, lh, warpDev, eqrW, eqrH, softmaxCoef, 255.0f / 65535.0f, rt.wf, outDev);
  }
  g_launchCount += 4;
}

}  // namespace s360
