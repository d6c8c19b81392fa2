// Device-side bicubic gathers with cv::remap semantics (imgwarp.cpp remapBicubic):
// 1/32-px coordinate quantisation, A=-0.75 kernel, 15-bit integer weights normalised to
// 2^15 for 8U sources, float weights for float sources, BORDER_CONSTANT(0) or BORDER_WRAP.
// Every image on the device is BGRA8 (uchar4) so a tap is one aligned 32-bit load; the
// 3-channel reference calls (ImageWarper.cpp:131-137,173 with CV_8UC3) use the same
// per-channel arithmetic and simply ignore .w.
#pragma once
#include "common.cuh"

namespace s360 {

struct RemapTables {       // built once per context on the host, resident in HBM (~100 KB)
  const short* wi;         // [1024][16]
  const float* wf;         // [1024][16]
  // Deghost terms whose arguments take few distinct values are tabulated on the host with the
  // C library the CPU path uses (tanhf/exp are not correctly rounded, and the blends truncate
  // float->uchar, so a 1-ulp difference would flip pixels whose value sits on an integer):
  const float* tanh10;     // [766] tanhf((n/255.0f) * 10.0f)   NovelView.cpp:131-138
  const float* tanh5;      // [766] tanhf((n/255.0f) * 5.0f)    CvUtil.cpp:234-241
  const float* flatSoftL;  // [256] softmaxL of flattenLayersDeghostPreferBase for top alpha a (CvUtil.cpp:242-248)
};

enum { BORDER_CONST = 0, BORDER_WRAPPED = 3 };

template <int BORDER>
__device__ __forceinline__ int borderIdx(int p, int len) {
  if ((unsigned)p < (unsigned)len) return p;
  if (BORDER == BORDER_WRAPPED) return wrapIdx(p, len);
  return -1;
}

template <int BORDER>
__device__ __forceinline__ uchar4 remapCubicBGRA(const uchar4* __restrict__ src, int w, int h, int pitchPx, float mx,
                                                 float my, const short* __restrict__ tabI) {
  int sx, sy, tab;
  quantMap(mx, my, sx, sy, tab);
  sx -= 1;
  sy -= 1;
  // 16 shorts = two 16-byte loads
  const int4* wp = reinterpret_cast<const int4*>(tabI + tab * 16);
  const int4 wa = __ldg(wp), wb = __ldg(wp + 1);
  const int wraw[8] = {wa.x, wa.y, wa.z, wa.w, wb.x, wb.y, wb.z, wb.w};
  int wt[16];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    wt[2 * i] = (int)(short)(wraw[i] & 0xffff);
    wt[2 * i + 1] = wraw[i] >> 16;
  }
  int s0 = 0, s1 = 0, s2 = 0, s3 = 0;
  if ((unsigned)sx < (unsigned)max(w - 3, 0) && (unsigned)sy < (unsigned)max(h - 3, 0)) {
    const uchar4* p = src + (size_t)sy * pitchPx + sx;
#pragma unroll
    for (int r = 0; r < 4; ++r) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const uchar4 v = __ldg(p + j);
        const int wv = wt[r * 4 + j];
        s0 += v.x * wv; s1 += v.y * wv; s2 += v.z * wv; s3 += v.w * wv;
      }
      p += pitchPx;
    }
  } else {
    if (BORDER == BORDER_CONST && (sx >= w || sx + 4 <= 0 || sy >= h || sy + 4 <= 0)) return make_uchar4(0, 0, 0, 0);
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int yy = borderIdx<BORDER>(sy + r, h);
      if (yy < 0) continue;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int xx = borderIdx<BORDER>(sx + j, w);
        if (xx < 0) continue;
        const uchar4 v = __ldg(src + (size_t)yy * pitchPx + xx);
        const int wv = wt[r * 4 + j];
        s0 += v.x * wv; s1 += v.y * wv; s2 += v.z * wv; s3 += v.w * wv;
      }
    }
  }
  return make_uchar4((unsigned char)satU8((s0 + (1 << 14)) >> 15), (unsigned char)satU8((s1 + (1 << 14)) >> 15),
                     (unsigned char)satU8((s2 + (1 << 14)) >> 15), (unsigned char)satU8((s3 + (1 << 14)) >> 15));
}

// CV_32FC2 source, BORDER_CONSTANT(0). Interior: four row sums accumulated top to bottom;
// border: sequential accumulation of the in-range taps (as OpenCV's two code paths do).
__device__ __forceinline__ float2 remapCubicF2(const float2* __restrict__ src, int w, int h, int pitchPx, float mx, float my,
                                               const float* __restrict__ tabF) {
  int sx, sy, tab;
  quantMap(mx, my, sx, sy, tab);
  sx -= 1;
  sy -= 1;
  if (sx >= w || sx + 4 <= 0 || sy >= h || sy + 4 <= 0) return make_float2(0.f, 0.f);
  const float4* wp = reinterpret_cast<const float4*>(tabF + tab * 16);
  float wt[16];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float4 q = __ldg(wp + i);
    wt[4 * i] = q.x; wt[4 * i + 1] = q.y; wt[4 * i + 2] = q.z; wt[4 * i + 3] = q.w;
  }
  float ax = 0.f, ay = 0.f;
  if ((unsigned)sx < (unsigned)max(w - 3, 0) && (unsigned)sy < (unsigned)max(h - 3, 0)) {
    const float2* p = src + (size_t)sy * pitchPx + sx;
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const float2 a = __ldg(p), b = __ldg(p + 1), c = __ldg(p + 2), d = __ldg(p + 3);
      const float rx = a.x * wt[r * 4] + b.x * wt[r * 4 + 1] + c.x * wt[r * 4 + 2] + d.x * wt[r * 4 + 3];
      const float ry = a.y * wt[r * 4] + b.y * wt[r * 4 + 1] + c.y * wt[r * 4 + 2] + d.y * wt[r * 4 + 3];
      if (r == 0) { ax = rx; ay = ry; }
      else { ax = ax + rx; ay = ay + ry; }
      p += pitchPx;
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
        const float2 v = __ldg(src + (size_t)yy * pitchPx + xx);
        ax = ax + v.x * wt[r * 4 + j];
        ay = ay + v.y * wt[r * 4 + j];
      }
    }
  }
  return make_float2(ax, ay);
}

}  // namespace s360
