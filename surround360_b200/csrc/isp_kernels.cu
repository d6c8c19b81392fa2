// Soft ISP kernels and host pipeline (see isp.h). Reference: camera_isp/CameraIsp.h; every kernel cites the member it restates.
// Float arithmetic in the reference's order, compiled with -fmad=false; indices (reflect) as MathUtil.h:41-44.
#include <cmath>
#include <cstring>
#include <stdexcept>
#include "expf_bits.cuh"
#include "isp.h"
#include "json_mini.h"
#include "kernels.h"

namespace s360 {

// ------------------------------------------------------------------------------------------------------------------------------
struct IspDev {
  int w, h, inW, inH, resize;
  float areaRecip;
  float black[3], blackScale[3], wb[3], cmin[3], cmax[3];
  unsigned char chan[4];     // channel (0 red, 1 green, 2 blue) of the pixel at (i % 2) * 2 + (j % 2)
  float ccm[9];              // compositeCCM, row major
  float maxVal;
};
constexpr int kIspMaxBatch = 32;   // images of one size and configuration per launch (the rig has 17 cameras)
struct IspBatch { const void* raw[kIspMaxBatch]; void* out[kIspMaxBatch]; size_t strideBytes; };
__device__ __forceinline__ int ispReflect(int x, int r) { return x < 0 ? -x : x >= r ? 2 * r - x - 2 : x; }   // MathUtil.h:41-44
__device__ __forceinline__ float ispClamp(float x, float a, float b) { return x < a ? a : x > b ? b : x; }    // MathUtil.h:36-39
__device__ __forceinline__ int ispChan(const IspDev& P, int i, int j) { return P.chan[(i & 1) * 2 + (j & 1)]; }

// How k_isp_prep reads one raw sample (row pointer, column): plain 8 / 16-bit planes (what loadImage receives, CameraIsp.h:831-857),
// or the footage container's packed samples widened to 16 bits on the fly exactly as the Unpacker's RawConverter widens them before
// loadImage (RawConverter.cpp:14-59): 8-bit x 0x101; 12-bit pairs in three bytes b0 b1 b2 -> (b0 << 4 | b1 & 15), (b2 << 4 | b1 >> 4),
// each replicated to 16 bits (u << 4 | u >> 8). The values are exact integers, so the fused read equals convert + 16-bit load.
struct RawPlain8 { static __device__ __forceinline__ float at(const unsigned char* row, int x) { return float(row[x]); } };
struct RawPlain16 { static __device__ __forceinline__ float at(const unsigned char* row, int x) { return float(reinterpret_cast<const unsigned short*>(row)[x]); } };
__device__ __forceinline__ unsigned footage12Sample(const unsigned char* row, int x) {
  const unsigned char* p = row + (x >> 1) * 3;
  const unsigned b1 = p[1];
  const unsigned u = (x & 1) ? ((unsigned)p[2] << 4 | b1 >> 4) : ((unsigned)p[0] << 4 | (b1 & 0xFu));
  return (u << 4 | u >> 8) & 0xFFFFu;
}
struct RawFootage8 { static __device__ __forceinline__ float at(const unsigned char* row, int x) { return float((unsigned)row[x] * 0x101u); } };
struct RawFootage12 { static __device__ __forceinline__ float at(const unsigned char* row, int x) { return float(footage12Sample(row, x)); } };

// loadImage/resizeInput (CameraIsp.h:338-358) + blackLevelAdjust (:1106-1126) + antiVignette (:1145-1154) + whiteBalance (:1005-1021)
// + clampAndStretch (:1128-1143): raw samples -> the float Bayer plane the demosaic reads
// Four pixels per thread, 32 columns apart (every load / store instruction of a warp stays contiguous): with one 2-byte sample per
// thread the kernel had too few bytes in flight to cover the DRAM latency (final call of round 2: 450 GB/s, 7 % of the HBM peak).
constexpr int kIspPrepPx = 4;
template <typename A>
__global__ void __launch_bounds__(256)
k_isp_prep(IspBatch B, IspDev P, const float* __restrict__ vigH, const float* __restrict__ vigV, float* __restrict__ rawAll) {
  const int j0 = blockIdx.x * (32 * kIspPrepPx) + threadIdx.x, i = blockIdx.y * blockDim.y + threadIdx.y;
  if (i >= P.h) return;
  const unsigned char* __restrict__ in = reinterpret_cast<const unsigned char*>(B.raw[blockIdx.z]);
  const size_t strideBytes = B.strideBytes;
  float* __restrict__ raw = rawAll + (size_t)blockIdx.z * P.w * P.h;
  const int r = P.resize > 1 ? 2 : 1;
  float sum[kIspPrepPx];
#pragma unroll
  for (int u = 0; u < kIspPrepPx; ++u) {
    const int j = j0 + 32 * u;
    sum[u] = 0.0f;
    if (j < P.w) {
      for (int k = 0; k < P.resize; ++k) {
        const int ip = i * P.resize + k * 2;
        const int ipp = ispReflect(ip + (i % r), P.inH);
        const unsigned char* row = in + (size_t)ipp * strideBytes;
        for (int l = 0; l < P.resize; ++l) {
          const int jp = j * P.resize + l * 2;
          const int jpp = ispReflect(jp + (j % r), P.inW);
          sum[u] += A::at(row, jpp);
        }
      }
    }
  }
#pragma unroll
  for (int u = 0; u < kIspPrepPx; ++u) {
    const int j = j0 + 32 * u;
    if (j >= P.w) continue;
    float v = sum[u] * P.areaRecip;
    const int ch = ispChan(P, i, j);
    if (v < 1.0f) v = (v - P.black[ch]) * P.blackScale[ch];
    v = v * (vigH[j * 3 + ch] * vigV[i * 3 + ch]);
    v = v * P.wb[ch];
    v = ispClamp(v, 0.0f, 1.0f);
    const float lo = P.cmin[ch], hi = P.cmax[ch];
    const float t = ispClamp(v, lo, hi);
    raw[(size_t)i * P.w + j] = (t - lo) / (hi - lo);
  }
}

// colorCorrect (CameraIsp.h:1214-1242): composite CCM then the tone-curve LUT; the float index truncates like the reference's
// implicit float -> size_t conversion
__device__ __forceinline__ float4 ispColorCorrect(const IspDev& P, const float* __restrict__ lut, float p0, float p1, float p2) {
  const float kRange = float(kToneCurveLutSize - 1);
  const float t0 = ispClamp(P.ccm[0] * p0 + P.ccm[1] * p1 + P.ccm[2] * p2, 0.0f, kRange);
  const float t1 = ispClamp(P.ccm[3] * p0 + P.ccm[4] * p1 + P.ccm[5] * p2, 0.0f, kRange);
  const float t2 = ispClamp(P.ccm[6] * p0 + P.ccm[7] * p1 + P.ccm[8] * p2, 0.0f, kRange);
  return make_float4(__ldg(lut + (size_t)t0 * 3), __ldg(lut + (size_t)t1 * 3 + 1), __ldg(lut + (size_t)t2 * 3 + 2), 0.0f);
}
__device__ __forceinline__ float ispLerpHalf(float a, float b) { return a * 0.5f + b * 0.5f; }   // lerp(a, b, 0.5f), MathUtil.h:56-59
__device__ __forceinline__ bool ispRedGreenRow(const IspDev& P, int i) {   // CameraIsp.h:94-96
  return (ispChan(P, i, 0) == 0 && ispChan(P, i, 1) == 1) || (ispChan(P, i, 1) == 0 && ispChan(P, i, 0) == 1);
}

// demosaicBilinearFilter (CameraIsp.h:89-148) + colorCorrect. Every neighbour the reference reads from its r / g / b planes is a
// position where that plane holds the raw sample (reflect keeps the Bayer parity), so the raw plane is read directly.
__global__ void __launch_bounds__(256)
k_isp_demosaic_bilinear(const float* __restrict__ rawAll, IspDev P, const float* __restrict__ lut, float4* __restrict__ rgbAll) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x, i = blockIdx.y * blockDim.y + threadIdx.y;
  if (j >= P.w || i >= P.h) return;
  const int w = P.w, h = P.h;
  const size_t img = (size_t)blockIdx.z * w * h;
  const float* __restrict__ raw = rawAll + img;
  float4* __restrict__ rgb = rgbAll + img;
  const int i_1 = ispReflect(i - 1, h), i1 = ispReflect(i + 1, h), j_1 = ispReflect(j - 1, w), j1 = ispReflect(j + 1, w);
  auto R = [&](int y, int x) { return __ldg(raw + (size_t)y * w + x); };
  const int ch = ispChan(P, i, j);
  float r, g, b;
  if (ch == 0) {
    r = R(i, j);
    g = ispLerpHalf(ispLerpHalf(R(i_1, j), R(i1, j)), ispLerpHalf(R(i, j_1), R(i, j1)));
    b = ispLerpHalf(ispLerpHalf(R(i_1, j_1), R(i1, j_1)), ispLerpHalf(R(i_1, j1), R(i1, j1)));
  } else if (ch == 1) {
    g = R(i, j);
    if (ispRedGreenRow(P, i)) {
      b = (R(i_1, j) + R(i1, j)) / 2.0f;
      r = (R(i, j_1) + R(i, j1)) / 2.0f;
    } else {
      r = (R(i_1, j) + R(i1, j)) / 2.0f;
      b = (R(i, j_1) + R(i, j1)) / 2.0f;
    }
  } else {
    b = R(i, j);
    g = ispLerpHalf(ispLerpHalf(R(i_1, j), R(i1, j)), ispLerpHalf(R(i, j_1), R(i, j1)));
    r = ispLerpHalf(ispLerpHalf(R(i_1, j_1), R(i1, j_1)), ispLerpHalf(R(i_1, j1), R(i1, j1)));
  }
  rgb[(size_t)i * w + j] = ispColorCorrect(P, lut, r, g, b);
}

// demosaicEdgeAware, first loop (CameraIsp.h:191-230): vertical / horizontal green estimates and their gradients; the flag is the
// term the homogeneity test counts (dH <= dV)
__global__ void __launch_bounds__(256)
k_isp_green_candidates(const float* __restrict__ rawAll, IspDev P, float* __restrict__ gVAll, float* __restrict__ gHAll, unsigned char* __restrict__ flagAll) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x, i = blockIdx.y * blockDim.y + threadIdx.y;
  if (j >= P.w || i >= P.h) return;
  const int w = P.w, h = P.h;
  const size_t img = (size_t)blockIdx.z * w * h;
  const float* __restrict__ raw = rawAll + img;
  float* __restrict__ gV = gVAll + img;
  float* __restrict__ gH = gHAll + img;
  unsigned char* __restrict__ flag = flagAll + img;
  const int i_1 = ispReflect(i - 1, h), i1 = ispReflect(i + 1, h), i_2 = ispReflect(i - 2, h), i2 = ispReflect(i + 2, h);
  const int j_1 = ispReflect(j - 1, w), j1 = ispReflect(j + 1, w), j_2 = ispReflect(j - 2, w), j2 = ispReflect(j + 2, w);
  auto R = [&](int y, int x) { return __ldg(raw + (size_t)y * w + x); };
  const float c = R(i, j);
  float v, hz, dV, dH;
  if (ispChan(P, i, j) == 1) {
    v = c;
    hz = c;
    dV = (fabsf(R(i2, j) - c) + fabsf(c - R(i_2, j))) / 2.0f;
    dH = (fabsf(R(i, j2) - c) + fabsf(c - R(i, j_2))) / 2.0f;
  } else {
    const float n = R(i_1, j), s = R(i1, j), wv = R(i, j_1), e = R(i, j1);
    v = (n + s) / 2.0f;
    hz = (wv + e) / 2.0f;
    dV = fabsf(n - s) / 2.0f;
    dH = fabsf(wv - e) / 2.0f;
    const float u2 = R(i_2, j), d2 = R(i2, j), l2 = R(i, j_2), r2 = R(i, j2);
    v = v + (2.0f * c - u2 - d2) / 4.0f;
    hz = hz + (2.0f * c - l2 - r2) / 4.0f;
    dV = dV + fabsf(-2.0f * c + u2 + d2) / 2.0f;
    dH = dH + fabsf(-2.0f * c + l2 + r2) / 2.0f;
  }
  const size_t o = (size_t)i * w + j;
  gV[o] = v;
  gH[o] = hz;
  flag[o] = dH <= dV ? 1 : 0;
}

// second and third loop (CameraIsp.h:231-262): 9 x 9 homogeneity vote picks the green estimate; red - green / blue - green at
// the red / blue sites. The block stages the votes of its 8 x 32 pixels plus a 4-pixel apron (reflected indices) in shared memory.
__global__ void __launch_bounds__(256)
k_isp_green_select(const float* __restrict__ rawAll, const float* __restrict__ gVAll, const float* __restrict__ gHAll, const unsigned char* __restrict__ flagAll,
                   IspDev P, float* __restrict__ greenAll, float* __restrict__ diffAll) {
  __shared__ unsigned char sm[16][40];
  const int w = P.w, h = P.h;
  const size_t img = (size_t)blockIdx.z * w * h;
  const float* __restrict__ raw = rawAll + img;
  const float* __restrict__ gV = gVAll + img;
  const float* __restrict__ gH = gHAll + img;
  const unsigned char* __restrict__ flag = flagAll + img;
  float* __restrict__ green = greenAll + img;
  float* __restrict__ diff = diffAll + img;
  const int bx = blockIdx.x * 32, by = blockIdx.y * 8;
  for (int t = threadIdx.y * 32 + threadIdx.x; t < 16 * 40; t += 256) {
    const int ty = t / 40, tx = t - ty * 40;
    const int y = ispReflect(min(by + ty - 4, h + 3), h), x = ispReflect(min(bx + tx - 4, w + 3), w);   // positions past the image feed no valid output
    sm[ty][tx] = flag[(size_t)y * w + x];
  }
  __syncthreads();
  const int j = bx + threadIdx.x, i = by + threadIdx.y;
  if (j >= w || i >= h) return;
  int hCount = 0;
#pragma unroll
  for (int l = 0; l < 9; ++l)
#pragma unroll
    for (int k = 0; k < 9; ++k) hCount += sm[threadIdx.y + l][threadIdx.x + k];
  const size_t o = (size_t)i * w + j;
  const float g = hCount < 40 ? gV[o] : gH[o];   // diameterSquared / 2 = 81 / 2
  green[o] = g;
  diff[o] = ispChan(P, i, j) == 1 ? 0.0f : __ldg(raw + o) - g;
}

// last loop (CameraIsp.h:263-335): constant-hue interpolation of red and blue from the colour differences, then colorCorrect.
// The sums keep the reference's terms and order, including the repeated (i1, j2) term of the six-tap average.
__global__ void __launch_bounds__(256)
k_isp_demosaic_edge_aware(const float* __restrict__ greenAll, const float* __restrict__ diffAll, IspDev P, const float* __restrict__ lut, float4* __restrict__ rgbAll) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x, i = blockIdx.y * blockDim.y + threadIdx.y;
  if (j >= P.w || i >= P.h) return;
  const int w = P.w, h = P.h;
  const size_t img = (size_t)blockIdx.z * w * h;
  const float* __restrict__ green = greenAll + img;
  const float* __restrict__ diff = diffAll + img;
  float4* __restrict__ rgb = rgbAll + img;
  const int i_1 = ispReflect(i - 1, h), i1 = ispReflect(i + 1, h), i_2 = ispReflect(i - 2, h), i2 = ispReflect(i + 2, h);
  const int j_1 = ispReflect(j - 1, w), j1 = ispReflect(j + 1, w), j_2 = ispReflect(j - 2, w), j2 = ispReflect(j + 2, w);
  auto D = [&](int y, int x) { return __ldg(diff + (size_t)y * w + x); };
  const float g = __ldg(green + (size_t)i * w + j);
  const int ch = ispChan(P, i, j);
  float r, b;
  if (ch == 1) {
    const float ch1 = (D(i_1, j_2) + D(i_1, j) + D(i_1, j2) + D(i1, j_2) + D(i1, j2) + D(i1, j2)) / 6.0f + g;
    const float ch2 = (D(i_2, j_1) + D(i, j_1) + D(i2, j_1) + D(i_2, j1) + D(i, j1) + D(i2, j1)) / 6.0f + g;
    if (ispRedGreenRow(P, i)) { b = ch1; r = ch2; } else { r = ch1; b = ch2; }
  } else {
    const float cross = (D(i_1, j_1) + D(i1, j_1) + D(i_1, j1) + D(i1, j1)) / 4.0f + g;
    const float plus = (D(i, j) + D(i_2, j) + D(i2, j) + D(i, j_2) + D(i, j2)) / 5.0f + g;
    if (ch == 0) { b = cross; r = plus; } else { r = cross; b = plus; }
  }
  rgb[(size_t)i * w + j] = ispColorCorrect(P, lut, r, g, b);
}

// sharpenWithIirLowPass for one pixel (Filter.h:98-130) then getImage's conversion (CameraIsp.h:1275-1299): the float -> 8/16 bit
// stores truncate like the reference's implicit conversions; c0/c2 = 2/0 when swizzling to BGR. HIGH8: the 16-bit result narrowed
// the way cv::imread(CV_LOAD_IMAGE_COLOR) narrows the 16-bit PNG the reference's Unpacker writes (high byte; RigDescription.cpp:97-101
// reads what Unpacker.cpp:187-190 wrote) -- the bytes the renderer would load.
struct IspFinish { float aR, aG, aB, noiseCore, maxVal; int swizzle; };
template <typename T, bool HIGH8>
__device__ __forceinline__ void ispStorePixel(float4 p, bool sharpen, float l0, float l1, float l2, const IspFinish& F, T* __restrict__ o) {
  if (sharpen) {
    const float h0 = p.x - l0, h1 = p.y - l1, h2 = p.z - l2;
    const float n0 = 1.0f - expfHostBits(-(h0 * h0 * F.noiseCore));
    const float n1 = 1.0f - expfHostBits(-(h1 * h1 * F.noiseCore));
    const float n2 = 1.0f - expfHostBits(-(h2 * h2 * F.noiseCore));
    p.x = ispClamp(l0 + h0 * n0 * F.aR, 0.0f, F.maxVal);
    p.y = ispClamp(l1 + h1 * n1 * F.aG, 0.0f, F.maxVal);
    p.z = ispClamp(l2 + h2 * n2 * F.aB, 0.0f, F.maxVal);
  }
  const int c0 = F.swizzle ? 2 : 0, c2 = F.swizzle ? 0 : 2;
  if (HIGH8) {
    o[c0] = (T)((int)p.x >> 8);
    o[1] = (T)((int)p.y >> 8);
    o[c2] = (T)((int)p.z >> 8);
  } else {
    o[c0] = (T)(int)p.x;
    o[1] = (T)(int)p.y;
    o[c2] = (T)(int)p.z;
  }
}

// One pass of iirLowPass<ReflectBoundary, ReflectBoundary, Vec3f> (Filter.h:38-96) over the lines of an image: causal sweep into
// `buf`, anticausal sweep out of it, clamped to [0, maxVal]. A thread owns one line (a row for the horizontal pass, a column for the
// vertical one) of one image (blockIdx.y): element p of line l sits at l * lineStride + p * posStride, so the row pass needs no
// transposed copy -- a thread then reads its own 128-byte line per batch of 8 pixels, and the column pass is coalesced across the
// warp. The recurrence v = ip * (1 - alpha) + v * alpha is sequential by definition: the pixels are fetched kIspIirBatch at a time
// and the loads of the NEXT batch are in flight while the dependent lerps of the current one execute.
// FINISH (vertical pass): the low-pass value is consumed on the spot -- cored unsharp mask, clamp, narrowing, store into the output
// image -- instead of being written out and read back by another kernel.
constexpr int kIspIirBatch = 8;
template <bool FINISH, typename T, bool HIGH8>
__global__ void __launch_bounds__(128)
k_isp_iir_lines(const float4* __restrict__ inAll, int lines, int n, size_t lineStride, size_t posStride, size_t imgPixels, float alpha, float maxVal,
                float4* __restrict__ bufAll, float4* __restrict__ outAll, const float4* __restrict__ rgbAll, IspFinish F, IspBatch B) {
  const int l = blockIdx.x * blockDim.x + threadIdx.x;
  if (l >= lines) return;
  const size_t img = (size_t)blockIdx.y * imgPixels, base = (size_t)l * lineStride;
  const float4* __restrict__ in = inAll + img + base;
  float4* __restrict__ buf = bufAll + img + base;
  const float om = 1.0f - alpha;
  const float4 p0 = in[0];
  float v0 = p0.x, v1 = p0.y, v2 = p0.z;
  float4 cur[kIspIirBatch], nxt[kIspIirBatch];
  // causal: j = 1..n reads position reflect(j), writes buffer position j - 1
#pragma unroll
  for (int u = 0; u < kIspIirBatch; ++u)
    if (1 + u <= n) cur[u] = in[(size_t)ispReflect(1 + u, n) * posStride];
  for (int j0 = 1; j0 <= n; j0 += kIspIirBatch) {
#pragma unroll
    for (int u = 0; u < kIspIirBatch; ++u)
      if (j0 + kIspIirBatch + u <= n) nxt[u] = in[(size_t)ispReflect(j0 + kIspIirBatch + u, n) * posStride];
#pragma unroll
    for (int u = 0; u < kIspIirBatch; ++u)
      if (j0 + u <= n) {
        v0 = cur[u].x * om + v0 * alpha;
        v1 = cur[u].y * om + v1 * alpha;
        v2 = cur[u].z * om + v2 * alpha;
        buf[(size_t)(j0 + u - 1) * posStride] = make_float4(v0, v1, v2, 0.0f);
      }
#pragma unroll
    for (int u = 0; u < kIspIirBatch; ++u) cur[u] = nxt[u];
  }
  // anticausal: j = n-2..-1 reads buffer position reflect(j), produces position j + 1. The first batch reads what this thread wrote last.
  float4 rc[kIspIirBatch], rn[kIspIirBatch];   // FINISH: the demosaiced pixels of the positions being produced
#pragma unroll
  for (int u = 0; u < kIspIirBatch; ++u)
    if (n - 2 - u >= -1) {
      cur[u] = buf[(size_t)ispReflect(n - 2 - u, n) * posStride];
      if (FINISH) rc[u] = rgbAll[img + base + (size_t)(n - 1 - u) * posStride];
    }
  for (int j0 = n - 2; j0 >= -1; j0 -= kIspIirBatch) {
#pragma unroll
    for (int u = 0; u < kIspIirBatch; ++u)
      if (j0 - kIspIirBatch - u >= -1) {
        nxt[u] = buf[(size_t)ispReflect(j0 - kIspIirBatch - u, n) * posStride];
        if (FINISH) rn[u] = rgbAll[img + base + (size_t)(j0 - kIspIirBatch - u + 1) * posStride];
      }
#pragma unroll
    for (int u = 0; u < kIspIirBatch; ++u)
      if (j0 - u >= -1) {
        v0 = cur[u].x * om + v0 * alpha;
        v1 = cur[u].y * om + v1 * alpha;
        v2 = cur[u].z * om + v2 * alpha;
        const size_t o = base + (size_t)(j0 - u + 1) * posStride;
        const float l0 = ispClamp(v0, 0.0f, maxVal), l1 = ispClamp(v1, 0.0f, maxVal), l2 = ispClamp(v2, 0.0f, maxVal);
        if (FINISH) ispStorePixel<T, HIGH8>(rc[u], true, l0, l1, l2, F, reinterpret_cast<T*>(B.out[blockIdx.y]) + o * 3);
        else outAll[img + o] = make_float4(l0, l1, l2, 0.0f);
      }
#pragma unroll
    for (int u = 0; u < kIspIirBatch; ++u) { cur[u] = nxt[u]; if (FINISH) rc[u] = rn[u]; }
  }
}

// getImage's conversion alone (sharpening disabled: one of the three amounts is zero, CameraIsp.h:1245)
template <typename T, bool HIGH8>
__global__ void __launch_bounds__(256)
k_isp_store(const float4* __restrict__ rgb, size_t n, IspFinish F, IspBatch B) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  ispStorePixel<T, HIGH8>(rgb[(size_t)blockIdx.y * n + i], false, 0.f, 0.f, 0.f, F, reinterpret_cast<T*>(B.out[blockIdx.y]) + i * 3);
}

// RawConverter::convert12Frame (RawConverter.cpp:34-59) as a stand-alone pass (the Unpacker's --output_raw_dir dump; the ISP reads
// packed samples directly, RawFootage12). The frame is dense (three bytes per pixel pair, no row padding), so for widths that are a
// multiple of 8 it is a linear stream: a thread turns 12 bytes (three aligned 32-bit loads) into eight 16-bit samples (one 16-byte
// store): 3.5 B per pixel of compulsory traffic, all of it coalesced. blockIdx.y = image.
struct RawBatch { const void* src[kIspMaxBatch]; void* dst[kIspMaxBatch]; };
__device__ __forceinline__ unsigned rep12(unsigned u) { return (u << 4 | u >> 8) & 0xFFFFu; }
__global__ void __launch_bounds__(256)
k_raw_convert12_x8(RawBatch B, size_t groups) {
  const size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= groups) return;
  const unsigned* __restrict__ in = reinterpret_cast<const unsigned*>(B.src[blockIdx.y]) + g * 3;
  const unsigned w0 = __ldg(in), w1 = __ldg(in + 1), w2 = __ldg(in + 2);
  // bytes b0..b11 little endian in w0 w1 w2; pair k uses bytes 3k, 3k+1, 3k+2
  auto byteAt = [&](int i) -> unsigned { const unsigned w = i < 4 ? w0 : i < 8 ? w1 : w2; return (w >> (8 * (i & 3))) & 0xFFu; };
  unsigned px[8];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const unsigned b0 = byteAt(3 * k), b1 = byteAt(3 * k + 1), b2 = byteAt(3 * k + 2);
    px[2 * k] = rep12(b0 << 4 | (b1 & 0xFu));
    px[2 * k + 1] = rep12(b2 << 4 | b1 >> 4);
  }
  uint4 o;
  o.x = px[0] | px[1] << 16; o.y = px[2] | px[3] << 16; o.z = px[4] | px[5] << 16; o.w = px[6] | px[7] << 16;
  reinterpret_cast<uint4*>(B.dst[blockIdx.y])[g] = o;
}
// any even width (or unaligned frames): one thread per pixel pair, byte loads
__global__ void __launch_bounds__(256)
k_raw_convert12_pairs(RawBatch B, size_t pairs) {
  const size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= pairs) return;
  const unsigned char* __restrict__ p = reinterpret_cast<const unsigned char*>(B.src[blockIdx.y]) + g * 3;
  const unsigned b0 = p[0], b1 = p[1], b2 = p[2];
  unsigned short* o = reinterpret_cast<unsigned short*>(B.dst[blockIdx.y]) + g * 2;
  o[0] = (unsigned short)rep12(b0 << 4 | (b1 & 0xFu));
  o[1] = (unsigned short)rep12(b2 << 4 | b1 >> 4);
}
// RawConverter::convert8Frame (RawConverter.cpp:14-32): value * 0x101; four samples per thread
__global__ void __launch_bounds__(256)
k_raw_convert8(RawBatch B, size_t n) {
  const size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const unsigned char* __restrict__ in = reinterpret_cast<const unsigned char*>(B.src[blockIdx.y]);
  unsigned short* out = reinterpret_cast<unsigned short*>(B.dst[blockIdx.y]);
  for (size_t i = g * 4; i < n && i < g * 4 + 4; ++i) out[i] = (unsigned short)(in[i] * 0x101u);
}

void rawConvertBatch(cudaStream_t st, int n, const void* const* srcDev, int bits, int w, int h, void* const* dstDev) {
  if (n < 1 || n > kIspMaxBatch) throw std::runtime_error("raw batch size out of range (1.." + std::to_string(kIspMaxBatch) + ")");
  if (bits != 8 && bits != 12) throw std::runtime_error("footage samples are 8 or 12 bits per pixel");
  if (w < 1 || h < 1) throw std::runtime_error("bad image size");
  // odd widths: the reference's running byte index makes rows straddle bytes and its last pixel reads one byte past the frame
  if (bits == 12 && (w & 1)) throw std::runtime_error("12-bit footage needs an even width");
  if (!srcDev || !dstDev) throw std::runtime_error("null image");
  RawBatch B;
  std::memset(&B, 0, sizeof B);
  bool aligned = true;
  for (int k = 0; k < n; ++k) {
    if (!srcDev[k] || !dstDev[k]) throw std::runtime_error("null image");
    B.src[k] = srcDev[k]; B.dst[k] = dstDev[k];
    aligned = aligned && ((uintptr_t)srcDev[k] & 3) == 0 && ((uintptr_t)dstDev[k] & 15) == 0;
  }
  const size_t px = (size_t)w * h;
  if (bits == 8) {
    S360_PROF(st, "k_raw_convert8", (double)px * n * 3);
    k_raw_convert8<<<dim3((unsigned)((px / 4 + 256) / 256), n), 256, 0, st>>>(B, px);
  } else if ((px & 7) == 0 && aligned) {
    S360_PROF(st, "k_raw_convert12", (double)px * n * 3.5);
    k_raw_convert12_x8<<<dim3((unsigned)((px / 8 + 255) / 256), n), 256, 0, st>>>(B, px / 8);
  } else {
    S360_PROF(st, "k_raw_convert12", (double)px * n * 3.5);
    k_raw_convert12_pairs<<<dim3((unsigned)((px / 2 + 255) / 256), n), 256, 0, st>>>(B, px / 2);
  }
  ++g_launchCount;
  S360_CUDA_CHECK(cudaGetLastError());
}

// ------------------------------------------------------------------------------------------------------------------------------
// host: configuration, tables (CameraIsp.h:390-715)
// ------------------------------------------------------------------------------------------------------------------------------
namespace {
using jsonmini::Json;
using jsonmini::JsonReader;
inline float hLerp(float x0, float x1, float alpha) { return x0 * (1.0f - alpha) + x1 * alpha; }   // MathUtil.h:56-59
inline float hClamp(float x, float a, float b) { return x < a ? a : x > b ? b : x; }
// Four point DeCasteljau (CameraIsp.h:361-371)
inline float hBezier(float a, float b, float c, float d, float t) {
  return hLerp(hLerp(hLerp(a, b, t), hLerp(b, c, t), t), hLerp(hLerp(b, c, t), hLerp(c, d, t), t), t);
}
inline float highKey(float boost, float x) {   // :373-379
  const float a = 0.5f, b = hClamp(0.6666f, 0.0f, 1.0f), c = hClamp(0.8333f + boost, 0.0f, 1.0f), d = 1.0f;
  return x > 0.5f ? hBezier(a, b, c, d, (x - 0.5f) * 2.0f) : 0;
}
inline float lowKey(float boost, float x) {   // :381-387
  const float a = 0.0f, b = hClamp(0.1666f + boost, 0.0f, 1.0f), c = hClamp(0.3333f, 0.0f, 1.0f), d = 0.5f;
  return x <= 0.5f ? hBezier(a, b, c, d, x * 2.0f) : 0;
}
IspP3 readP3(const Json& j) {   // JsonUtil.cpp getVector: three doubles narrowed to float
  if (j.kind != Json::Array || j.items.size() != 3) throw std::runtime_error("JSON error. Expecting a vector to be a list with 3 elements.");
  IspP3 p;
  p.x = (float)j.items[0]->number; p.y = (float)j.items[1]->number; p.z = (float)j.items[2]->number;
  return p;
}
std::vector<IspP3> readP3List(const Json& j) {   // getCoordList
  if (j.kind != Json::Array) throw std::runtime_error("JSON error. Expecting a list of 3-element lists.");
  std::vector<IspP3> v;
  for (auto& it : j.items) v.push_back(readP3(*it));
  return v;
}
// cv::gemm's fixed-size branch for 3 x 3 CV_32F (matmul.simd.hpp): each element is the float sum a0*b0 + a1*b1 + a2*b2
void mat3Mul(const float* a, const float* b, float* d) {
  float t[9];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) t[i * 3 + j] = a[i * 3] * b[j] + a[i * 3 + 1] * b[3 + j] + a[i * 3 + 2] * b[6 + j];
  std::memcpy(d, t, sizeof t);
}
}  // namespace

// BezierCurve::operator()(t) (MathUtil.h:208-215): recursive de Casteljau over all control points, per channel
static float bezierRec(const std::vector<IspP3>& p, int c, int i, int j, float t) {
  if (i == j) return c == 0 ? p[i].x : c == 1 ? p[i].y : p[i].z;
  return hLerp(bezierRec(p, c, i, j - 1, t), bezierRec(p, c, i + 1, j, t), t);
}
void ispBezier(const std::vector<IspP3>& pts, float t, float out[3]) {
  for (int c = 0; c < 3; ++c) out[c] = bezierRec(pts, c, 0, (int)pts.size() - 1, t);
}

IspPipeline::IspPipeline(const std::string& jsonInput, int outputBpp) : outputBpp_(outputBpp) {
  if (outputBpp != 8 && outputBpp != 16) throw std::runtime_error("output bits per pixel must be 8 or 16");
  auto root = JsonReader(jsonInput).read();
  // defaults (CameraIsp.h:441-463)
  bitsPerPixel_ = 8;
  maxPixelValue_ = (1 << bitsPerPixel_) - 1;
  clampMax_ = {1.0f, 1.0f, 1.0f};
  vignetteRollOffH_.push_back({1.0f, 1.0f, 1.0f});
  vignetteRollOffV_.push_back({1.0f, 1.0f, 1.0f});
  whiteBalanceGain_ = {1.0f, 1.0f, 1.0f};
  const float eye[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
  std::memcpy(ccm_, eye, sizeof eye);
  sharpeningSupport_ = 10.0f / 2048.0f;
  gamma_ = {1.0f, 1.0f, 1.0f};
  bayerPattern_ = "GBRG";
  const Json* cfg = root->kind == Json::Object ? root->find("CameraIsp") : nullptr;
  if (cfg) {
    if (cfg->kind != Json::Object) throw std::runtime_error("JSON error. \"CameraIsp\" must be an object");
    if (auto* v = cfg->find("bitsPerPixel")) bitsPerPixel_ = (int)v->number;
    maxPixelValue_ = (1 << bitsPerPixel_) - 1;
    if (auto* v = cfg->find("blackLevel")) blackLevel_ = readP3(*v);
    if (auto* v = cfg->find("clampMin")) clampMin_ = readP3(*v);
    if (auto* v = cfg->find("clampMax")) clampMax_ = readP3(*v);
    if (auto* v = cfg->find("stuckPixelThreshold")) stuckPixelThreshold_ = (int)v->number;
    if (auto* v = cfg->find("stuckPixelDarknessThreshold")) stuckPixelDarknessThreshold_ = (float)v->number;
    if (auto* v = cfg->find("stuckPixelRadius")) stuckPixelRadius_ = 2 * (int)v->number;
    if (auto* v = cfg->find("vignetteRollOffH")) vignetteRollOffH_ = readP3List(*v);
    if (auto* v = cfg->find("vignetteRollOffV")) vignetteRollOffV_ = readP3List(*v);
    if (auto* v = cfg->find("whiteBalanceGain")) whiteBalanceGain_ = readP3(*v);
    if (auto* v = cfg->find("ccm")) {   // getMatrix
      if (v->kind != Json::Array || v->items.size() != 3) throw std::runtime_error("JSON error. Expecting a matrix being a 3 x 3 element list: [[row1], [row2], [row3]].");
      for (int r = 0; r < 3; ++r) {
        const Json& row = *v->items[r];
        if (row.kind != Json::Array || row.items.size() != 3) throw std::runtime_error("JSON error. Expecting every matrix row to be a list with 3 elements.");
        for (int c = 0; c < 3; ++c) ccm_[r * 3 + c] = (float)row.items[c]->number;
      }
    }
    if (auto* v = cfg->find("saturation")) saturation_ = (float)v->number;
    if (auto* v = cfg->find("gamma")) gamma_ = readP3(*v);
    if (auto* v = cfg->find("lowKeyBoost")) lowKeyBoost_ = readP3(*v);
    if (auto* v = cfg->find("highKeyBoost")) highKeyBoost_ = readP3(*v);
    if (auto* v = cfg->find("contrast")) contrast_ = (float)v->number;
    if (auto* v = cfg->find("sharpening")) sharpening_ = readP3(*v);
    if (auto* v = cfg->find("sharpeningSupport")) sharpeningSupport_ = (float)v->number;
    if (auto* v = cfg->find("noiseCore")) noiseCore_ = (float)v->number;
    if (auto* v = cfg->find("bayerPattern")) { if (v->kind != Json::String) throw std::runtime_error("JSON error. bayerPattern must be a string"); bayerPattern_ = v->text; }
  }
  setup();
}

void IspPipeline::setup() {
  // Bayer pattern tables (CameraIsp.h:613-663); an unknown pattern leaves the reference's tables uninitialised -- refuse it here
  struct Pat { const char* name; bool red[4], green[4]; };
  static const Pat pats[4] = {{"RGGB", {1, 0, 0, 0}, {0, 1, 1, 0}}, {"GRBG", {0, 1, 0, 0}, {1, 0, 0, 1}}, {"GBRG", {0, 0, 1, 0}, {1, 0, 0, 1}}, {"BGGR", {0, 0, 0, 1}, {0, 1, 1, 0}}};
  const Pat* pat = nullptr;
  for (auto& p : pats)
    if (bayerPattern_.find(p.name) != std::string::npos) { pat = &p; break; }
  if (!pat) throw std::runtime_error("unknown bayerPattern: " + bayerPattern_);
  for (int k = 0; k < 4; ++k) { redBayerPixel_[k / 2][k % 2] = pat->red[k]; greenBayerPixel_[k / 2][k % 2] = pat->green[k]; }
  // saturation in YUV (CameraIsp.h:675-684; matrices of ColorspaceConversion.h:22-36): satMat = yuv2rgb * diag(1, s, s) * rgb2yuv
  const float rgb2yuv[9] = {0.299f, 0.587f, 0.114f, -0.14713f, -0.28886f, 0.436f, 0.615f, -0.51499f, -0.10001f};
  const float yuv2rgb[9] = {1.0f, 0.0f, 1.13983f, 1.0f, -0.39465f, -0.58060f, 1.0f, 2.03211f, 0.0f};
  float satMat[9] = {1.0f, 0, 0, 0, saturation_, 0, 0, 0, saturation_}, tmp[9];
  mat3Mul(yuv2rgb, satMat, tmp);
  mat3Mul(tmp, rgb2yuv, satMat);
  // transpose(ccm, compositeCCM); compositeCCM *= satMat; compositeCCM *= float(kToneCurveLutSize - 1) (:686-692)
  float t[9];
  for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) t[r * 3 + c] = ccm_[c * 3 + r];
  mat3Mul(t, satMat, compositeCCM_);
  const float scale = float(kToneCurveLutSize - 1);
  for (float& v : compositeCCM_) v = v * scale;
  buildToneCurveLut();
}

void IspPipeline::buildToneCurveLut() {
  toneCurveLut_.assign((size_t)kToneCurveLutSize * 3, 0.0f);
  const float range = float((1 << outputBpp_) - 1);
  const float dx = 1.0f / float(kToneCurveLutSize - 1);
  const float angle = M_PI * 0.25f * contrast_;
  const float slope = tanf(angle);
  const float bias = 0.5f * (1.0f - slope);
  for (int i = 0; i < kToneCurveLutSize; ++i) {
    const float x = dx * i;
    float* o = &toneCurveLut_[(size_t)i * 3];
    if (disableToneCurve_) {
      const float y = x * range;
      o[0] = o[1] = o[2] = y;
    } else {
      float r = powf(x, gamma_.x), g = powf(x, gamma_.y), b = powf(x, gamma_.z);
      r = lowKey(lowKeyBoost_.x, r) + highKey(highKeyBoost_.x, r);
      g = lowKey(lowKeyBoost_.y, g) + highKey(highKeyBoost_.y, g);
      b = lowKey(lowKeyBoost_.z, b) + highKey(highKeyBoost_.z, b);
      o[0] = hClamp((slope * r + bias) * range, 0.0f, range);
      o[1] = hClamp((slope * g + bias) * range, 0.0f, range);
      o[2] = hClamp((slope * b + bias) * range, 0.0f, range);
    }
  }
  lutDirty_ = true;
}

void IspPipeline::setDemosaicFilter(int f) {
  if (f < 0 || f >= 3) throw std::runtime_error("expecting Demosaic filter in [0,2]");
  demosaicFilter_ = f;
}
void IspPipeline::setResize(int r) {
  if (r != 1 && r != 2 && r != 4 && r != 8) throw std::runtime_error("expecting a resize value of 1, 2, 4, or 8. got " + std::to_string(r));
  resize_ = r;
}

static void vignetteTables(const std::vector<IspP3>& ptsH, const std::vector<IspP3>& ptsV, int w, int h, std::vector<float>& out) {
  const int maxDimension = std::max(w, h);   // CameraIsp.h:834
  out.resize((size_t)(w + h) * 3);
  for (int x = 0; x < w; ++x) ispBezier(ptsH, float(x) / float(maxDimension), &out[(size_t)x * 3]);
  for (int y = 0; y < h; ++y) ispBezier(ptsV, float(y) / float(maxDimension), &out[(size_t)(w + y) * 3]);
}

void IspPipeline::tables(int w, int h, float* ccm9, float* lut, float* vigH, float* vigV) {
  if (w < 1 || h < 1) throw std::runtime_error("bad table size");
  std::vector<float> vig;
  vignetteTables(vignetteRollOffH_, vignetteRollOffV_, w, h, vig);
  if (ccm9) std::memcpy(ccm9, compositeCCM_, sizeof compositeCCM_);
  if (lut) std::memcpy(lut, toneCurveLut_.data(), toneCurveLut_.size() * sizeof(float));
  if (vigH) std::memcpy(vigH, vig.data(), (size_t)w * 3 * sizeof(float));
  if (vigV) std::memcpy(vigV, vig.data() + (size_t)w * 3, (size_t)h * 3 * sizeof(float));
}

IspPipeline::~IspPipeline() {
  if (dLut_) cudaFree(dLut_);
  if (dVig_) cudaFree(dVig_);
  if (dWork_) cudaFree(dWork_);
}

// per pixel and image: rgb, tA and scratch, one float4 each (48 B); the demosaic's temporaries (raw, gV, gH, green in tA; the vote
// byte and the colour difference in scratch) are dead before the sharpen stage re-uses tA (horizontal low pass) and scratch (the
// recurrences' buffer). Every array holds the n images of a batch back to back.
void IspPipeline::ensureDevice(int w, int h, int n) {
  const size_t px = (size_t)w * h, need = px * 48 * n;
  if (need > workBytes_) {
    if (dWork_) S360_CUDA_CHECK(cudaFree(dWork_));
    dWork_ = nullptr; workBytes_ = 0;
    S360_CUDA_CHECK(cudaMalloc(&dWork_, need));
    workBytes_ = need;
  }
  if (!dLut_) S360_CUDA_CHECK(cudaMalloc(&dLut_, (size_t)kToneCurveLutSize * 3 * sizeof(float)));
  if (w != vigW_ || h != vigH_ || !dVig_) {
    if (dVig_) S360_CUDA_CHECK(cudaFree(dVig_));
    dVig_ = nullptr;
    S360_CUDA_CHECK(cudaMalloc(&dVig_, (size_t)(w + h) * 3 * sizeof(float)));
    vigW_ = 0; vigH_ = 0;
  }
}

void IspPipeline::process(cudaStream_t st, const void* rawDev, int inW, int inH, int depthBits, size_t strideBytes, void* outDev, bool swizzle, bool highByte) {
  processBatch(st, 1, &rawDev, inW, inH, depthBits, strideBytes, &outDev, swizzle, highByte);
}

void IspPipeline::processBatch(cudaStream_t st, int n, const void* const* rawDev, int inW, int inH, int depthBits, size_t strideBytes, void* const* outDev,
                               bool swizzle, bool highByte, int footageBits) {
  if (n < 1 || n > kIspMaxBatch) throw std::runtime_error("ISP batch size out of range (1.." + std::to_string(kIspMaxBatch) + ")");
  if (highByte && outputBpp_ != 16) throw std::runtime_error("the high-byte output needs an ISP created with 16 output bits");
  if (depthBits != 8 && depthBits != 16) throw std::runtime_error("input is larger that 16 bits per pixel");   // the reference's message, CameraIsp.h:855
  if (footageBits) {   // packed container samples, seen by loadImage as the 16-bit plane RawConverter makes of them (Unpacker.cpp:141-180)
    if (footageBits != 8 && footageBits != 12) throw std::runtime_error("footage samples are 8 or 12 bits per pixel");
    if (footageBits == 12 && (inW & 1)) throw std::runtime_error("12-bit footage needs an even width");
    if (depthBits != 16) throw std::runtime_error("footage samples are loaded as 16-bit input");
  }
  if (!rawDev || !outDev) throw std::runtime_error("null image");
  for (int k = 0; k < n; ++k) if (!rawDev[k] || !outDev[k]) throw std::runtime_error("null image");
  if (demosaicFilter_ == 1) throw std::runtime_error("the frequency-domain demosaic filter (cv::dct) is not implemented: use 0 (bilinear) or 2 (edge aware)");
  if (stuckPixelRadius_ > 0 && stuckPixelThreshold_ <= 1)
    throw std::runtime_error("stuck-pixel removal with stuckPixelThreshold <= 1 is a sequential in-place scan and is not implemented");
  // (with stuckPixelThreshold >= 2 the reference's candidate loop `for (k = size - 1; k <= size - threshold; k--)` never runs:
  //  removeStuckPixels changes nothing, CameraIsp.h:1090-1097)
  const int w = inW / resize_, h = inH / resize_;   // loadImage, CameraIsp.h:832-833
  if (w < 8 || h < 8) throw std::runtime_error("image too small for the ISP (8 x 8 after the resize at least)");
  if (strideBytes < (footageBits ? (size_t)inW * footageBits / 8 : (size_t)inW * (depthBits / 8))) throw std::runtime_error("row stride smaller than a row");
  // loadImage: the input depth overrides bitsPerPixel; only the 16-bit case recomputes maxPixelValue (CameraIsp.h:843-853)
  if (depthBits == 8) bitsPerPixel_ = 8;
  else { bitsPerPixel_ = 16; maxPixelValue_ = (1 << bitsPerPixel_) - 1; }
  ensureDevice(w, h, n);
  if (lutDirty_) {
    S360_CUDA_CHECK(cudaMemcpyAsync(dLut_, toneCurveLut_.data(), toneCurveLut_.size() * sizeof(float), cudaMemcpyHostToDevice, st));
    S360_CUDA_CHECK(cudaStreamSynchronize(st));   // the host vector may be rebuilt right after the call
    lutDirty_ = false;
  }
  if (vigW_ != w || vigH_ != h) {
    std::vector<float> vig;
    vignetteTables(vignetteRollOffH_, vignetteRollOffV_, w, h, vig);
    S360_CUDA_CHECK(cudaMemcpyAsync(dVig_, vig.data(), vig.size() * sizeof(float), cudaMemcpyHostToDevice, st));
    S360_CUDA_CHECK(cudaStreamSynchronize(st));
    vigW_ = w; vigH_ = h;
  }
  IspDev P;
  P.w = w; P.h = h; P.inW = inW; P.inH = inH; P.resize = resize_;
  P.areaRecip = 1.0f / (maxPixelValue_ * float(resize_ * resize_));   // CameraIsp.h:340
  const float bl[3] = {blackLevel_.x, blackLevel_.y, blackLevel_.z};
  for (int c = 0; c < 3; ++c) {   // blackLevelAdjust's constants, :1107-1112
    P.black[c] = bl[c] / float(maxPixelValue_);
    P.blackScale[c] = 1.0f / (1.0f - P.black[c]);
  }
  P.wb[0] = whiteBalanceGain_.x; P.wb[1] = whiteBalanceGain_.y; P.wb[2] = whiteBalanceGain_.z;
  P.cmin[0] = clampMin_.x; P.cmin[1] = clampMin_.y; P.cmin[2] = clampMin_.z;
  P.cmax[0] = clampMax_.x; P.cmax[1] = clampMax_.y; P.cmax[2] = clampMax_.z;
  for (int k = 0; k < 4; ++k) P.chan[k] = redBayerPixel_[k / 2][k % 2] ? 0 : greenBayerPixel_[k / 2][k % 2] ? 1 : 2;
  std::memcpy(P.ccm, compositeCCM_, sizeof P.ccm);
  P.maxVal = (1 << outputBpp_) - 1.0f;   // sharpen(), CameraIsp.h:1248
  IspBatch B;
  std::memset(&B, 0, sizeof B);
  for (int k = 0; k < n; ++k) { B.raw[k] = rawDev[k]; B.out[k] = outDev[k]; }
  B.strideBytes = strideBytes;

  const size_t px = (size_t)w * h, all = px * n;
  float4* rgb = reinterpret_cast<float4*>(dWork_);
  float4* tA = rgb + all;
  float4* scratch = tA + all;
  float* raw = reinterpret_cast<float*>(tA);         // demosaic temporaries, dead before the sharpen stage writes tA / scratch
  float* gV = raw + all;
  float* gH = gV + all;
  float* green = gH + all;
  unsigned char* flag = reinterpret_cast<unsigned char*>(scratch);   // bytes [0, all)
  float* diff = reinterpret_cast<float*>(scratch) + all;             // bytes [4 all, 8 all)
  const float* vigH = dVig_;
  const float* vigV = dVig_ + (size_t)w * 3;
  const dim3 b(32, 8), g(iDivUp(w, 32), iDivUp(h, 8), n);
  {
    S360_PROF(st, "k_isp_prep", (double)all * (footageBits ? footageBits / 8.0 : depthBits / 8) * resize_ * resize_ + (double)all * 4);
    const dim3 gp(iDivUp(w, 32 * kIspPrepPx), iDivUp(h, 8), n);
    if (footageBits == 12) k_isp_prep<RawFootage12><<<gp, b, 0, st>>>(B, P, vigH, vigV, raw);
    else if (footageBits == 8) k_isp_prep<RawFootage8><<<gp, b, 0, st>>>(B, P, vigH, vigV, raw);
    else if (depthBits == 8) k_isp_prep<RawPlain8><<<gp, b, 0, st>>>(B, P, vigH, vigV, raw);
    else k_isp_prep<RawPlain16><<<gp, b, 0, st>>>(B, P, vigH, vigV, raw);
  }
  if (demosaicFilter_ == 0) {
    S360_PROF(st, "k_isp_demosaic_bilinear", (double)all * (4 + 16));
    k_isp_demosaic_bilinear<<<g, b, 0, st>>>(raw, P, dLut_, rgb);
    g_launchCount += 2;
  } else {
    { S360_PROF(st, "k_isp_green_candidates", (double)all * (4 + 9)); k_isp_green_candidates<<<g, b, 0, st>>>(raw, P, gV, gH, flag); }
    { S360_PROF(st, "k_isp_green_select", (double)all * (13 + 8)); k_isp_green_select<<<g, b, 0, st>>>(raw, gV, gH, flag, P, green, diff); }
    { S360_PROF(st, "k_isp_demosaic_edge_aware", (double)all * (8 + 16)); k_isp_demosaic_edge_aware<<<g, b, 0, st>>>(green, diff, P, dLut_, rgb); }
    g_launchCount += 4;
  }
  IspFinish Fin;
  Fin.aR = 1.0f + sharpening_.x; Fin.aG = 1.0f + sharpening_.y; Fin.aB = 1.0f + sharpening_.z;
  Fin.noiseCore = noiseCore_; Fin.maxVal = P.maxVal; Fin.swizzle = swizzle ? 1 : 0;
  const int outBytes = highByte ? 1 : outputBpp_ / 8;
  if (sharpening_.x != 0.0 && sharpening_.y != 0.0 && sharpening_.z != 0.0) {   // sharpen(), CameraIsp.h:1244-1259
    const float alpha = powf(sharpeningSupport_, 1.0f / 4.0f);                  // Filter.h:46
    float4* lpH = tA;   // horizontal pass: rgb -> lpH (buffer: scratch); the demosaic's temporaries in tA are dead by now
    {
      S360_PROF(st, "k_isp_iir_lines(rows)", (double)all * 64);
      k_isp_iir_lines<false, unsigned char, false><<<dim3(iDivUp(h, 128), n), 128, 0, st>>>(rgb, h, w, (size_t)w, 1, px, alpha, P.maxVal, scratch, lpH, nullptr, Fin, B);
    }
    {   // vertical pass of lpH, fused with the unsharp mask and the output conversion
      S360_PROF(st, "k_isp_iir_lines(cols+finish)", (double)all * (64 + 3 * outBytes));
      const dim3 gv(iDivUp(w, 128), n);
      if (highByte) k_isp_iir_lines<true, unsigned char, true><<<gv, 128, 0, st>>>(lpH, w, h, 1, (size_t)w, px, alpha, P.maxVal, scratch, nullptr, rgb, Fin, B);
      else if (outputBpp_ == 8) k_isp_iir_lines<true, unsigned char, false><<<gv, 128, 0, st>>>(lpH, w, h, 1, (size_t)w, px, alpha, P.maxVal, scratch, nullptr, rgb, Fin, B);
      else k_isp_iir_lines<true, unsigned short, false><<<gv, 128, 0, st>>>(lpH, w, h, 1, (size_t)w, px, alpha, P.maxVal, scratch, nullptr, rgb, Fin, B);
    }
    g_launchCount += 2;
  } else {
    S360_PROF(st, "k_isp_store", (double)all * (16 + 3 * outBytes));
    const dim3 blocks((unsigned)((px + 255) / 256), n);
    if (highByte) k_isp_store<unsigned char, true><<<blocks, 256, 0, st>>>(rgb, px, Fin, B);
    else if (outputBpp_ == 8) k_isp_store<unsigned char, false><<<blocks, 256, 0, st>>>(rgb, px, Fin, B);
    else k_isp_store<unsigned short, false><<<blocks, 256, 0, st>>>(rgb, px, Fin, B);
    ++g_launchCount;
  }
  S360_CUDA_CHECK(cudaGetLastError());
}

}  // namespace s360
