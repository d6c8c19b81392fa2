// sm_100a kernels for alpha feathering, layer flattening and IIR sharpening.
#include <vector>
#include "kernels.h"

namespace s360 {

static inline dim3 grid2d(int w, int h, dim3 b) { return dim3(iDivUp(w, b.x), iDivUp(h, b.y)); }

// ------------------------------------------------------------------------------------
// featherAlphaChannel (CvUtil.cpp:140-157): erode(MORPH_CROSS 2r+1) then GaussianBlur
// (r x r, sigma r/2) on the 8-bit alpha plane. The blur follows OpenCV's bit-exact 8U
// path: 8.8 fixed-point taps, BORDER_REFLECT_101, horizontal pass in 16 bits, vertical
// pass rounded by (v + 2^15) >> 16.
// ------------------------------------------------------------------------------------
__device__ __forceinline__ uchar4 readMaybeFlipped(const uchar4* __restrict__ src, int w, int srcH, int flip, int x, int y) {
  return flip ? __ldg(src + (size_t)(srcH - 1 - y) * w + (w - 1 - x)) : __ldg(src + (size_t)y * w + x);
}
__global__ void k_extract_alpha(const uchar4* __restrict__ src, int w, int srcH, int rows, int flip, uint8_t* __restrict__ a) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
  if (x >= w || y >= rows) return;
  a[(size_t)y * w + x] = readMaybeFlipped(src, w, srcH, flip, x, y).w;
}
__global__ void k_erode_cross(const uint8_t* __restrict__ a, int w, int h, int r, uint8_t* __restrict__ out) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
  if (x >= w || y >= h) return;
  int m = 255;
  const int y0 = max(y - r, 0), y1 = min(y + r, h - 1);
  for (int yy = y0; yy <= y1; ++yy) m = min(m, (int)__ldg(a + (size_t)yy * w + x));
  const int x0 = max(x - r, 0), x1 = min(x + r, w - 1);
  const uint8_t* row = a + (size_t)y * w;
  for (int xx = x0; xx <= x1; ++xx) m = min(m, (int)__ldg(row + xx));
  out[(size_t)y * w + x] = (uint8_t)m;
}
// Fast path for the default 63-wide cross (r = 31): the window minimum is built by doubling
// (windows of 2, 4, ... 32 samples; 63 = two overlapping 32-windows), two windows per instruction
// with the packed 16-bit minimum. k_erode_v31: a thread owns one column of a 32-row strip (94 rows
// in registers). k_erode_h31: a block stages 32 rows x 190 columns in shared memory, a thread owns a
// 32-column strip of one row, and the result is combined with the vertical minimum.
constexpr int kErR = 31, kErStrip = 32;
__device__ __forceinline__ void erodeWindows32(unsigned (&v)[2 * kErR + 1]) {
  // in: v[i] = s[i] | s[i+31] << 16 for a strip of 94 samples s; out (i <= 31): v[i] = min s[i..i+31] | min s[i+31..i+62] << 16
#pragma unroll
  for (int sft = 1; sft <= 16; sft <<= 1) {
#pragma unroll
    for (int i = 0; i + sft < 2 * kErR + 1; ++i) v[i] = __vminu2(v[i], v[i + sft]);
  }
}
__global__ void __launch_bounds__(128)
k_erode_v31(const uint8_t* __restrict__ a, int w, int h, uint8_t* __restrict__ out) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y0 = blockIdx.y * kErStrip;
  if (x >= w) return;
  unsigned v[2 * kErR + 1];
#pragma unroll
  for (int i = 0; i < 2 * kErR + 1; ++i) v[i] = 0;
#pragma unroll
  for (int r = 0; r < kErStrip + 2 * kErR; ++r) {
    const int y = y0 - kErR + r;
    const unsigned s = (y >= 0 && y < h) ? (unsigned)__ldg(a + (size_t)y * w + x) : 255u;  // outside the image: neutral
    if (r <= 2 * kErR) v[r] |= s;
    if (r >= kErR) v[r - kErR] |= s << 16;
  }
  erodeWindows32(v);
#pragma unroll
  for (int r = 0; r < kErStrip; ++r) {
    const int y = y0 + r;
    if (y < h) out[(size_t)y * w + x] = (uint8_t)min(v[r] & 0xffffu, v[r] >> 16);
  }
}
constexpr int kErTileW = 128, kErPitch = 196;  // 49 words per staged row: odd, so the 32 rows of a warp hit 32 banks
__global__ void __launch_bounds__(128)
k_erode_h31(const uint8_t* __restrict__ a, const uint8_t* __restrict__ vmin, int w, int h, uint8_t* __restrict__ out) {
  __shared__ uint8_t sIn[kErStrip * kErPitch];
  __shared__ uint8_t sOut[kErStrip * (kErTileW + 4)];
  const int x0 = blockIdx.x * kErTileW, y0 = blockIdx.y * kErStrip;
  for (int i = threadIdx.x; i < kErStrip * (kErTileW + 2 * kErR); i += 128) {
    const int r = i / (kErTileW + 2 * kErR), c = i - r * (kErTileW + 2 * kErR);
    const int x = x0 - kErR + c, y = y0 + r;
    sIn[r * kErPitch + c] = (x >= 0 && x < w && y < h) ? __ldg(a + (size_t)y * w + x) : (uint8_t)255;
  }
  __syncthreads();
  {
    const int r = threadIdx.x & 31, strip = threadIdx.x >> 5;
    const uint8_t* rowp = sIn + r * kErPitch + strip * 32;
    unsigned v[2 * kErR + 1];
#pragma unroll
    for (int i = 0; i < 2 * kErR + 1; ++i) v[i] = 0;
#pragma unroll
    for (int c = 0; c < 32 + 2 * kErR; ++c) {
      const unsigned s = rowp[c];
      if (c <= 2 * kErR) v[c] |= s;
      if (c >= kErR) v[c - kErR] |= s << 16;
    }
    erodeWindows32(v);
#pragma unroll
    for (int j = 0; j < 32; ++j) sOut[r * (kErTileW + 4) + strip * 32 + j] = (uint8_t)min(v[j] & 0xffffu, v[j] >> 16);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < kErStrip * kErTileW; i += 128) {
    const int r = i / kErTileW, c = i - r * kErTileW;
    const int x = x0 + c, y = y0 + r;
    if (x < w && y < h) {
      const size_t idx = (size_t)y * w + x;
      out[idx] = (uint8_t)min((int)sOut[r * (kErTileW + 4) + c], (int)__ldg(vmin + idx));
    }
  }
}
__global__ void k_blur8_rows(const uint8_t* __restrict__ a, int w, int h, const int* __restrict__ fk, int n, uint16_t* __restrict__ out) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
  if (x >= w || y >= h) return;
  const int r = n / 2;
  const uint8_t* row = a + (size_t)y * w;
  unsigned s = 0;
  for (int j = 0; j < n; ++j) s += (unsigned)__ldg(row + reflect101(x + j - r, w)) * (unsigned)__ldg(fk + j);
  out[(size_t)y * w + x] = (uint16_t)min(s, 65535u);
}
// vertical pass + write the feathered alpha next to the (possibly flipped) colours into the
// 1.2x wrap-extended image: ext(y, x) = cropped(y, x % w) (TRSP:404-411)
__global__ void k_blur8_cols_extend(const uint16_t* __restrict__ t, int w, int h, const int* __restrict__ fk, int n,
                                    const uchar4* __restrict__ src, int srcH, int flip, uchar4* __restrict__ ext, int extW) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
  if (x >= w || y >= h) return;
  const int r = n / 2;
  unsigned s = 0;
  for (int j = 0; j < n; ++j) s += (unsigned)__ldg(t + (size_t)reflect101(y + j - r, h) * w + x) * (unsigned)__ldg(fk + j);
  uchar4 c = readMaybeFlipped(src, w, srcH, flip, x, y);
  c.w = (unsigned char)min((s + (1u << 15)) >> 16, 255u);
  for (int xx = x; xx < extW; xx += w) ext[(size_t)y * extW + xx] = c;
}
// Fast paths for the default 31-tap kernel (exact integer arithmetic: any evaluation order gives the
// same sums). Rows: a block stages 8 rows x (128 + 30) samples as ints; a thread produces 4
// neighbouring outputs from 34 staged values. Columns: a thread owns one column of a 16-row strip and
// keeps the 46 rows it needs in registers.
constexpr int kB8N = 31, kB8R = 15, kB8TileW = 128, kB8Strip = 16;
__global__ void __launch_bounds__(256)
k_blur8_rows31(const uint8_t* __restrict__ a, int w, int h, const int* __restrict__ fk, uint16_t* __restrict__ out) {
  __shared__ __align__(16) int sm[8][kB8TileW + 32];
  __shared__ int sk[kB8N + 1];
  const int x0 = blockIdx.x * kB8TileW, y = blockIdx.y * 8 + threadIdx.y;
  const uint8_t* row = a + (size_t)min(y, h - 1) * w;
  for (int i = threadIdx.x; i < kB8TileW + 2 * kB8R; i += 32) {
    const int xs = reflect101(min(x0 - kB8R + i, w + kB8R - 1), w);  // columns past w+14 feed no valid output
    sm[threadIdx.y][i] = (int)__ldg(row + xs);
  }
  if (threadIdx.y == 0 && threadIdx.x < kB8N) sk[threadIdx.x] = __ldg(fk + threadIdx.x);
  __syncthreads();
  const int x = x0 + 4 * threadIdx.x;
  if (y >= h || x >= w) return;
  int v[36];
  const int4* p = reinterpret_cast<const int4*>(&sm[threadIdx.y][4 * threadIdx.x]);
#pragma unroll
  for (int j = 0; j < 9; ++j) {
    const int4 q = p[j];
    v[4 * j] = q.x; v[4 * j + 1] = q.y; v[4 * j + 2] = q.z; v[4 * j + 3] = q.w;
  }
  unsigned s0 = 0, s1 = 0, s2 = 0, s3 = 0;
#pragma unroll
  for (int t = 0; t < kB8N; ++t) {
    const unsigned c = (unsigned)sk[t];
    s0 += (unsigned)v[t] * c; s1 += (unsigned)v[t + 1] * c; s2 += (unsigned)v[t + 2] * c; s3 += (unsigned)v[t + 3] * c;
  }
  uint16_t* o = out + (size_t)y * w + x;
  o[0] = (uint16_t)min(s0, 65535u);
  if (x + 1 < w) o[1] = (uint16_t)min(s1, 65535u);
  if (x + 2 < w) o[2] = (uint16_t)min(s2, 65535u);
  if (x + 3 < w) o[3] = (uint16_t)min(s3, 65535u);
}
__global__ void __launch_bounds__(128)
k_blur8_cols_extend31(const uint16_t* __restrict__ t, int w, int h, const int* __restrict__ fk, const uchar4* __restrict__ src, int srcH,
                      int flip, uchar4* __restrict__ ext, int extW) {
  __shared__ int sk[kB8N + 1];
  if (threadIdx.x < kB8N) sk[threadIdx.x] = __ldg(fk + threadIdx.x);
  __syncthreads();
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y0 = blockIdx.y * kB8Strip;
  if (x >= w) return;
  unsigned v[kB8Strip + 2 * kB8R];
#pragma unroll
  for (int r = 0; r < kB8Strip + 2 * kB8R; ++r) {
    const int yy = reflect101(min(y0 - kB8R + r, h + kB8R - 1), h);  // rows past h+14 feed no valid output
    v[r] = (unsigned)__ldg(t + (size_t)yy * w + x);
  }
  unsigned c[kB8N];
#pragma unroll
  for (int j = 0; j < kB8N; ++j) c[j] = (unsigned)sk[j];
#pragma unroll
  for (int i = 0; i < kB8Strip; ++i) {
    const int y = y0 + i;
    if (y < h) {
      unsigned s = 0;
#pragma unroll
      for (int j = 0; j < kB8N; ++j) s += v[i + j] * c[j];
      uchar4 px = readMaybeFlipped(src, w, srcH, flip, x, y);
      px.w = (unsigned char)min((s + (1u << 15)) >> 16, 255u);
      for (int xx = x; xx < extW; xx += w) ext[(size_t)y * extW + xx] = px;
    }
  }
}
size_t featherScratchBytes(int w, int rows) { return (size_t)w * rows * 2 + (size_t)w * rows * 2 + 256; }
void launchFeatherExtend(cudaStream_t st, const uchar4* src, int w, int srcH, int rows, int flip, int erodeSize,
                         const int* blurKernelFixed, uchar4* extOut, int extW, uint8_t* scratch) {
  dim3 b(32, 8);
  const size_t plane = (size_t)w * rows;
  uint8_t* a0 = scratch;
  uint8_t* a1 = scratch + plane;
  uint16_t* t = reinterpret_cast<uint16_t*>(scratch + 2 * plane);
  { S360_PROF(st, "k_extract_alpha", (double)plane * 5);
    k_extract_alpha<<<grid2d(w, rows, b), b, 0, st>>>(src, w, srcH, rows, flip, a0); }
  if (erodeSize == kErR) {
    // the u16 scratch `t` (not needed before the blur) first holds the vertical minimum; a1 receives the eroded plane
    uint8_t* vtmp = reinterpret_cast<uint8_t*>(t);
    { S360_PROF(st, "k_erode_v31", (double)plane * 2);
      k_erode_v31<<<dim3(iDivUp(w, 128), iDivUp(rows, kErStrip)), 128, 0, st>>>(a0, w, rows, vtmp); }
    { S360_PROF(st, "k_erode_h31", (double)plane * 3);
      k_erode_h31<<<dim3(iDivUp(w, kErTileW), iDivUp(rows, kErStrip)), 128, 0, st>>>(a0, vtmp, w, rows, a1); }
    ++g_launchCount;
  } else {
    { S360_PROF(st, "k_erode_cross", (double)plane * 2);
      k_erode_cross<<<grid2d(w, rows, b), b, 0, st>>>(a0, w, rows, erodeSize, a1); }
  }
  if (erodeSize == kB8N) {
    { S360_PROF(st, "k_blur8_rows31", (double)plane * 3);
      k_blur8_rows31<<<dim3(iDivUp(w, kB8TileW), iDivUp(rows, 8)), b, 0, st>>>(a1, w, rows, blurKernelFixed, t); }
    { S360_PROF(st, "k_blur8_cols_extend31", (double)plane * 6 + (double)extW * rows * 4);
      k_blur8_cols_extend31<<<dim3(iDivUp(w, 128), iDivUp(rows, kB8Strip)), 128, 0, st>>>(t, w, rows, blurKernelFixed, src, srcH, flip, extOut, extW); }
  } else {
    { S360_PROF(st, "k_blur8_rows", (double)plane * 3);
      k_blur8_rows<<<grid2d(w, rows, b), b, 0, st>>>(a1, w, rows, blurKernelFixed, erodeSize, t); }
    { S360_PROF(st, "k_blur8_cols_extend", (double)plane * 6 + (double)extW * rows * 4);
      k_blur8_cols_extend<<<grid2d(w, rows, b), b, 0, st>>>(t, w, rows, blurKernelFixed, erodeSize, src, srcH, flip, extOut, extW); }
  }
  g_launchCount += 4;
}
void launchFeatherAlpha(cudaStream_t st, const uchar4* src, int w, int h, int erodeSize, const int* blurKernelFixed, uchar4* dst,
                        uint8_t* scratch) {
  launchFeatherExtend(st, src, w, h, h, 0, erodeSize, blurKernelFixed, dst, w, scratch);
}
__global__ void k_extend_wrap(const uchar4* __restrict__ src, int w, int rows, uchar4* __restrict__ ext, int extW) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
  if (x >= extW || y >= rows) return;
  ext[(size_t)y * extW + x] = __ldg(src + (size_t)y * w + (x % w));
}
void launchExtendWrap(cudaStream_t st, const uchar4* src, int w, int rows, uchar4* extOut, int extW) {
  dim3 b(32, 8);
  { S360_PROF(st, "k_extend_wrap", ((double)w + extW) * rows * 4);
    k_extend_wrap<<<grid2d(extW, rows, b), b, 0, st>>>(src, w, rows, extOut, extW); }
  ++g_launchCount;
}

// ------------------------------------------------------------------------------------
// pole removal (PoleRemoval.cpp:32-188; circleAlphaCut / cutRedMaskOutOfAlphaChannel CvUtil.cpp:201-222)
// ------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned char circleAlpha(int x, int y, int w, int h, float radius) {
  const float dx = float(x) - float(w) / 2.0f, dy = float(y) - float(h) / 2.0f;
  const float r = sqrtf(dx * dx + dy * dy);
  return r < radius ? 255 : 0;
}
__global__ void k_pr_prepare(const uint8_t* __restrict__ src, const uint8_t* __restrict__ mask, int w, int h, float radius, uchar4* __restrict__ dst) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
  if (x >= w || y >= h) return;
  const size_t i = (size_t)y * w + x;
  const uint8_t* p = src + i * 3;
  const uint8_t* m = mask + i * 3;
  unsigned char a = circleAlpha(x, y, w, h, radius);
  if (m[0] == 0 && m[1] == 0 && m[2] == 255) a = 0;
  dst[i] = make_uchar4(p[0], p[1], p[2], a);
}
void launchPoleRemovalPrepare(cudaStream_t st, const uint8_t* srcBGR, const uint8_t* maskBGR, int w, int h, float radius, uchar4* dst) {
  dim3 b(32, 8);
  { S360_PROF(st, "k_pr_prepare", (double)w * h * 10);
    k_pr_prepare<<<grid2d(w, h, b), b, 0, st>>>(srcBGR, maskBGR, w, h, radius, dst); }
  ++g_launchCount;
}
__global__ void k_flip_both(const uchar4* __restrict__ src, int w, int h, uchar4* __restrict__ dst) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
  if (x >= w || y >= h) return;
  dst[(size_t)y * w + x] = __ldg(src + (size_t)(h - 1 - y) * w + (w - 1 - x));
}
void launchFlipBoth(cudaStream_t st, const uchar4* src, int w, int h, uchar4* dst) {
  dim3 b(32, 8);
  { S360_PROF(st, "k_flip_both", (double)w * h * 8);
    k_flip_both<<<grid2d(w, h, b), b, 0, st>>>(src, w, h, dst); }
  ++g_launchCount;
}
__global__ void __launch_bounds__(256)
k_pr_merge(const uchar4* __restrict__ img1, const uchar4* __restrict__ img2, const float2* __restrict__ flow, int w, int h, float radius,
           uchar4* __restrict__ dst, const short* __restrict__ tabI) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
  if (x >= w || y >= h) return;
  const size_t i = (size_t)y * w + x;
  const float2 fl = __ldg(flow + i);
  const uchar4 p2 = remapCubicBGRA<BORDER_CONST>(img2, w, h, w, float(x) + fl.x, float(y) + fl.y, tabI);
  uchar4 p1 = __ldg(img1 + i);
  const float alpha = (float)p1.w / 255.0f, alpha2 = (float)p2.w / 255.0f;
  if (alpha < 1.0f && alpha2 > 0.0f) {
    const float a1 = alpha, a2 = 1.0f - alpha;
    p1.x = (unsigned char)truncU8(a1 * float(p1.x) + a2 * float(p2.x));
    p1.y = (unsigned char)truncU8(a1 * float(p1.y) + a2 * float(p2.y));
    p1.z = (unsigned char)truncU8(a1 * float(p1.z) + a2 * float(p2.z));
  }
  p1.w = circleAlpha(x, y, w, h, radius);   // circleAlphaCut overwrites the alpha of every pixel (PoleRemoval.cpp:181)
  dst[i] = p1;
}
void launchPoleRemovalMerge(cudaStream_t st, const uchar4* img1, const uchar4* img2, const float2* flow, int w, int h, float radius,
                            uchar4* dst, const RemapTables& rt) {
  dim3 b(32, 8);
  { S360_PROF(st, "k_pr_merge", (double)w * h * 20);
    k_pr_merge<<<grid2d(w, h, b), b, 0, st>>>(img1, img2, flow, w, h, radius, dst, rt.wi); }
  ++g_launchCount;
}

// ------------------------------------------------------------------------------------
// flattenLayersDeghostPreferBase (CvUtil.cpp:224-260)
// ------------------------------------------------------------------------------------
__device__ __forceinline__ uchar4 flattenPixel(uchar4 base, uchar4 top, const RemapTables& rt) {
  const int diffSum = abs((int)base.x - (int)top.x) + abs((int)base.y - (int)top.y) + abs((int)base.z - (int)top.z);
  const float deghostCoef = __ldg(rt.tanh5 + diffSum);   // tanhf((diffSum / 255.0f) * 5.0f)
  const float alphaR = (float)top.w / 255.0f;
  const float alphaL = 1.0f - alphaR;
  const float softmaxL = __ldg(rt.flatSoftL + top.w);    // float(expL / (expL + expR + 1e-5)) in double
  const float softmaxR = 1.0f - softmaxL;
  const float wL = lerpf(alphaL, softmaxL, deghostCoef), wR = lerpf(alphaR, softmaxR, deghostCoef);
  return make_uchar4((unsigned char)truncU8(float(base.x) * wL + float(top.x) * wR),
                     (unsigned char)truncU8(float(base.y) * wL + float(top.y) * wR),
                     (unsigned char)truncU8(float(base.z) * wL + float(top.z) * wR), max(top.w, base.w));
}
__global__ void k_flatten(const uchar4* __restrict__ base, const uchar4* __restrict__ top, size_t n, uchar4* __restrict__ out, RemapTables rt) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  out[i] = flattenPixel(__ldg(base + i), __ldg(top + i), rt);
}
void launchFlatten(cudaStream_t st, const uchar4* base, const uchar4* top, int w, int h, uchar4* out, const RemapTables& rt) {
  const size_t n = (size_t)w * h;
  { S360_PROF(st, "k_flatten", (double)n * 12);
    k_flatten<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(base, top, n, out, rt); }
  ++g_launchCount;
}
// TRSP:864-885: sph = flatten(sph, topWarped); sph = flip(flatten(flip(sph), bottomWarped)).
// Layers are poleRows tall and zero below (copyMakeBorder, TRSP:538-546).
__global__ void __launch_bounds__(256)
k_flatten_poles(const uchar4* __restrict__ base, int w, int h, const uchar4* __restrict__ topLayer, const uchar4* __restrict__ botLayer,
                int poleRows, uchar4* __restrict__ out, RemapTables rt) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
  if (x >= w || y >= h) return;
  uchar4 v = __ldg(base + (size_t)y * w + x);
  const uchar4 zero = make_uchar4(0, 0, 0, 0);
  if (topLayer) v = flattenPixel(v, y < poleRows ? __ldg(topLayer + (size_t)y * w + x) : zero, rt);
  if (botLayer) {
    const int fy = h - 1 - y, fx = w - 1 - x;
    v = flattenPixel(v, fy < poleRows ? __ldg(botLayer + (size_t)fy * w + fx) : zero, rt);
  }
  out[(size_t)y * w + x] = v;
}
void launchFlattenPoles(cudaStream_t st, const uchar4* base, int w, int h, const uchar4* topLayer, const uchar4* botLayer,
                        int poleRows, uchar4* out, const RemapTables& rt) {
  dim3 b(32, 8);
  { S360_PROF(st, "k_flatten_poles", (double)w * h * 8 + (double)w * poleRows * 4 * ((topLayer ? 1 : 0) + (botLayer ? 1 : 0)));
    k_flatten_poles<<<grid2d(w, h, b), b, 0, st>>>(base, w, h, topLayer, botLayer, poleRows, out, rt); }
  ++g_launchCount;
}

// ------------------------------------------------------------------------------------
// sharpen (TRSP:688-696, Filter.h:39-137). The row pass of iirLowPass is run as a column
// pass on the transposed image so that both passes walk memory with lanes on consecutive
// addresses: transpose -> columns(wrap) -> transpose -> columns(reflect) -> unsharp.
// Each recurrence is sequential by definition (one thread per line, 3 channels of ILP);
// parallelism comes from the 4096..8400 independent lines per eye.
// ------------------------------------------------------------------------------------
__global__ void k_transpose4(const uchar4* __restrict__ srcAll, int w, int h, uchar4* __restrict__ dstAll) {
  __shared__ uchar4 tile[32][33];
  const uchar4* __restrict__ src = srcAll + (size_t)w * h * blockIdx.z;
  uchar4* __restrict__ dst = dstAll + (size_t)w * h * blockIdx.z;
  const int bx = blockIdx.x * 32, by = blockIdx.y * 32;
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const int x = bx + threadIdx.x, y = by + r;
    if (x < w && y < h) tile[r][threadIdx.x] = src[(size_t)y * w + x];
  }
  __syncthreads();
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const int y = by + threadIdx.x, x = bx + r;  // dst is h-wide, w-tall
    if (x < w && y < h) dst[(size_t)x * h + y] = tile[threadIdx.x][r];
  }
}
template <bool WRAP>
__device__ __forceinline__ int iirBoundary(int i, int n) {
  if (WRAP) return i < 0 ? n + i : i >= n ? i - n : i;
  return i < 0 ? -i : i >= n ? 2 * n - i - 2 : i;
}
__device__ __forceinline__ float clamp255(float v) { return v < 0.0f ? 0.0f : v > 255.0f ? 255.0f : v; }
// in/out: w columns, n rows (line length n along y), `eyes` images back to back (blockIdx.y).
// buf: per eye 3 planes of [n][w] floats. One thread per line; the recurrence is sequential by
// definition, so the loads of the next kIirBatch rows are issued together before the dependent
// multiply-adds consume them (otherwise every step would expose a full memory round trip).
constexpr int kIirBatch = 8;
template <bool WRAP>
__global__ void __launch_bounds__(128)
k_iir_columns(const uchar4* __restrict__ inAll, int w, int n, float alpha, float* __restrict__ bufAll, uchar4* __restrict__ outAll) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  if (x >= w) return;
  const size_t plane = (size_t)w * n;
  const uchar4* __restrict__ in = inAll + plane * blockIdx.y;
  uchar4* __restrict__ out = outAll + plane * blockIdx.y;
  float* __restrict__ buf = bufAll + 3 * plane * blockIdx.y;
  const float om = 1.0f - alpha;
  const uchar4 p0 = in[x];
  float v0 = (float)p0.x, v1 = (float)p0.y, v2 = (float)p0.z;
  // Both passes are software-pipelined: the loads of the next batch are in flight while the dependent multiply-adds of the current one
  // run (one thread per line leaves only a few hundred warps on the GPU, so nothing else hides the round trip).
  auto loadIn = [&](int j0, uchar4* d) {
#pragma unroll
    for (int u = 0; u < kIirBatch; ++u)
      if (j0 + u <= n) d[u] = in[(size_t)iirBoundary<WRAP>(j0 + u, n) * w + x];
  };
  uchar4 nx[kIirBatch];
  loadIn(1, nx);
  for (int j0 = 1; j0 <= n; j0 += kIirBatch) {  // causal: j = 1..n reads row boundary(j), writes row boundary(j-1)
    uchar4 ip[kIirBatch];
#pragma unroll
    for (int u = 0; u < kIirBatch; ++u) ip[u] = nx[u];
    if (j0 + kIirBatch <= n) loadIn(j0 + kIirBatch, nx);
#pragma unroll
    for (int u = 0; u < kIirBatch; ++u)
      if (j0 + u <= n) {
        v0 = (float)ip[u].x * om + v0 * alpha;
        v1 = (float)ip[u].y * om + v1 * alpha;
        v2 = (float)ip[u].z * om + v2 * alpha;
        const size_t o = (size_t)iirBoundary<WRAP>(j0 + u - 1, n) * w + x;
        buf[o] = v0; buf[plane + o] = v1; buf[2 * plane + o] = v2;
      }
  }
  auto loadBuf = [&](int j0, float* d0, float* d1, float* d2) {
#pragma unroll
    for (int u = 0; u < kIirBatch; ++u)
      if (j0 - u >= -1) {
        const size_t o = (size_t)iirBoundary<WRAP>(j0 - u, n) * w + x;
        d0[u] = buf[o]; d1[u] = buf[plane + o]; d2[u] = buf[2 * plane + o];
      }
  };
  float n0[kIirBatch], n1[kIirBatch], n2[kIirBatch];
  loadBuf(n - 2, n0, n1, n2);
  for (int j0 = n - 2; j0 >= -1; j0 -= kIirBatch) {  // anticausal: j = n-2..-1 reads buffer row boundary(j), writes row j+1
    float b0[kIirBatch], b1[kIirBatch], b2[kIirBatch];
#pragma unroll
    for (int u = 0; u < kIirBatch; ++u) { b0[u] = n0[u]; b1[u] = n1[u]; b2[u] = n2[u]; }
    if (j0 - kIirBatch >= -1) loadBuf(j0 - kIirBatch, n0, n1, n2);
#pragma unroll
    for (int u = 0; u < kIirBatch; ++u)
      if (j0 - u >= -1) {
        v0 = b0[u] * om + v0 * alpha;
        v1 = b1[u] * om + v1 * alpha;
        v2 = b2[u] * om + v2 * alpha;
        out[(size_t)(j0 - u + 1) * w + x] =
            make_uchar4((unsigned char)truncU8(clamp255(v0)), (unsigned char)truncU8(clamp255(v1)), (unsigned char)truncU8(clamp255(v2)), 255);
      }
  }
}
__global__ void k_unsharp(uchar4* __restrict__ img, const uchar4* __restrict__ lp, size_t n, float amount) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uchar4 p = img[i];
  const uchar4 l = lp[i];
  auto one = [&](unsigned char pc, unsigned char lc) {
    const float lf = (float)lc;
    const float hp = (float)pc - lf;
    const float ng = 1.0f - expf(-((hp * hp) * 100.0f));
    return (unsigned char)truncU8(clamp255(lf + hp * ng * amount));
  };
  p.x = one(p.x, l.x); p.y = one(p.y, l.y); p.z = one(p.z, l.z);
  img[i] = p;
}
// img: `eyes` images of w x h back to back, sharpened in place.
// scratchF: eyes * 3 * w * h floats; scratchLp: 2 * eyes * w * h uchar4.
void launchSharpen(cudaStream_t st, uchar4* img, int w, int h, int eyes, float amount, float* scratchF, uchar4* scratchLp) {
  uchar4* tA = scratchLp;
  uchar4* tB = scratchLp + (size_t)eyes * w * h;
  const float alpha = powf(0.25f, 1.0f / 4.0f);
  dim3 tb(32, 8);
  const size_t n = (size_t)eyes * w * h;
  // algorithmic bytes: the reference's images are 3-channel (sharpen: 7 planes x 3 B per pixel in all); the transposes
  // exist only because the recurrence runs along columns here
  { S360_PROF(st, "k_transpose4", (double)n * 8);
    k_transpose4<<<dim3(iDivUp(w, 32), iDivUp(h, 32), eyes), tb, 0, st>>>(img, w, h, tA); }                 // tA: h wide, w tall
  { S360_PROF(st, "k_iir_columns", (double)n * 6);
    k_iir_columns<true><<<dim3(iDivUp(h, 128), eyes), 128, 0, st>>>(tA, h, w, alpha, scratchF, tB); }       // horizontal pass (wrap)
  { S360_PROF(st, "k_transpose4", (double)n * 8);
    k_transpose4<<<dim3(iDivUp(h, 32), iDivUp(w, 32), eyes), tb, 0, st>>>(tB, h, w, tA); }                 // tA: w wide, h tall
  { S360_PROF(st, "k_iir_columns", (double)n * 6);
    k_iir_columns<false><<<dim3(iDivUp(w, 128), eyes), 128, 0, st>>>(tA, w, h, alpha, scratchF, tB); }      // vertical pass (reflect)
  { S360_PROF(st, "k_unsharp", (double)n * 9);
    k_unsharp<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(img, tB, n, amount); }
  g_launchCount += 5;
}

// ------------------------------------------------------------------------------------
// Encode hand-off: the finished BGR8 frame -> NV12 (a full-resolution luma plane and an interleaved half-resolution
// Cb / Cr plane), the input surface a hardware encoder registers from a device pointer -- so that the equirects never leave HBM between
// the renderer and the encoder (the reference writes PNGs and lets ffmpeg convert to yuv420p on the CPU, scripts/run_all.py:74-88).
// Arithmetic: the 8-bit integer form of the limited-range ("TV") matrix, BT.601 (what swscale applies to untagged RGB by default) or
// BT.709; chroma = the matrix applied to the rounded mean of the 2 x 2 block (centre siting). A thread owns one 2 x 2 block: 12 bytes in,
// 6 bytes out, nothing else -- HBM-bound streaming.
// ------------------------------------------------------------------------------------
struct YuvCoef { int yr, yg, yb, ur, ug, ub, vr, vg, vb; };
__global__ void __launch_bounds__(256)
k_bgr_to_nv12(const uint8_t* __restrict__ bgr, size_t bgrStride, int w, int h, YuvCoef K, uint8_t* __restrict__ yPlane, size_t yPitch,
              uint8_t* __restrict__ uvPlane, size_t uvPitch) {
  const int bx = blockIdx.x * blockDim.x + threadIdx.x, by = blockIdx.y * blockDim.y + threadIdx.y;   // 2 x 2 block index
  const int x = bx * 2, y = by * 2;
  if (x >= w || y >= h) return;
  const int x1 = min(x + 1, w - 1), y1 = min(y + 1, h - 1);   // odd sizes: the last column / row is its own neighbour
  const uint8_t* r0 = bgr + (size_t)y * bgrStride;
  const uint8_t* r1 = bgr + (size_t)y1 * bgrStride;
  int B[4], G[4], R[4];
  const uint8_t* px[4] = {r0 + 3 * x, r0 + 3 * x1, r1 + 3 * x, r1 + 3 * x1};
#pragma unroll
  for (int k = 0; k < 4; ++k) { B[k] = px[k][0]; G[k] = px[k][1]; R[k] = px[k][2]; }
  auto luma = [&](int k) { return (uint8_t)(((K.yr * R[k] + K.yg * G[k] + K.yb * B[k] + 128) >> 8) + 16); };
  yPlane[(size_t)y * yPitch + x] = luma(0);
  if (x + 1 < w) yPlane[(size_t)y * yPitch + x + 1] = luma(1);
  if (y + 1 < h) {
    yPlane[(size_t)(y + 1) * yPitch + x] = luma(2);
    if (x + 1 < w) yPlane[(size_t)(y + 1) * yPitch + x + 1] = luma(3);
  }
  const int rm = (R[0] + R[1] + R[2] + R[3] + 2) >> 2, gm = (G[0] + G[1] + G[2] + G[3] + 2) >> 2, bm = (B[0] + B[1] + B[2] + B[3] + 2) >> 2;
  uint8_t* uv = uvPlane + (size_t)by * uvPitch + 2 * bx;
  uv[0] = (uint8_t)(((K.ur * rm + K.ug * gm + K.ub * bm + 128) >> 8) + 128);
  uv[1] = (uint8_t)(((K.vr * rm + K.vg * gm + K.vb * bm + 128) >> 8) + 128);
}
void launchBgrToNv12(cudaStream_t st, const uint8_t* bgr, size_t bgrStride, int w, int h, int matrix, uint8_t* yPlane, size_t yPitch, uint8_t* uvPlane,
                     size_t uvPitch) {
  // round(256 * 219/255 * (Kr, Kg, Kb)) and round(256 * 224/255 * chroma rows), chroma rows summing to zero (neutral greys get Cb = Cr = 128): BT.601 (Kr .299, Kb .114), BT.709 (Kr .2126, Kb .0722)
  const YuvCoef k601{66, 129, 25, -38, -74, 112, 112, -94, -18}, k709{47, 157, 16, -26, -86, 112, 112, -102, -10};
  const dim3 b(32, 8), g(iDivUp(iDivUp(w, 2), 32), iDivUp(iDivUp(h, 2), 8));
  S360_PROF(st, "k_bgr_to_nv12", (double)w * h * 4.5);
  k_bgr_to_nv12<<<g, b, 0, st>>>(bgr, bgrStride, w, h, matrix == 1 ? k709 : k601, yPlane, yPitch, uvPlane, uvPitch);
  ++g_launchCount;
  S360_CUDA_CHECK(cudaGetLastError());
}

}  // namespace s360
