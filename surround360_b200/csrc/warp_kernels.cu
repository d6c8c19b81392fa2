// sm_100a kernels for the gather/warp stages of the render path: spherical projection,
// fused novel-view synthesis, pole warp, cubic resize, cubemap. All HBM/L2-gather bound
// integer/byte work; one output pixel per thread, BGRA8 words in and out, 2D blocks of
// 32x8 threads so a warp writes 128 contiguous bytes and its 4x4 gather footprints overlap
// in L1.
#include <vector>
#include "kernels.h"

namespace s360 {

std::atomic<unsigned long long> g_launchCount{0};

// ---- per-kernel profiling (kernels.h) ----
thread_local KernelProf* t_kernelProf = nullptr;
cudaEvent_t KernelProf::get() {
  if (!pool.empty()) { cudaEvent_t e = pool.back(); pool.pop_back(); return e; }
  cudaEvent_t e;
  S360_CUDA_CHECK(cudaEventCreate(&e));
  return e;
}
void KernelProf::collect() {
  for (auto& r : pending) {
    float ms = 0;
    if (cudaEventElapsedTime(&ms, r.a, r.b) == cudaSuccess) {
      Acc& a = acc[r.name];
      a.ms += ms; a.bytes += r.bytes; a.launches += 1;
    }
    pool.push_back(r.a); pool.push_back(r.b);
  }
  pending.clear();
}
void KernelProf::clear() { acc.clear(); }
KernelProf::~KernelProf() {
  for (auto& r : pending) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
  for (auto e : pool) cudaEventDestroy(e);
}
ProfScope::ProfScope(cudaStream_t s, const char* name, double bytes) : st(s), on(t_kernelProf && t_kernelProf->on) {
  if (!on) return;
  rec.name = name; rec.bytes = bytes; rec.a = t_kernelProf->get(); rec.b = t_kernelProf->get();
  cudaEventRecord(rec.a, st);
}
ProfScope::~ProfScope() {
  if (!on) return;
  cudaEventRecord(rec.b, st);
  t_kernelProf->pending.push_back(rec);
}

static inline dim3 grid2d(int w, int h, dim3 b) { return dim3(iDivUp(w, b.x), iDivUp(h, b.y)); }

// ------------------------------------------------------------------------------------
// source preparation
// ------------------------------------------------------------------------------------
__global__ void k_prep_source(const uint8_t* __restrict__ src, int cn, int w, int h, int feather, uchar4* __restrict__ dst) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
  if (x >= w || y >= h) return;
  const uint8_t* p = src + ((size_t)y * w + x) * cn;
  uchar4 v = make_uchar4(p[0], p[1], p[2], cn == 4 ? p[3] : 255);
  if (feather > 0) {
    // TRSP:117-124: alpha = uint8(255.0f * float(y + 0.5f) / float(feather)) on rows y and h-1-y
    const int yy = y < feather ? y : (h - 1 - y < feather ? h - 1 - y : -1);
    if (yy >= 0) v.w = (unsigned char)truncU8(255.0f * float(yy + 0.5f) / float(feather));
    // rows in both ramps (h < 2*feather): the reference writes row y first then h-1-y in loop order;
    // the later write wins. Not reachable with feather*2 <= h; keep the first-ramp value otherwise.
  }
  dst[(size_t)y * w + x] = v;
}
void launchPrepSource(cudaStream_t st, const uint8_t* src, int cn, int w, int h, int feather, uchar4* dst) {
  dim3 b(32, 8);
  S360_PROF(st, "k_prep_source", (double)w * h * (cn + 4));
  k_prep_source<<<grid2d(w, h, b), b, 0, st>>>(src, cn, w, h, feather, dst);
  ++g_launchCount;
}

__global__ void k_unpack_bgra(const uint8_t* __restrict__ src, int cn, size_t n, uchar4* __restrict__ dst) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint8_t* p = src + i * cn;
  dst[i] = make_uchar4(p[0], p[1], p[2], cn == 4 ? p[3] : 255);
}
void launchUnpackBGRA(cudaStream_t st, const uint8_t* src, int cn, int w, int h, uchar4* dst) {
  const size_t n = (size_t)w * h;
  S360_PROF(st, "k_unpack_bgra", (double)n * (cn + 4));
  k_unpack_bgra<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(src, cn, n, dst);
  ++g_launchCount;
}
__global__ void k_pack_bgr(const uchar4* __restrict__ src, int w, int h, int pitch, uint8_t* __restrict__ dst) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
  if (x >= w || y >= h) return;
  const uchar4 v = src[(size_t)y * pitch + x];
  uint8_t* d = dst + ((size_t)y * w + x) * 3;
  d[0] = v.x; d[1] = v.y; d[2] = v.z;
}
void launchPackBGR(cudaStream_t st, const uchar4* src, int w, int h, int pitch, uint8_t* dst) {
  dim3 b(32, 8);
  S360_PROF(st, "k_pack_bgr", (double)w * h * 7);
  k_pack_bgr<<<grid2d(w, h, b), b, 0, st>>>(src, w, h, pitch, dst);
  ++g_launchCount;
}
__global__ void k_pack_bgra(const uchar4* __restrict__ src, int w, int h, int pitch, uchar4* __restrict__ dst) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
  if (x >= w || y >= h) return;
  dst[(size_t)y * w + x] = src[(size_t)y * pitch + x];
}
void launchPackBGRA(cudaStream_t st, const uchar4* src, int w, int h, int pitch, uint8_t* dst) {
  dim3 b(32, 8);
  S360_PROF(st, "k_pack_bgra", (double)w * h * 8);
  k_pack_bgra<<<grid2d(w, h, b), b, 0, st>>>(src, w, h, pitch, reinterpret_cast<uchar4*>(dst));
  ++g_launchCount;
}

// ------------------------------------------------------------------------------------
// spherical projection: ImageWarper.cpp:143-174 fused (map + remap), TRSP:629-637/671-678
// ------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_project_spherical(const uchar4* __restrict__ src, int sw, int sh, DevCamera cam, SphericalTables tb, uchar4* __restrict__ dst,
                    int dw, int dh, int dpitch, int poleMode, int featherStart, int featherSize, const short* __restrict__ tabI) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
  if (x >= dw || y >= dh) return;
  const float cx = __ldg(tb.cosX + x), sx = __ldg(tb.sinX + x), cy = __ldg(tb.cosY + y), sy = __ldg(tb.sinY + y);
  // unit = (cos y * cos x, cos y * sin x, sin y) as float products, then double * int(kNearInfinity)
  const double ux = (double)(cy * cx) * 1000000.0, uy = (double)(cy * sx) * 1000000.0, uz = (double)sy * 1000000.0;
  double px, py;
  cameraPixel(cam, ux, uy, uz, px, py);
  const float mx = (float)(px - 0.5), my = (float)(py - 0.5);
  uchar4 v = remapCubicBGRA<BORDER_CONST>(src, sw, sh, sw, mx, my, tabI);
  if (poleMode) {
    if (poleMode == 1) v.w = 255;
    if (y >= featherStart) {
      const float alpha = 1.0f - float(y - featherStart) / float(featherSize);
      const unsigned char ramp = (unsigned char)truncU8(255.0f * alpha);
      v.w = poleMode == 1 ? ramp : min(v.w, ramp);
    }
  }
  dst[(size_t)y * dpitch + x] = v;
}
void launchProjectSpherical(cudaStream_t st, const uchar4* src, int sw, int sh, const DevCamera& cam, const SphericalTables& tb,
                            uchar4* dst, int dw, int dh, int dstPitch, int poleMode, int featherStart, int featherSize,
                            const RemapTables& rt) {
  dim3 b(32, 8);
  S360_PROF(st, "k_project_spherical", ((double)sw * sh + (double)dw * dh) * 4);   // the source once + the projection written
  k_project_spherical<<<grid2d(dw, dh, b), b, 0, st>>>(src, sw, sh, cam, tb, dst, dw, dh, dstPitch, poleMode, featherStart,
                                                         featherSize, rt.wi);
  ++g_launchCount;
}

// ------------------------------------------------------------------------------------
// novel view: NovelView.cpp:174-224 (x4) + :101-154 (x2) per output pixel
// ------------------------------------------------------------------------------------
struct LazyPix {
  uchar4 c;
  float mag;
};
// one renderLazyNovelView sample: flow remap at (wx, wy), composed image remap, alpha scale
__device__ __forceinline__ LazyPix lazySample(const uchar4* __restrict__ img, int imgPitch, const float2* __restrict__ flow, int w,
                                              int h, float wx, float wy, float t, const RemapTables& rt) {
  const float2 fd = remapCubicF2(flow, w, h, w, wx, wy, rt.wf);
  const float mx = wx + fd.x * t, my = wy + fd.y * t;
  LazyPix r;
  r.c = remapCubicBGRA<BORDER_CONST>(img, w, h, imgPitch, mx, my, rt.wi);
  r.c.w = (unsigned char)((unsigned)__float2int_rz((1.0f - t) * (float)r.c.w) & 0xff);
  r.mag = sqrtf(fd.x * fd.x + fd.y * fd.y);
  return r;
}
// combineLazyViews for one pixel
__device__ __forceinline__ uchar4 combineLazy(const LazyPix& L, const LazyPix& R, int imageCols, const RemapTables& rt) {
  const uchar4 cL = L.c, cR = R.c;
  const unsigned char outAlpha = (float)max(cL.w, cR.w) / 255.0f > 0.1 ? 255 : 0;
  if (cL.w == 0 && cR.w == 0) return make_uchar4(0, 0, 0, outAlpha);
  if (cL.w == 0) return make_uchar4(cR.x, cR.y, cR.z, outAlpha);
  if (cR.w == 0) return make_uchar4(cL.x, cL.y, cL.z, outAlpha);
  const float magL = L.mag / float(imageCols), magR = R.mag / float(imageCols);
  float blendL = float(cL.w), blendR = float(cR.w);
  const float norm = blendL + blendR;
  blendL /= norm;
  blendR /= norm;
  const int diffSum = abs((int)cL.x - (int)cR.x) + abs((int)cL.y - (int)cR.y) + abs((int)cL.z - (int)cR.z);
  const float deghostCoef = __ldg(rt.tanh10 + diffSum);  // tanhf((diffSum / 255.0f) * 10.0f)
  const double expL = exp((double)(10.0f * blendL) * (1.0 + (double)(20.0f * magL)));
  const double expR = exp((double)(10.0f * blendR) * (1.0 + (double)(20.0f * magR)));
  const double sumExp = expL + expR + 0.00001;
  const float softmaxL = float(expL / sumExp), softmaxR = float(expR / sumExp);
  const float wL = lerpf(blendL, softmaxL, deghostCoef), wR = lerpf(blendR, softmaxR, deghostCoef);
  return make_uchar4((unsigned char)truncU8(float(cL.x) * wL + float(cR.x) * wR), (unsigned char)truncU8(float(cL.y) * wL + float(cR.y) * wR),
                     (unsigned char)truncU8(float(cL.z) * wL + float(cR.z) * wR), 255);
}

// the chunk column u of pair `pairIdx` for one eye (sign = +1 left eye, -1 right eye)
__device__ __forceinline__ uchar4 novelPixel(const uchar4* __restrict__ ovL, const uchar4* __restrict__ ovR, int pitch,
                                             const float2* __restrict__ fLR, const float2* __restrict__ fRL, int ovW, int h, int u,
                                             int v, int camW, int numNovel, int bufW, float disp, const RemapTables& rt) {
  float wx, wy, t;
  if (u < numNovel) {  // TRSP:273-285
    t = float(u) / float(numNovel);
    const float slabShift = float(camW) * 0.5f - float(numNovel - u);
    wx = slabShift + disp;
    wy = float(v);
  } else {  // unfilled LazyNovelViewBuffer entries are Point3f(0,0,0)
    wx = 0.f; wy = 0.f; t = 0.f;
  }
  const LazyPix a = lazySample(ovL, pitch, fRL, ovW, h, wx, wy, t, rt);
  const LazyPix b = lazySample(ovR, pitch, fLR, ovW, h, wx, wy, 1.0f - t, rt);
  return combineLazy(a, b, bufW, rt);  // combineLazyViews normalises by the novel view's width (NovelView.cpp:120-121)
}

__device__ __forceinline__ int offsetWrapIndex(int x, int cols, float offset) {  // CvUtil.cpp:99-113
  float srcX = float(x) - offset;
  if (srcX < 0) srcX += cols;
  if (srcX >= cols) srcX -= cols;
  const int sx = satShort(cvRoundF(srcX));
  return wrapIdx(sx, cols);
}

__global__ void __launch_bounds__(256) k_novel_view(NovelViewArgs a, RemapTables rt) {
  const int X = blockIdx.x * blockDim.x + threadIdx.x, v = blockIdx.y * blockDim.y + threadIdx.y;
  if (X >= a.eqrW || v >= a.camH) return;
#pragma unroll
  for (int eye = 0; eye < 2; ++eye) {
    const float off = eye == 0 ? a.shiftPx : -a.shiftPx;
    const int sxp = offsetWrapIndex(X, a.eqrW, off);  // column of the stacked panorama
    const int pair = sxp / a.bufW, u = sxp - pair * a.bufW;
    const int right = (pair + 1) % a.numCams;
    const uchar4* ovL = a.proj + (size_t)pair * a.camW * a.camH + (a.camW - a.overlapW);
    const uchar4* ovR = a.proj + (size_t)right * a.camW * a.camH;
    const float2* fLR = a.flowLR + (size_t)pair * a.overlapW * a.camH;
    const float2* fRL = a.flowRL + (size_t)pair * a.overlapW * a.camH;
    const uchar4 c = novelPixel(ovL, ovR, a.camW, fLR, fRL, a.overlapW, a.camH, u, v, a.camW, a.numNovel, a.bufW,
                                eye == 0 ? a.slabDisp : -a.slabDisp, rt);
    uchar4* pano = eye == 0 ? a.panoL : a.panoR;
    pano[(size_t)(v + a.padAbove) * a.eqrW + X] = c;
  }
}
void launchNovelView(cudaStream_t st, const NovelViewArgs& a, const RemapTables& rt) {
  dim3 b(32, 8);
  // per pair: two overlap images (4 B) and two flows (8 B) in; both eyes' panorama strips out
  S360_PROF(st, "k_novel_view", (double)a.numCams * a.overlapW * a.camH * (2 * 4 + 2 * 8) + 2.0 * a.eqrW * a.camH * 4);
  k_novel_view<<<grid2d(a.eqrW, a.camH, b), b, 0, st>>>(a, rt);
  ++g_launchCount;
}

__global__ void __launch_bounds__(256)
k_novel_chunk(const uchar4* ovL, const uchar4* ovR, int pitch, const float2* fLR, const float2* fRL, int w, int h, int bufW, int camW,
              int numNovel, float slabDisp, uchar4* chunkL, uchar4* chunkR, RemapTables rt) {
  const int u = blockIdx.x * blockDim.x + threadIdx.x, v = blockIdx.y * blockDim.y + threadIdx.y;
  if (u >= bufW || v >= h) return;
  chunkL[(size_t)v * bufW + u] = novelPixel(ovL, ovR, pitch, fLR, fRL, w, h, u, v, camW, numNovel, bufW, slabDisp, rt);
  chunkR[(size_t)v * bufW + u] = novelPixel(ovL, ovR, pitch, fLR, fRL, w, h, u, v, camW, numNovel, bufW, -slabDisp, rt);
}
void launchNovelViewChunk(cudaStream_t st, const uchar4* ovL, const uchar4* ovR, int pitch, const float2* flowLR, const float2* flowRL,
                          int w, int h, int bufW, int camW, int numNovel, float slabDisp, uchar4* chunkL, uchar4* chunkR,
                          const RemapTables& rt) {
  dim3 b(32, 8);
  S360_PROF(st, "k_novel_chunk", (double)w * h * 24 + 2.0 * bufW * h * 4);
  k_novel_chunk<<<grid2d(bufW, h, b), b, 0, st>>>(ovL, ovR, pitch, flowLR, flowRL, w, h, bufW, camW, numNovel, slabDisp, chunkL,
                                                   chunkR, rt);
  ++g_launchCount;
}

__global__ void k_offset_wrap(const uchar4* __restrict__ src, int w, int h, float offset, uchar4* __restrict__ dst) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
  if (x >= w || y >= h) return;
  dst[(size_t)y * w + x] = src[(size_t)y * w + offsetWrapIndex(x, w, offset)];
}
void launchOffsetWrap(cudaStream_t st, const uchar4* src, int w, int h, float offset, uchar4* dst) {
  dim3 b(32, 8);
  S360_PROF(st, "k_offset_wrap", (double)w * h * 8);
  k_offset_wrap<<<grid2d(w, h, b), b, 0, st>>>(src, w, h, offset, dst);
  ++g_launchCount;
}

// ------------------------------------------------------------------------------------
// pole warp: TRSP:483-546 (flow-ramped warp, cubic remap, seam blend, alpha ramp)
// ------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_pole_warp(PoleWarpArgs a, const short* __restrict__ tabI) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
  if (x >= a.cols || y >= a.rows) return;
  const float phi = a.poleCameraRadiusDeg * float(y + 0.5f) / float(a.rows);
  const float alphaFlow = 1.0f - rampf(phi, a.phiRampStart, a.phiMid);
  const float k = 1.0f - alphaFlow;
  auto sample = [&](int xx) {
    const float2 f = __ldg(a.flow + (size_t)y * a.extW + xx);
    const float mx = float(xx) + k * f.x, my = float(y) + k * f.y;
    return remapCubicBGRA<BORDER_CONST>(a.extFish, a.extW, a.rows, a.extW, mx, my, tabI);
  };
  uchar4 s = sample(x);
  const int maxBlendX = (int)(float(a.cols) * (1.2f - 1.0f));
  if (x < maxBlendX) {
    const uchar4 wp = sample(x + a.cols);
    const float alpha = 1.0f - rampf(float(x), float(maxBlendX) * 0.333f, float(maxBlendX) * 0.667f);
    s.x = (unsigned char)truncU8(float(wp.x) * alpha + float(s.x) * (1.0f - alpha));
    s.y = (unsigned char)truncU8(float(wp.y) * alpha + float(s.y) * (1.0f - alpha));
    s.z = (unsigned char)truncU8(float(wp.z) * alpha + float(s.z) * (1.0f - alpha));
  }
  const float alphaSide = 1.0f - rampf(phi, a.phiMid, a.phiRampEnd);
  s.w = (unsigned char)truncU8(float(s.w) * alphaSide);
  a.warped[(size_t)y * a.cols + x] = s;
}
void launchPoleWarp(cudaStream_t st, const PoleWarpArgs& a, const RemapTables& rt) {
  dim3 b(32, 8);
  S360_PROF(st, "k_pole_warp", (double)a.extW * a.rows * (8 + 4) + (double)a.cols * a.rows * 4);
  k_pole_warp<<<grid2d(a.cols, a.rows, b), b, 0, st>>>(a, rt.wi);
  ++g_launchCount;
}

// ------------------------------------------------------------------------------------
// cv::resize INTER_CUBIC 8U: 11-bit coefficients, integer horizontal pass; vertical pass in
// float for element index (x*cn+c) < (dw*cn & ~7) (OpenCV's SIMD body) and exact integers
// for the row tail. (PixFlow.h:98-99 with cn=4, TRSP:943-956 with cn=3.)
// ------------------------------------------------------------------------------------
__device__ __forceinline__ void cubicTaps11(int d, double scale, int& s, int* c) {
  float f;
  resizeCoord(d, scale, s, f);
  float k[4];
  cubicCoeffs(f, k);
#pragma unroll
  for (int i = 0; i < 4; ++i) c[i] = satShort(cvRoundF(k[i] * 2048.f));
}
template <int CN>
__global__ void __launch_bounds__(256)
k_resize_cubic8(const uchar4* __restrict__ src, int sw, int sh, int spitch, uchar4* __restrict__ dst, int dw, int dh, int dpitch,
                uint8_t* __restrict__ dstPacked3, double scaleX, double scaleY) {
  // source coordinate and 11-bit taps depend on the column (row) only: evaluated once per block (32 x 8 threads)
  __shared__ int sSx[32], sSy[8], sAx[32][4], sBy[8][4];
  const int tx = threadIdx.x, ty = threadIdx.y;
  const int x = blockIdx.x * 32 + tx, y = blockIdx.y * 8 + ty;
  if (ty == 0) cubicTaps11(min(x, dw - 1), scaleX, sSx[tx], sAx[tx]);
  else if (ty == 1 && tx < 8) cubicTaps11(min((int)blockIdx.y * 8 + tx, dh - 1), scaleY, sSy[tx], sBy[tx]);
  __syncthreads();
  if (x >= dw || y >= dh) return;
  const int sx = sSx[tx], sy = sSy[ty];
  int ax[4], by[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) { ax[j] = sAx[tx][j]; by[j] = sBy[ty][j]; }
  int rows[4][4];  // [tap row][channel]
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int yy = clampi(sy - 1 + r, 0, sh - 1);
    int v0 = 0, v1 = 0, v2 = 0, v3 = 0;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int xx = clampi(sx - 1 + j, 0, sw - 1);
      const uchar4 p = __ldg(src + (size_t)yy * spitch + xx);
      v0 += p.x * ax[j]; v1 += p.y * ax[j]; v2 += p.z * ax[j]; v3 += p.w * ax[j];
    }
    rows[r][0] = v0; rows[r][1] = v1; rows[r][2] = v2; rows[r][3] = v3;
  }
  const float scale = 1.f / (2048 * 2048);
  const float fb0 = by[0] * scale, fb1 = by[1] * scale, fb2 = by[2] * scale, fb3 = by[3] * scale;
  const int simdEnd = (dw * CN) & ~7;
  unsigned char o[4] = {0, 0, 0, 255};
#pragma unroll
  for (int c = 0; c < CN; ++c) {
    const int i = x * CN + c;
    if (i < simdEnd) {
      const float f = (float)rows[0][c] * fb0 + ((float)rows[1][c] * fb1 + ((float)rows[2][c] * fb2 + (float)rows[3][c] * fb3));
      o[c] = (unsigned char)satU8(cvRoundF(f));
    } else {
      const int s = rows[0][c] * by[0] + rows[1][c] * by[1] + rows[2][c] * by[2] + rows[3][c] * by[3];
      o[c] = (unsigned char)satU8((s + (1 << 21)) >> 22);
    }
  }
  if (dstPacked3) {
    uint8_t* d = dstPacked3 + ((size_t)y * dw + x) * 3;
    d[0] = o[0]; d[1] = o[1]; d[2] = o[2];
  } else {
    dst[(size_t)y * dpitch + x] = make_uchar4(o[0], o[1], o[2], o[3]);
  }
}
void launchResizeCubic8(cudaStream_t st, const uchar4* src, int sw, int sh, int spitch, int cn, uchar4* dst, int dw, int dh, int dpitch,
                        uint8_t* dstPacked3) {
  dim3 b(32, 8);
  S360_PROF(st, "k_resize_cubic8", (double)sw * sh * 4 + (double)dw * dh * (dstPacked3 ? 3 : 4));
  if (cn == 4) k_resize_cubic8<4><<<grid2d(dw, dh, b), b, 0, st>>>(src, sw, sh, spitch, dst, dw, dh, dpitch, dstPacked3, resizeScale(sw, dw),
                                                                resizeScale(sh, dh));
  else k_resize_cubic8<3><<<grid2d(dw, dh, b), b, 0, st>>>(src, sw, sh, spitch, dst, dw, dh, dpitch, dstPacked3, resizeScale(sw, dw),
                                                                resizeScale(sh, dh));
  ++g_launchCount;
}

// ------------------------------------------------------------------------------------
// cubemap: ImageWarper.cpp:26-141 + CvUtil.cpp:117-138
// ------------------------------------------------------------------------------------
enum { FACE_BACK, FACE_LEFT, FACE_TOP, FACE_BOTTOM, FACE_FRONT, FACE_RIGHT };
// acosf as the CPU path's C library computes it (fdlibm e_acosf.c rational approximation;
// only IEEE + - * / sqrt, so with -fmad=false the device result is the same float).
// tools/check_acosf_port.c verifies this restatement against glibc 2.39 on every float in [-1, 1].
__device__ __forceinline__ float acosfRef(float x) {
  const float one = 1.0f, pi = 3.1415925026e+00f, pio2_hi = 1.5707962513e+00f, pio2_lo = 7.5497894159e-08f,
              pS0 = 1.6666667163e-01f, pS1 = -3.2556581497e-01f, pS2 = 2.0121252537e-01f, pS3 = -4.0055535734e-02f,
              pS4 = 7.9153501429e-04f, pS5 = 3.4793309169e-05f, qS1 = -2.4033949375e+00f, qS2 = 2.0209457874e+00f,
              qS3 = -6.8828397989e-01f, qS4 = 7.7038154006e-02f;
  const int hx = __float_as_int(x), ix = hx & 0x7fffffff;
  if (ix == 0x3f800000) return hx > 0 ? 0.0f : pi + 2.0f * pio2_lo;
  if (ix > 0x3f800000) return (x - x) / (x - x);
  if (ix < 0x3f000000) {
    if (ix <= 0x32800000) return pio2_hi + pio2_lo;
    const float z = x * x;
    const float p = z * (pS0 + z * (pS1 + z * (pS2 + z * (pS3 + z * (pS4 + z * pS5)))));
    const float q = one + z * (qS1 + z * (qS2 + z * (qS3 + z * qS4)));
    const float r = p / q;
    return pio2_hi - (x - (pio2_lo - x * r));
  }
  if (hx < 0) {
    const float z = (one + x) * 0.5f;
    const float p = z * (pS0 + z * (pS1 + z * (pS2 + z * (pS3 + z * (pS4 + z * pS5)))));
    const float q = one + z * (qS1 + z * (qS2 + z * (qS3 + z * qS4)));
    const float s = sqrtf(z);
    const float r = p / q;
    const float w = r * s - pio2_lo;
    return pi - 2.0f * (s + w);
  }
  const float z = (one - x) * 0.5f;
  const float s = sqrtf(z);
  const float df = __int_as_float(__float_as_int(s) & 0xfffff000);
  const float c = (z - df * df) / (s + df);
  const float p = z * (pS0 + z * (pS1 + z * (pS2 + z * (pS3 + z * (pS4 + z * pS5)))));
  const float q = one + z * (qS1 + z * (qS2 + z * (qS3 + z * qS4)));
  const float r = p / q;
  const float w = r * s + c;
  return 2.0f * (df + w);
}
__device__ __forceinline__ void cubemapCoord(float x, float y, int face, int srcW, int srcH, float fov, float& srcX, float& srcY) {
  float d0 = x, d1 = y, d2 = 0.5f, o0 = x, o1 = y, o2 = 0.5f;
  switch (face) {
    case FACE_BACK: o0 = d0; o1 = d2; o2 = -d1; break;
    case FACE_LEFT: o0 = -d2; o1 = d0; o2 = -d1; break;
    case FACE_TOP: break;
    case FACE_BOTTOM: o0 = d0; o1 = -d1; o2 = -d2; break;
    case FACE_FRONT: o0 = -d0; o1 = -d2; o2 = -d1; break;
    case FACE_RIGHT: o0 = d2; o1 = -d0; o2 = -d1; break;
  }
  const float r = sqrtf(o0 * o0 + o1 * o1);
  const double nrm = sqrt((double)o0 * o0 + (double)o1 * o1 + (double)o2 * o2);
  const float phi = acosfRef((float)((double)o2 / nrm));
  float theta = r > 0.0f ? acosfRef(fabsf(o0 / r)) : 0.0f;
  const double PI = 3.14159265358979323846;
  if (o0 > 0 && o1 > 0) {}
  else if (o0 <= 0 && o1 > 0) theta = (float)(PI - (double)theta);
  else if (o0 <= 0 && o1 <= 0) theta = (float)(PI + (double)theta);
  else theta = (float)(2 * PI - (double)theta);
  const float phiP = phi < 0.0f ? 0.0f : phi > fov ? fov : phi;
  const float twoPi = (float)((double)2.0f * PI);
  const float thetaP = theta < 0.0f ? 0.0f : theta > twoPi ? twoPi : theta;
  srcX = (float)((double)(float(srcW) * thetaP) / ((double)2.0f * PI));
  srcY = float(srcH) * phiP / fov;
}
__global__ void __launch_bounds__(256)
k_cubemap(const uchar4* __restrict__ eye, int w, int h, int fw, int fh, int photo, uint8_t* __restrict__ out, const short* __restrict__ tabI) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x, j = blockIdx.y * blockDim.y + threadIdx.y, f = blockIdx.z;
  if (i >= fw || j >= fh) return;
  // face order of convertSphericalToCubemapBicubicRemap: RIGHT, LEFT, TOP, BOTTOM, BACK, FRONT
  const int faces[6] = {FACE_RIGHT, FACE_LEFT, FACE_TOP, FACE_BOTTOM, FACE_BACK, FACE_FRONT};
  const float dy = 1.0f / float(fw), dx = 1.0f / float(fh);
  float sx, sy;
  cubemapCoord(float(i) * dy - 0.5f, float(j) * dx - 0.5f, faces[f], w, h, (float)3.14159265358979323846, sx, sy);
  const uchar4 v = remapCubicBGRA<BORDER_WRAPPED>(eye, w, h, w, sx, sy, tabI);
  size_t ox, oy, outW;
  if (photo) {  // stackVertical of the six faces
    outW = fw; ox = i; oy = (size_t)f * fh + j;
  } else {  // "video": rows {faces[1],faces[0],faces[2]} / {3,4,5}, each flipped horizontally
    const int slot[6] = {1, 0, 2, 3, 4, 5};  // image index f -> position in its 3-wide strip / strip
    const int pos = slot[f];
    outW = (size_t)3 * fw; ox = (size_t)(pos % 3) * fw + (fw - 1 - i); oy = (size_t)(pos / 3) * fh + j;
  }
  uint8_t* d = out + (oy * outW + ox) * 3;
  d[0] = v.x; d[1] = v.y; d[2] = v.z;
}
void launchCubemap(cudaStream_t st, const uchar4* eye, int w, int h, int faceW, int faceH, int photo, uint8_t* outPacked3,
                   const RemapTables& rt) {
  dim3 b(32, 8);
  dim3 g(iDivUp(faceW, 32), iDivUp(faceH, 8), 6);
  S360_PROF(st, "k_cubemap", (double)w * h * 4 + 6.0 * faceW * faceH * 3);
  k_cubemap<<<g, b, 0, st>>>(eye, w, h, faceW, faceH, photo, outPacked3, rt.wi);
  ++g_launchCount;
}

}  // namespace s360
