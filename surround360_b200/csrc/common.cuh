// Shared device/host helpers for the sm_100a kernels of the render path.
// All float arithmetic in this tree is compiled with -fmad=false: the reference CPU build
// has no FMA contraction (surround360_render/CMakeLists.txt:33-35), and several stages
// branch on strict float compares (PixFlow.h:431), so op-for-op IEEE evaluation is what
// makes the results reproducible against the CPU path.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>
#include <stdexcept>
#include <string>

#define S360_CUDA_CHECK(expr)                                                                        \
  do {                                                                                               \
    cudaError_t _e = (expr);                                                                         \
    if (_e != cudaSuccess)                                                                           \
      throw std::runtime_error(std::string("CUDA error: ") + cudaGetErrorString(_e) + " at " + __FILE__ + ":" + \
                               std::to_string(__LINE__));                                            \
  } while (0)

namespace s360 {

// number of SMs on B200; grids of persistent / grid-stride kernels are sized from it
constexpr int kNumSMs = 148;

__host__ __device__ __forceinline__ int iDivUp(int a, int b) { return (a + b - 1) / b; }

#ifdef __CUDACC__
// ---- OpenCV scalar semantics -------------------------------------------------------
// cvRound: round-half-even, INT_MIN on overflow/NaN (SSE cvtss2si)
__device__ __forceinline__ int cvRoundF(float v) {
  if (!(v > -2147483648.f && v < 2147483648.f)) return (int)0x80000000;
  return __float2int_rn(v);
}
__device__ __forceinline__ int cvFloorF(float v) { return __float2int_rd(v); }
__device__ __forceinline__ int satShort(int v) { return min(max(v, -32768), 32767); }
__device__ __forceinline__ int satU8(int v) { return min(max(v, 0), 255); }
// float -> uchar by truncation (C++ conversion used by the reference's hand loops)
__device__ __forceinline__ unsigned truncU8(float f) { return (unsigned)__float2int_rz(f) & 0xffu; }
__device__ __forceinline__ int reflect101(int p, int len) {
  if ((unsigned)p < (unsigned)len) return p;
  if (len == 1) return 0;
  do {
    if (p < 0) p = -p;
    else p = len - 1 - (p - len) - 1;
  } while ((unsigned)p >= (unsigned)len);
  return p;
}
__device__ __forceinline__ int wrapIdx(int p, int len) {
  if (p < 0) p -= ((p - len + 1) / len) * len;
  if (p >= len) p %= len;
  return p;
}
__device__ __forceinline__ int clampi(int v, int lo, int hi) { return min(max(v, lo), hi); }

#endif  // __CUDACC__

inline double resizeScale(int ssize, int dsize) { return 1. / ((double)dsize / (double)ssize); }

// cv interpolateCubic (A = -0.75); float, op order as written
__host__ __device__ __forceinline__ void cubicCoeffs(float x, float* k) {
  const float A = -0.75f;
  k[0] = ((A * (x + 1) - 5 * A) * (x + 1) + 8 * A) * (x + 1) - 4 * A;
  k[1] = ((A + 2) * x - (A + 3)) * x * x + 1;
  k[2] = ((A + 2) * (1 - x) - (A + 3)) * (1 - x) * (1 - x) + 1;
  k[3] = 1.f - k[0] - k[1] - k[2];
}

#ifdef __CUDACC__
// cv::resize source coordinate: f = (float)((d+0.5)*scale - 0.5), scale = 1/((double)dsize/ssize).
// The scale is the same for every pixel of a launch: the host evaluates it (resizeScale below, same
// IEEE double operations) and passes it in, which keeps two fp64 divisions out of every thread.
__device__ __forceinline__ void resizeCoord(int d, double scale, int& s, float& f) {
  f = (float)((d + 0.5) * scale - 0.5);
  s = cvFloorF(f);
  f -= (float)s;
}

__device__ __forceinline__ float rampf(float x, float a, float b) { return fmaxf(0.0f, fminf(1.0f, (x - a) / (b - a))); }
__device__ __forceinline__ float lerpf(float x0, float x1, float a) { return x0 * (1.0f - a) + x1 * a; }

// cv::remap 1/32-px quantisation of a float map coordinate
__device__ __forceinline__ void quantMap(float mx, float my, int& sx, int& sy, int& tab) {
  const int ix = cvRoundF(mx * 32.f), iy = cvRoundF(my * 32.f);
  tab = (iy & 31) * 32 + (ix & 31);
  sx = satShort(ix >> 5);
  sy = satShort(iy >> 5);
}

#endif  // __CUDACC__

struct DevCamera {  // Camera (Camera.h:77-95) flattened for the device
  double pos[3];
  double rot[9];  // row-major
  double principal[2], distortion[2], focal[2];
  int type;  // 0 FTHETA, 1 RECTILINEAR
  int pad;
};

// Camera::pixel (Camera.h:133-140, 225-261) in double
__host__ __device__ __forceinline__ void cameraPixel(const DevCamera& c, double rx, double ry, double rz, double& px, double& py) {
  const double vx = rx - c.pos[0], vy = ry - c.pos[1], vz = rz - c.pos[2];
  const double cx = c.rot[0] * vx + c.rot[1] * vy + c.rot[2] * vz;
  const double cy = c.rot[3] * vx + c.rot[4] * vy + c.rot[5] * vz;
  const double cz = c.rot[6] * vx + c.rot[7] * vy + c.rot[8] * vz;
  double sx, sy;
  if (c.type == 0) {
    const double norm = sqrt(cx * cx + cy * cy);
    const double r = atan2(norm, -cz);
    const double r2 = r * r;
    const double s = (1 + r2 * (c.distortion[0] + r2 * c.distortion[1])) * r / norm;
    sx = s * cx;
    sy = s * cy;
  } else {
    const double plx = cx / -cz, ply = cy / -cz;
    const double r2 = plx * plx + ply * ply;
    const double f = 1 + r2 * (c.distortion[0] + r2 * c.distortion[1]);
    sx = f * plx;
    sy = f * ply;
  }
  px = c.focal[0] * sx + c.principal[0];
  py = c.focal[1] * sy + c.principal[1];
}

}  // namespace s360
