// tools/emu: a minimal CPU execution model for this repository's CUDA kernels -- DEVELOPMENT TOOL ONLY.
// It lets the exact kernel source (rewritten by emu_prep.py: launches and inline PTX only) run on host
// threads so that a change can be checked bit-for-bit against the oracle before GPU time is spent on it.
// Nothing in the package, the C ABI, the tests or bench.py loads it; the product has no CPU path.
//
// Model: a launch runs synchronously. Kernels without barriers run their threads one after another; CUDA threads of the
// other kernels are fibers: a block with __syncthreads runs round-robin on one OS thread and switches at the barrier;
// kernels with warp shuffles / inter-block spin-waits (k_sweep*) get one OS thread per warp whose 32 lane-fibers advance
// in lock-step at the shuffles, every warp of the grid resident. __shared__ becomes a static array (one block at a time).
#pragma once
#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <memory>
#include <thread>
#include <vector>

#define __CUDACC__ 1
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline __attribute__((always_inline))
#define __noinline__ __attribute__((noinline))
#define __launch_bounds__(...)
#define __shared__ static
#define __align__(n) __attribute__((aligned(n)))

struct float2 { float x, y; };
struct float4 { float x, y, z, w; };
struct int4 { int x, y, z, w; };
struct uchar4 { unsigned char x, y, z, w; };
struct uint3 { unsigned x, y, z; };
struct uint4 { unsigned x, y, z, w; };
struct dim3 {
  unsigned x, y, z;
  dim3(unsigned a = 1, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {}
};
static inline float2 make_float2(float x, float y) { return {x, y}; }
static inline float4 make_float4(float x, float y, float z, float w) { return {x, y, z, w}; }
static inline uchar4 make_uchar4(unsigned char x, unsigned char y, unsigned char z, unsigned char w) { return {x, y, z, w}; }
static inline int4 make_int4(int x, int y, int z, int w) { return {x, y, z, w}; }

extern thread_local uint3 threadIdx, blockIdx;
extern thread_local dim3 blockDim, gridDim;

// ---- device intrinsics ----
template <class T> static inline T __ldg(const T* p) { return *p; }
static inline unsigned __float_as_uint(float f) { unsigned u; std::memcpy(&u, &f, 4); return u; }
static inline float __uint_as_float(unsigned u) { float f; std::memcpy(&f, &u, 4); return f; }
static inline double __longlong_as_double(long long v) { double d; std::memcpy(&d, &v, 8); return d; }
static inline long long __double_as_longlong(double d) { long long v; std::memcpy(&v, &d, 8); return v; }
static inline int __float_as_int(float f) { int u; std::memcpy(&u, &f, 4); return u; }
static inline float __int_as_float(int u) { float f; std::memcpy(&f, &u, 4); return f; }
static inline int __float2int_rn(float v) { return (int)std::nearbyintf(v); }   // default rounding mode: to nearest even
static inline int __float2int_rd(float v) { return (int)std::floor(v); }
static inline int __float2int_rz(float v) { return (int)v; }
static inline float __fmaf_rn(float a, float b, float c) { return std::fmaf(a, b, c); }
static inline unsigned __vminu2(unsigned a, unsigned b) {
  return std::min(a & 0xffffu, b & 0xffffu) | (std::min(a >> 16, b >> 16) << 16);
}
static inline int atomicOr(int* p, int v) { return __atomic_fetch_or(p, v, __ATOMIC_SEQ_CST); }
static inline int atomicExch(int* p, int v) { return __atomic_exchange_n(p, v, __ATOMIC_SEQ_CST); }
using std::max;
using std::min;
static inline int min(int a, unsigned b) { return (int)std::min<long long>(a, b); }
static inline unsigned min(unsigned a, int b) { return (unsigned)std::min<long long>(a, b); }

namespace emu {
void syncthreads();
float shfl_up(float v, unsigned delta);
float shfl_down(float v, unsigned delta);
int shfl_idx(int v, int lane);
float2 ld_volatile(const float2* p);
void st_volatile(float2* p, float2 v);
enum Mode { SEQUENTIAL, BLOCK_BARRIER, GRID_WARPS };
void launch(const char* name, dim3 grid, dim3 block, Mode mode, const std::function<void()>& body);
}  // namespace emu
static inline void __syncthreads() { emu::syncthreads(); }
static inline void __nanosleep(unsigned) {}
static inline float __shfl_up_sync(unsigned, float v, unsigned d) { return emu::shfl_up(v, d); }
static inline float __shfl_down_sync(unsigned, float v, unsigned d) { return emu::shfl_down(v, d); }
static inline int __shfl_sync(unsigned, int v, int lane) { return emu::shfl_idx(v, lane); }
static inline float __shfl_sync(unsigned, float v, int lane) { return __int_as_float(emu::shfl_idx(__float_as_int(v), lane)); }
static inline void __syncwarp() { (void)emu::shfl_idx(0, 0); }
static inline int __any_sync(unsigned, int pred) { int acc = 0; for (int l = 0; l < 32; ++l) acc |= emu::shfl_idx(pred ? 1 : 0, l); return acc; }   // full, converged warps only   // every live lane of the warp arrives before any continues

// ---- runtime API subset (device memory is host memory) ----
typedef void* cudaStream_t;
typedef void* cudaEvent_t;
enum cudaError_t { cudaSuccess = 0, cudaErrorUnknown = 999 };
enum cudaMemcpyKind { cudaMemcpyHostToHost, cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice, cudaMemcpyDefault };
static inline const char* cudaGetErrorString(cudaError_t) { return "emulator"; }
static inline cudaError_t cudaGetLastError() { return cudaSuccess; }
static inline cudaError_t cudaMalloc(void** p, size_t n) { *p = std::calloc(1, n ? n : 1); return *p ? cudaSuccess : cudaErrorUnknown; }
template <class T> static inline cudaError_t cudaMalloc(T** p, size_t n) { return cudaMalloc((void**)p, n); }
static inline cudaError_t cudaFree(void* p) { std::free(p); return cudaSuccess; }
static inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t = nullptr) { std::memmove(d, s, n); return cudaSuccess; }
static inline cudaError_t cudaMemcpy(void* d, const void* s, size_t n, cudaMemcpyKind) { std::memmove(d, s, n); return cudaSuccess; }
static inline cudaError_t cudaMemsetAsync(void* d, int v, size_t n, cudaStream_t = nullptr) { std::memset(d, v, n); return cudaSuccess; }
static inline cudaError_t cudaMemset(void* d, int v, size_t n) { std::memset(d, v, n); return cudaSuccess; }
static inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaDeviceSynchronize() { return cudaSuccess; }
static inline cudaError_t cudaSetDevice(int) { return cudaSuccess; }
static inline cudaError_t cudaGetDeviceCount(int* n) { *n = 1; return cudaSuccess; }
struct cudaDeviceProp { char name[256]; int multiProcessorCount; size_t totalGlobalMem; int major, minor; };
static inline cudaError_t cudaGetDeviceProperties(cudaDeviceProp* p, int) { std::memset(p, 0, sizeof(*p)); std::strcpy(p->name, "host emulation"); p->multiProcessorCount = 148; p->major = 10; return cudaSuccess; }
enum { cudaStreamNonBlocking = 1 };
static inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, unsigned) { *s = nullptr; return cudaSuccess; }
static inline cudaError_t cudaStreamDestroy(cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaMemcpy2DAsync(void* d, size_t dp, const void* s, size_t sp, size_t wbytes, size_t rows, cudaMemcpyKind, cudaStream_t = nullptr) {
  for (size_t r = 0; r < rows; ++r) std::memmove((char*)d + r * dp, (const char*)s + r * sp, wbytes);
  return cudaSuccess;
}
static inline cudaError_t cudaMemcpy2D(void* d, size_t dp, const void* s, size_t sp, size_t wbytes, size_t rows, cudaMemcpyKind k) { return cudaMemcpy2DAsync(d, dp, s, sp, wbytes, rows, k); }
// CUDA graphs are not modelled: the emulated context renders with plain launches (s360_ctx_set_graphs(ctx, 0))
typedef void* cudaGraph_t;
typedef void* cudaGraphExec_t;
enum cudaStreamCaptureMode { cudaStreamCaptureModeThreadLocal };
static inline cudaError_t cudaStreamBeginCapture(cudaStream_t, cudaStreamCaptureMode) { return cudaErrorUnknown; }
static inline cudaError_t cudaStreamEndCapture(cudaStream_t, cudaGraph_t*) { return cudaErrorUnknown; }
static inline cudaError_t cudaGraphInstantiate(cudaGraphExec_t*, cudaGraph_t, unsigned long long) { return cudaErrorUnknown; }
static inline cudaError_t cudaGraphLaunch(cudaGraphExec_t, cudaStream_t) { return cudaErrorUnknown; }
static inline cudaError_t cudaGraphDestroy(cudaGraph_t) { return cudaSuccess; }
static inline cudaError_t cudaGraphExecDestroy(cudaGraphExec_t) { return cudaSuccess; }
static inline cudaError_t cudaEventCreate(cudaEvent_t* e) { *e = nullptr; return cudaSuccess; }
static inline cudaError_t cudaEventDestroy(cudaEvent_t) { return cudaSuccess; }
static inline cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t = nullptr) { return cudaSuccess; }
static inline cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t, cudaEvent_t) { *ms = 0; return cudaSuccess; }
