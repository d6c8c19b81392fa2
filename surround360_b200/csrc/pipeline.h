// Context and per-frame orchestration (host C++). Mirrors renderStereoPanorama()
// (surround360_render/source/test/TestRenderStereoPanorama.cpp:716-972) stage by stage,
// enqueueing the sm_100a kernels on the context's stream. No CPU fallback anywhere.
#pragma once
#include <map>
#include <memory>
#include <string>
#include <vector>
#include "../../include/s360.h"
#include "kernels.h"
#include "rig.h"

namespace s360 {

struct DeviceBuffer {  // RAII cudaMalloc
  void* p = nullptr;
  size_t bytes = 0;
  DeviceBuffer() {}
  explicit DeviceBuffer(size_t n) { alloc(n); }
  ~DeviceBuffer() { release(); }
  DeviceBuffer(const DeviceBuffer&) = delete;
  DeviceBuffer& operator=(const DeviceBuffer&) = delete;
  DeviceBuffer(DeviceBuffer&& o) noexcept : p(o.p), bytes(o.bytes) { o.p = nullptr; o.bytes = 0; }
  void alloc(size_t n);
  void ensure(size_t n) { if (n > bytes) { release(); alloc(n); } }
  void release();
  template <class T> T* as() const { return reinterpret_cast<T*>(p); }
};

struct Geometry {  // TRSP:155-162, 309-316, 340-348, 607-610
  int camW, camH, overlapW, numNovel, poleH, extW;
  float hRadians, vRadians, fovHRad, slabDisp, shiftPx;
  // pole ramp constants, TRSP:455-481
  float poleCameraRadiusDeg, phiRampStart, phiMid, phiRampEnd;
};
Geometry computeGeometry(const Rig& rig, const s360_params& p);

struct FramePlan;  // all device buffers and tables for one (params, image size)

}  // namespace s360

struct s360_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  s360::Rig rig;
  bool rigLoaded = false;
  s360::DeviceBuffer tabI, tabF, tabT;  // cv::remap weight tables, libm tables
  s360::RemapTables rt{};
  std::map<int, std::unique_ptr<s360::DeviceBuffer>> blurKernels;  // erodeSize -> 8.8 fixed-point taps
  std::unique_ptr<s360::FramePlan> plan;
  s360::DeviceBuffer previewWarp;       // s360_hyper_preview: equirect -> camera warp of the three preview cameras
  int previewWarpW = 0, previewWarpH = 0;
  s360::DeviceBuffer poleMasks;         // two BGR8 masks (bottom, bottom2), s360_set_pole_masks
  int poleMaskW = 0, poleMaskH = 0;
  bool keepIntermediates = false;
  bool profile = false;
  bool profileKernels = false;          // s360_profile_enable(ctx, 2): events around every kernel launch (s360_profile_kernels)
  bool useGraphs = true;
  bool latencyMode = false;             // s360_ctx_set_latency_mode: the flow batches run the latency variant of the sweep
  bool temporal = false;
  s360::FlowBatch::SweepProfile sweepProfile;
  std::map<std::string, std::unique_ptr<s360::FlowBatch>> flowOps;   // s360_flow_compute_dev: one solver per (algorithm, size)
  s360::KernelProf kernelProf;          // per-kernel CUDA-event times + algorithmic bytes while `profile` is on
  float stageMs[7] = {0, 0, 0, 0, 0, 0, 0};
  std::string idScratch;

  const int* blurKernelFor(int erodeSize);
  s360_ctx();
  ~s360_ctx();
};

namespace s360 {
void frameOutputSize(s360_ctx* c, const s360_params& p, int* eqW, int* eqH, int* cubeW, int* cubeH);
// images: device BGR8 pointers (one per rig camera); outputs device packed BGR
void renderFrameDevice(s360_ctx* c, const s360_params& p, const uint8_t* const* imagesDev, int imgW, int imgH,
                       uint8_t* equirectDev, uint8_t* cubemapDev);
void renderFrameHost(s360_ctx* c, const s360_params& p, const uint8_t* const* images, int imgW, int imgH, uint8_t* equirect,
                     uint8_t* cubemap);
int frameIntermediate(s360_ctx* c, const char* name, int idx, void* out, int* w, int* h, int* ch, int* isFloat);
// video mode through files (TRSP:200-255, 413-452): plan without rendering, previous-frame state from host buffers
void prepareFrame(s360_ctx* c, const s360_params& p, int imgW, int imgH);
void temporalSizes(s360_ctx* c, int* overlapW, int* overlapH, int* extW, int* extH);
void temporalImportSide(s360_ctx* c, int pair, const float* flowLtoR, const float* flowRtoL, const uint8_t* overlapL, const uint8_t* overlapR);
void temporalImportPole(s360_ctx* c, int slot, const float* flow, const uint8_t* extFish);
void temporalCommit(s360_ctx* c);
void poleToSide(s360_ctx* c, const s360_params& p, const uint8_t* sideHost, int sw, int sh, const uint8_t* fishHost, int fw, int fh,
                uint8_t* warpedOut, float* flowOut);
std::vector<int> gaussKernelFixed8(int n, double sigma);
void collectProfile(s360_ctx* c);
void checkFlowErrors(s360_ctx* c);
}  // namespace s360
