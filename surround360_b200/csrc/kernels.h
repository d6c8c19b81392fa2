// Internal launcher API of the sm_100a kernels (host-callable, all asynchronous on `st`).
// Device images are BGRA8 (uchar4) with a pitch in pixels; flows are float2.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <atomic>
#include <map>
#include <string>
#include <vector>
#include "common.cuh"
#include "remap.cuh"

namespace s360 {

// global launch counter (bench.py "gpu_launches")
extern std::atomic<unsigned long long> g_launchCount;

// ---------------- warp_kernels.cu ----------------
// BGR8/BGRA8 host-layout source -> BGRA8 with alpha 255 (or source alpha) and the
// projectSideToSpherical ramp on the first/last `feather` rows (TRSP:110-125).
void launchPrepSource(cudaStream_t st, const uint8_t* src, int srcChannels, int w, int h, int feather, uchar4* dst);

struct SphericalTables {  // per-destination-column / per-row trig, made on the host with
  const float* cosX;      // cosf/sinf exactly as ImageWarper.cpp:154-163 evaluates them
  const float* sinX;
  const float* cosY;
  const float* sinY;
};
// bicubicRemapToSpherical. poleMode 0: keep remapped alpha (side cameras). poleMode 1:
// alpha = 255 then the std_alpha_feather ramp of TRSP:629-637 / 671-678 from row featherStart.
// poleMode 2 (bottom image with pole removal): remapped alpha, min() with the ramp (TRSP:629-637).
void launchProjectSpherical(cudaStream_t st, const uchar4* src, int sw, int sh, const DevCamera& cam,
                            const SphericalTables& tb, uchar4* dst, int dw, int dh, int dstPitch, int poleMode,
                            int featherStart, int featherSize, const RemapTables& rt);

struct NovelViewArgs {
  const uchar4* proj;     // all side projections, [numCams][camH][camW]
  const float2* flowLR;   // [numCams][camH][overlapW]
  const float2* flowRL;
  int numCams, camW, camH, overlapW, numNovel, bufW;
  float slabDisp, shiftPx;
  uchar4* panoL;          // eqrW x eqrH, rows [padAbove, padAbove+camH) written
  uchar4* panoR;
  int eqrW, padAbove;
};
// renderLazyNovelView x4 + combineLazyViews x2 + stackHorizontal + offsetHorizontalWrap
// (NovelView.cpp:101-268, TRSP:259-292,380-384) in one pass, both eyes.
void launchNovelView(cudaStream_t st, const NovelViewArgs& a, const RemapTables& rt);

// One camera pair, chunk-local output (C-ABI s360_novel_view_chunk).
void launchNovelViewChunk(cudaStream_t st, const uchar4* ovL, const uchar4* ovR, int pitch, const float2* flowLR,
                          const float2* flowRL, int w, int h, int bufW, int camW, int numNovel, float slabDisp,
                          uchar4* chunkL, uchar4* chunkR, const RemapTables& rt);

void launchOffsetWrap(cudaStream_t st, const uchar4* src, int w, int h, float offset, uchar4* dst);

struct PoleWarpArgs {  // TRSP:454-546
  const uchar4* extFish;  // extW x rows
  const float2* flow;     // extW x rows
  int cols, rows, extW;
  float poleCameraRadiusDeg, phiRampStart, phiMid, phiRampEnd;
  uchar4* warped;         // cols x rows
};
void launchPoleWarp(cudaStream_t st, const PoleWarpArgs& a, const RemapTables& rt);

// cv::resize INTER_CUBIC for 8U with `cn` (3 or 4) channel semantics on BGRA storage.
// If dstPacked3 != nullptr the result is written as packed BGR8 rows (pitch dstW*3 bytes).
void launchResizeCubic8(cudaStream_t st, const uchar4* src, int sw, int sh, int spitch, int cn, uchar4* dst, int dw, int dh,
                        int dpitch, uint8_t* dstPacked3);

// convertSphericalToCubemapBicubicRemap + stackOutputCubemapFaces for one eye, packed BGR out.
void launchCubemap(cudaStream_t st, const uchar4* eye, int w, int h, int faceW, int faceH, int photo, uint8_t* outPacked3,
                   const RemapTables& rt);

void launchPackBGR(cudaStream_t st, const uchar4* src, int w, int h, int pitch, uint8_t* dst);
void launchUnpackBGRA(cudaStream_t st, const uint8_t* src, int channels, int w, int h, uchar4* dst);  // alpha=255 if 3ch
// encode hand-off: BGR8 (row stride in bytes) -> NV12 planes; matrix 0 = BT.601, 1 = BT.709, limited range
void launchBgrToNv12(cudaStream_t st, const uint8_t* bgr, size_t bgrStride, int w, int h, int matrix, uint8_t* yPlane, size_t yPitch, uint8_t* uvPlane,
                     size_t uvPitch);
void launchPackBGRA(cudaStream_t st, const uchar4* src, int w, int h, int pitch, uint8_t* dst);

// ---------------- finish_kernels.cu ----------------
// featherAlphaChannel on a crop of `rows` rows, optionally read through flip(-1) of a
// srcH-row image, written into a 1.2x wrap-extended image (TRSP:396-411). scratch: 3*w*rows bytes + 2*w*rows u16
void launchFeatherExtend(cudaStream_t st, const uchar4* src, int w, int srcH, int rows, int flip, int erodeSize,
                         const int* blurKernelFixed, uchar4* extOut, int extW, uint8_t* scratch);
void launchExtendWrap(cudaStream_t st, const uchar4* src, int w, int rows, uchar4* extOut, int extW);
// plain featherAlphaChannel (C-ABI)
void launchFeatherAlpha(cudaStream_t st, const uchar4* src, int w, int h, int erodeSize, const int* blurKernelFixed,
                        uchar4* dst, uint8_t* scratch);
size_t featherScratchBytes(int w, int rows);

// flattenLayersDeghostPreferBase; top may be smaller (topRows) -> zero layer below it.
// Fused frame variant: base, then top layer, then bottom layer through flip(-1) (TRSP:864-885).
void launchFlatten(cudaStream_t st, const uchar4* base, const uchar4* top, int w, int h, uchar4* out, const RemapTables& rt);
void launchFlattenPoles(cudaStream_t st, const uchar4* base, int w, int h, const uchar4* topLayer, const uchar4* botLayer,
                        int poleRows, uchar4* out, const RemapTables& rt);

// sharpenThread: iirLowPass (wrap / reflect) + sharpenWithIirLowPass, in place on BGRA storage, for
// `eyes` images stored back to back. scratch: eyes*w*h*3 floats + 2*eyes*w*h uchar4
void launchSharpen(cudaStream_t st, uchar4* img, int w, int h, int eyes, float amount, float* scratchF, uchar4* scratchLp);

// pole removal (PoleRemoval.cpp:32-188). prepare: BGR + red mask -> BGRA with the usable-radius circle and the
// mask cut out of alpha (CvUtil.cpp:201-222). merge: warp image2 by the flow (cubic), alpha-weighted
// combination, new circle alpha (PoleRemoval.cpp:128-181).
void launchPoleRemovalPrepare(cudaStream_t st, const uint8_t* srcBGR, const uint8_t* maskBGR, int w, int h, float radius, uchar4* dst);
void launchFlipBoth(cudaStream_t st, const uchar4* src, int w, int h, uchar4* dst);
void launchPoleRemovalMerge(cudaStream_t st, const uchar4* img1, const uchar4* img2, const float2* flow, int w, int h, float radius,
                            uchar4* dst, const RemapTables& rt);

// ---------------- flow_kernels.cu ----------------
struct FlowParams {  // OpticalFlowFactory.h:25-61
  float pyrScaleFactor = 0.9f, smoothnessCoef = 0.001f, verticalRegularizationCoef = 0.01f,
        horizontalRegularizationCoef = 0.01f, gradientStepSize = 0.5f, downscaleFactor = 0.5f;
  int maxPercentage = 0;
};
bool flowParamsByName(const char* name, FlowParams* out);

constexpr int kMaxFlowBatch = 32;
struct FlowImagePtrs {  // per-flow source views (BGRA8), passed by value to kernels
  const uchar4* i0[kMaxFlowBatch];
  const uchar4* i1[kMaxFlowBatch];
  int pitch0[kMaxFlowBatch];
  int pitch1[kMaxFlowBatch];
  int hint[kMaxFlowBatch];
};
struct FlowPrevPtrs {  // optional temporal inputs
  const float2* prevFlow[kMaxFlowBatch];
  const uchar4* prevI1[kMaxFlowBatch];
  int prevPitch[kMaxFlowBatch];
  bool enabled = false;
};


// ------------------------------------------------------------------------------------
// Per-kernel profiling (bench.py's `roofline_all`): when a context enables it, every launch of the frame is bracketed by two
// CUDA events on the launching stream and tagged with the kernel's name and its ALGORITHMIC bytes (compulsory reads + writes
// of that launch). Off by default: the scope object is two pointer tests.
// ------------------------------------------------------------------------------------
struct KernelProf {
  struct Rec { const char* name; double bytes; cudaEvent_t a, b; };
  struct Acc { double ms = 0, bytes = 0; unsigned long long launches = 0; };
  bool on = false;
  std::vector<Rec> pending;
  std::vector<cudaEvent_t> pool;
  std::map<std::string, Acc> acc;
  cudaEvent_t get();
  void collect();   // after the stream has been synchronised
  void clear();
  ~KernelProf();
};
extern thread_local KernelProf* t_kernelProf;   // set by the pipeline for the duration of a frame's launches
struct ProfScope {
  cudaStream_t st;
  KernelProf::Rec rec;
  bool on;
  ProfScope(cudaStream_t s, const char* name, double bytes);
  ~ProfScope();
};
#define S360_PROF_CAT2(a, b) a##b
#define S360_PROF_CAT(a, b) S360_PROF_CAT2(a, b)
// usage: { S360_PROF(st, "k_name", bytes); k_name<<<...>>>(...); }
#define S360_PROF(st, name, bytes) ::s360::ProfScope S360_PROF_CAT(_prof_, __LINE__)(st, name, (double)(bytes))

struct Rig;
// TestHyperPreview (preview_kernels.cu)
void buildPreviewWarp(const Rig& rig, int eqrW, int eqrH, std::vector<float2>& out);
void launchPreview(cudaStream_t st, const uint16_t* const rawDev[3], int imgW, int imgH, float gamma, float4* layers, const float2* warpDev, int eqrW,
                   int eqrH, float softmaxCoef, const RemapTables& rt, float* outDev);

class FlowBatch {  // PixFlow::computeOpticalFlow for n same-sized pairs, solved concurrently
 public:
  // temporal: allocate the previous-flow / motion pyramids; levelDump: allocate per-level dump planes
  // externalWorkspace: do not allocate; the owner calls bindWorkspace() with >= bytes() of device memory that
  // may be shared with other batches which never run at the same time
  FlowBatch(int n, int w, int h, const FlowParams& p, bool temporal = false, bool levelDump = false, bool externalWorkspace = false);
  void bindWorkspace(void* ws, size_t bytes);
  ~FlowBatch();
  FlowBatch(const FlowBatch&) = delete;
  // flowOut: [n][h][w] float2 contiguous
  void run(cudaStream_t st, const FlowImagePtrs& imgs, const FlowPrevPtrs& prev, float2* flowOut);
  int n() const { return n_; }
  int w() const { return w_; }
  int h() const { return h_; }
  size_t bytes() const { return totalBytes_; }
  int numLevels() const { return (int)lw_.size(); }
  double levelPixels() const;  // sum over levels of w*h (one flow)
  // copy the flow as it stands after each level (debug / tests); enabled before run()
  void enableLevelDump(bool on) { dump_ = on; }
  const float2* levelDump(int level) const { return dumpBuf_[level]; }
  int levelW(int l) const { return lw_[l]; }
  int levelH(int l) const { return lh_[l]; }
  // sweep-kernel profiling: CUDA events around every k_sweep launch (on the launching stream)
  struct SweepProfile { double ms = 0, levelPixels = 0, steps = 0; unsigned long long launches = 0, fallbacks = 0; };
  void enableProfile(bool on) { profile_ = on; }
  // latency mode: the sweeps run the cp.async-ring kernel (k_sweep<.., 4>), which shortens a frame rendered alone and costs throughput
  // when many frames are in flight (flow_kernels.cu); results are bit-identical
  void setLatencyMode(bool on);
  // call after the stream has been synchronised; accumulates and clears pending events
  void collectProfile(SweepProfile* acc);
  // after the stream has been synchronised: throws if a sweep band gave up waiting for its producer (see k_sweep's bounded spin)
  void checkSweepError();

 private:
  int n_, w_, h_, dw_, dh_;
  FlowParams P_;
  std::vector<int> lw_, lh_;
  size_t totalBytes_ = 0;
  bool dump_ = false;
  bool profile_ = false;
  bool latency_ = false;
  bool temporal_ = false, dumpAlloc_ = false;
  struct ProfRec { cudaEvent_t a, b; double px, steps; };
  std::vector<ProfRec> pending_;
  std::vector<cudaEvent_t> evPool_;
  cudaEvent_t getEvent();
  void buildResizeTabs();
  std::vector<float2*> dumpBuf_;
  // device buffers (per level, batched over n)
  std::vector<float*> I0_, I1_, A0_, A1_;
  std::vector<float4*> rowRec_, quad_;
  std::vector<float2*> prevPyr_;
  std::vector<float*> D_, motionPyr_;
  uchar4* small0_ = nullptr;
  uchar4* small1_ = nullptr;
  float2 *full_ = nullptr, *flowA_ = nullptr, *flowB_ = nullptr, *tmp_ = nullptr, *boundary_ = nullptr;
  int* progress_ = nullptr;
  int* sweepErr_ = nullptr;
  // TMA descriptors (CUtensorMap, 128 bytes each, kept opaque here) of the two flow ping-pong planes per level, for the A/B
  // variant of the blur's row pass (env S360_TMA=1); tmOk_[l] = the level's row pitch satisfies TMA's 16-byte stride rule
  struct alignas(64) TensorMapBytes { unsigned char b[128]; };
  std::vector<TensorMapBytes> tmA_, tmB_;
  std::vector<char> tmOk_;
  bool useTma_ = false;
  void buildTensorMaps();
  // frame-invariant resize coordinates / weights per pyramid level, tabulated on the host once (flow_kernels.cu: ResizeTabs)
  void* tabAlloc_ = nullptr;
  std::vector<const int*> cubSx_, cubSy_;      // [l]: cubic upsample level l+1 -> l
  std::vector<const float4*> cubAx_, cubBy_;
  const int *downSx_ = nullptr, *downSy_ = nullptr;   // 8-bit cubic downscale full size -> working size
  const int4 *downAx_ = nullptr, *downBy_ = nullptr;
  const float4* downFby_ = nullptr;
  std::vector<const int*> linSx_, linSy_;      // [l]: linear pyramid level l-1 -> l; [0]: the final upscale to the full size
  std::vector<const float*> linFx_, linFy_;
  size_t boundaryElems_ = 0;
  void* arena_ = nullptr;
  void* boundaryAlloc_ = nullptr;
  bool bound_ = false;
  std::vector<size_t> oI0_, oI1_, oA0_, oA1_, oPrev_, oMot_, oDump_;
  size_t oS0_ = 0, oS1_ = 0, oRec_ = 0, oQuad_ = 0, oD_ = 0, oFA_ = 0, oFull_ = 0, oFB_ = 0, oTmp_ = 0;
};

}  // namespace s360
