// C ABI (include/s360.h). Every entry point converts exceptions into a status code and a
// thread-local message; nothing here falls back to the CPU.
#include <cstdio>
#include <cstring>
#include <fstream>
#include <sstream>
#include "pipeline.h"
#include "isp.h"
#include "footage.h"

using namespace s360;

namespace s360 { void initContextTables(s360_ctx* c); void resetPlan(s360_ctx* c); void resetTemporal(s360_ctx* c); }

static thread_local std::string g_lastError;
#define S360_API_BEGIN try {
#define S360_API_END                                             \
  return 0;                                                      \
  } catch (const std::exception& e) { g_lastError = e.what(); return 1; } \
  catch (...) { g_lastError = "unknown error"; return 1; }

static void require(bool ok, const char* msg) { if (!ok) throw std::runtime_error(msg); }
static void useDevice(s360_ctx* c) { require(c != nullptr, "null context"); S360_CUDA_CHECK(cudaSetDevice(c->device)); }

extern "C" {

const char* s360_last_error(void) { return g_lastError.c_str(); }
const char* s360_version(void) { return "surround360_b200 0.1 (sm_100a)"; }

void s360_params_default(s360_params* p) {
  std::memset(p, 0, sizeof(*p));
  p->interpupilary_dist = 6.4; p->sharpening = 0.0; p->zero_parallax_dist = 10000.0;
  p->side_alpha_feather_size = 100; p->std_alpha_feather_size = 31;
  p->eqr_width = 256; p->eqr_height = 128; p->final_eqr_width = 3480; p->final_eqr_height = 960;
  p->cubemap_width = 1536; p->cubemap_height = 1536;
  std::strcpy(p->side_flow_alg, "pixflow_low");
  std::strcpy(p->polar_flow_alg, "pixflow_low");
  std::strcpy(p->poleremoval_flow_alg, "pixflow_low");
}

int s360_ctx_create(int device, s360_ctx** out) {
  S360_API_BEGIN
  require(out != nullptr, "null out pointer");
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0) throw std::runtime_error("no CUDA device available: this library has no CPU fallback");
  require(device >= 0 && device < count, "device index out of range");
  S360_CUDA_CHECK(cudaSetDevice(device));
  cudaDeviceProp prop;
  S360_CUDA_CHECK(cudaGetDeviceProperties(&prop, device));
  if (prop.major < 10) throw std::runtime_error(std::string("device '") + prop.name + "' is not sm_100a-class; kernels are built for B200 only");
  auto c = std::make_unique<s360_ctx>();
  c->device = device;
  S360_CUDA_CHECK(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
  initContextTables(c.get());
  *out = c.release();
  S360_API_END
}
void s360_ctx_destroy(s360_ctx* c) {
  if (!c) return;
  cudaSetDevice(c->device);
  cudaStreamSynchronize(c->stream);
  resetPlan(c);
  if (c->stream) cudaStreamDestroy(c->stream);
  delete c;
}
int s360_ctx_synchronize(s360_ctx* c) {
  S360_API_BEGIN
  useDevice(c);
  S360_CUDA_CHECK(cudaStreamSynchronize(c->stream));
  checkFlowErrors(c);
  S360_API_END
}
void* s360_ctx_stream(s360_ctx* c) { return c ? (void*)c->stream : nullptr; }
unsigned long long s360_ctx_launch_count(s360_ctx* c) { (void)c; return g_launchCount.load(); }
int s360_ctx_set_keep_intermediates(s360_ctx* c, int on) {
  S360_API_BEGIN
  require(c != nullptr, "null context");
  c->keepIntermediates = on != 0;
  S360_API_END
}

int s360_ctx_set_temporal(s360_ctx* c, int on) {
  S360_API_BEGIN
  require(c != nullptr, "null context");
  c->temporal = on != 0;
  resetTemporal(c);
  S360_API_END
}
int s360_ctx_reset_temporal(s360_ctx* c) {
  S360_API_BEGIN
  require(c != nullptr, "null context");
  resetTemporal(c);
  S360_API_END
}
int s360_frame_prepare(s360_ctx* c, const s360_params* p, int w, int h) {
  S360_API_BEGIN
  require(c != nullptr && p != nullptr, "null argument");
  require(w > 0 && h > 0, "bad image size");
  useDevice(c);
  prepareFrame(c, *p, w, h);
  S360_API_END
}
int s360_temporal_sizes(s360_ctx* c, int* ow, int* oh, int* ew, int* eh) {
  S360_API_BEGIN
  require(c != nullptr && ow && oh && ew && eh, "null argument");
  temporalSizes(c, ow, oh, ew, eh);
  S360_API_END
}
int s360_temporal_import_side(s360_ctx* c, int pair, const float* flr, const float* frl, const uint8_t* ol, const uint8_t* orr) {
  S360_API_BEGIN
  require(c != nullptr && flr && frl && ol && orr, "null argument");
  useDevice(c);
  temporalImportSide(c, pair, flr, frl, ol, orr);
  S360_API_END
}
int s360_temporal_import_pole(s360_ctx* c, int slot, const float* flow, const uint8_t* extFish) {
  S360_API_BEGIN
  require(c != nullptr && flow && extFish, "null argument");
  useDevice(c);
  temporalImportPole(c, slot, flow, extFish);
  S360_API_END
}
int s360_temporal_commit(s360_ctx* c) {
  S360_API_BEGIN
  require(c != nullptr, "null context");
  temporalCommit(c);
  S360_API_END
}
int s360_ctx_set_graphs(s360_ctx* c, int on) {
  S360_API_BEGIN
  require(c != nullptr, "null context");
  c->useGraphs = on != 0;
  S360_API_END
}
int s360_ctx_set_latency_mode(s360_ctx* c, int on) {
  S360_API_BEGIN
  require(c != nullptr, "null context");
  if (c->latencyMode != (on != 0)) {
    c->latencyMode = on != 0;
    resetPlan(c);   // the frame plan's flow batches and its captured graph hold the kernel choice
  }
  S360_API_END
}
int s360_profile_enable(s360_ctx* c, int on) {
  S360_API_BEGIN
  require(c != nullptr, "null context");
  c->profile = on != 0;
  c->profileKernels = on >= 2;
  c->sweepProfile = FlowBatch::SweepProfile();
  c->kernelProf.clear();
  S360_API_END
}
int s360_profile_kernels(s360_ctx* c, char* json, size_t cap, size_t* needed) {
  S360_API_BEGIN
  useDevice(c);
  require(c != nullptr && needed != nullptr, "null argument");
  collectProfile(c);
  std::string out = "[";
  bool first = true;
  for (const auto& kv : c->kernelProf.acc) {
    char buf[512];
    std::snprintf(buf, sizeof buf, "%s{\"kernel\": \"%s\", \"ms\": %.6f, \"algorithmic_bytes\": %.0f, \"launches\": %llu}", first ? "" : ", ",
                  kv.first.c_str(), kv.second.ms, kv.second.bytes, kv.second.launches);
    out += buf;
    first = false;
  }
  out += "]";
  *needed = out.size() + 1;
  if (json && cap >= out.size() + 1) std::memcpy(json, out.c_str(), out.size() + 1);
  S360_API_END
}
int s360_profile_sweep(s360_ctx* c, double* ms, unsigned long long* launches, double* levelPixels, double* steps) {
  S360_API_BEGIN
  useDevice(c);
  collectProfile(c);
  if (ms) *ms = c->sweepProfile.ms;
  if (launches) *launches = c->sweepProfile.launches;
  if (levelPixels) *levelPixels = c->sweepProfile.levelPixels;
  if (steps) *steps = c->sweepProfile.steps;
  S360_API_END
}
int s360_profile_sweep_fallbacks(s360_ctx* c, unsigned long long* fallbacks) {
  S360_API_BEGIN
  require(c != nullptr && fallbacks != nullptr, "null argument");
  *fallbacks = c->sweepProfile.fallbacks;   // kept for ABI stability: always 0 since the fast/generic sweep pair was removed
  S360_API_END
}

// ---- rig ----
int s360_rig_load_json(s360_ctx* c, const char* text) {
  S360_API_BEGIN
  require(c && text, "null argument");
  Rig r;
  r.parse(text);
  c->rig = r;
  c->rigLoaded = true;
  resetPlan(c);
  S360_API_END
}
int s360_rig_load_file(s360_ctx* c, const char* path) {
  S360_API_BEGIN
  require(c && path, "null argument");
  std::ifstream f(path);
  if (!f) throw std::runtime_error(std::string("could not read JSON file: ") + path);
  std::stringstream ss;
  ss << f.rdbuf();
  Rig r;
  r.parse(ss.str());
  c->rig = r;
  c->rigLoaded = true;
  resetPlan(c);
  S360_API_END
}
int s360_rig_num_cameras(s360_ctx* c) { return c && c->rigLoaded ? (int)c->rig.cameras.size() : -1; }
int s360_rig_num_side(s360_ctx* c) { return c && c->rigLoaded ? (int)c->rig.side.size() : -1; }
int s360_rig_special_camera(s360_ctx* c, int which) {
  try {
    require(c && c->rigLoaded, "no rig loaded");
    return which == 0 ? c->rig.topCamera() : which == 1 ? c->rig.bottomCamera() : c->rig.findLargestDistCamAxisToRigCenter();
  } catch (const std::exception& e) { g_lastError = e.what(); return -1; }
}
int s360_rig_side_index(s360_ctx* c, int i) {
  if (!c || !c->rigLoaded || i < 0 || i >= (int)c->rig.side.size()) return -1;
  return c->rig.side[i];
}
const char* s360_rig_camera_id(s360_ctx* c, int cam) {
  if (!c || !c->rigLoaded || cam < 0 || cam >= (int)c->rig.cameras.size()) return nullptr;
  return c->rig.cameras[cam].id.c_str();
}
int s360_camera_pixel(s360_ctx* c, int cam, const double xyz[3], double out[2]) {
  S360_API_BEGIN
  require(c && c->rigLoaded, "no rig loaded");
  require(cam >= 0 && cam < (int)c->rig.cameras.size(), "camera index out of range");
  c->rig.cameras[cam].pixel({xyz[0], xyz[1], xyz[2]}, out[0], out[1]);
  S360_API_END
}
int s360_rig_serialize(s360_ctx* c, void* buf, size_t cap, size_t* used) {
  S360_API_BEGIN
  require(c && c->rigLoaded, "no rig loaded");
  std::vector<uint8_t> b = c->rig.serialize();
  if (used) *used = b.size();
  if (buf) { require(cap >= b.size(), "buffer too small"); std::memcpy(buf, b.data(), b.size()); }
  S360_API_END
}
int s360_rig_deserialize(s360_ctx* c, const void* buf, size_t len) {
  S360_API_BEGIN
  require(c && buf, "null argument");
  Rig r;
  r.deserialize((const uint8_t*)buf, len);
  c->rig = r;
  c->rigLoaded = true;
  resetPlan(c);
  S360_API_END
}
int s360_set_pole_masks(s360_ctx* c, const uint8_t* m1, const uint8_t* m2, int w, int h) {
  S360_API_BEGIN
  useDevice(c);
  require(m1 && m2 && w > 0 && h > 0, "missing or bad pole mask");
  const size_t bytes = (size_t)w * h * 3;
  c->poleMasks.ensure(bytes * 2);
  S360_CUDA_CHECK(cudaMemcpy(c->poleMasks.p, m1, bytes, cudaMemcpyHostToDevice));
  S360_CUDA_CHECK(cudaMemcpy(c->poleMasks.as<uint8_t>() + bytes, m2, bytes, cudaMemcpyHostToDevice));
  c->poleMaskW = w; c->poleMaskH = h;
  resetPlan(c);   // the plan (and its captured graph) holds the mask pointers and sizes
  S360_API_END
}
int s360_geometry_query(s360_ctx* c, const s360_params* p, s360_geometry* out) {
  S360_API_BEGIN
  require(c && c->rigLoaded && p && out, "no rig loaded / null argument");
  const Geometry g = computeGeometry(c->rig, *p);
  out->cam_w = g.camW; out->cam_h = g.camH; out->overlap_w = g.overlapW; out->num_novel = g.numNovel;
  out->h_radians = g.hRadians; out->v_radians = g.vRadians; out->fov_h_radians = g.fovHRad; out->slab_disp = g.slabDisp;
  out->shift_px = g.shiftPx; out->pole_h = g.poleH;
  S360_API_END
}

// ---- operators ----
int s360_project_spherical(s360_ctx* c, int cam, const uint8_t* src, int sw, int sh, int sc, float l, float r, float t, float b, int feather,
                           uint8_t* dst, int dw, int dh) {
  S360_API_BEGIN
  useDevice(c);
  require(c->rigLoaded, "no rig loaded");
  require(cam >= 0 && cam < (int)c->rig.cameras.size(), "camera index out of range");
  require(src && dst && (sc == 3 || sc == 4) && sw > 0 && sh > 0 && dw > 0 && dh > 0, "bad image arguments");
  cudaStream_t st = c->stream;
  DeviceBuffer dRaw((size_t)sw * sh * sc), dSrc((size_t)sw * sh * 4), dDst((size_t)dw * dh * 4), dTab((size_t)(2 * dw + 2 * dh) * 4);
  std::vector<float> tab((size_t)2 * dw + 2 * dh);
  for (int x = 0; x < dw; ++x) {
    const float xFrac = (x + 0.5f) / dw;
    const float xAngle = (1 - xFrac) * l + xFrac * r;
    tab[x] = cosf(xAngle); tab[dw + x] = sinf(xAngle);
  }
  for (int y = 0; y < dh; ++y) {
    const float yFrac = (y + 0.5f) / dh;
    const float yAngle = (1 - yFrac) * t + yFrac * b;
    tab[2 * dw + y] = cosf(yAngle); tab[2 * dw + dh + y] = sinf(yAngle);
  }
  S360_CUDA_CHECK(cudaMemcpyAsync(dRaw.p, src, dRaw.bytes, cudaMemcpyHostToDevice, st));
  S360_CUDA_CHECK(cudaMemcpyAsync(dTab.p, tab.data(), tab.size() * 4, cudaMemcpyHostToDevice, st));
  launchPrepSource(st, dRaw.as<uint8_t>(), sc, sw, sh, feather, dSrc.as<uchar4>());
  SphericalTables tb{dTab.as<float>(), dTab.as<float>() + dw, dTab.as<float>() + 2 * dw, dTab.as<float>() + 2 * dw + dh};
  launchProjectSpherical(st, dSrc.as<uchar4>(), sw, sh, c->rig.cameras[cam].toDevice(), tb, dDst.as<uchar4>(), dw, dh, dw, 0, 0, 1, c->rt);
  S360_CUDA_CHECK(cudaMemcpyAsync(dst, dDst.p, (size_t)dw * dh * 4, cudaMemcpyDeviceToHost, st));
  S360_CUDA_CHECK(cudaStreamSynchronize(st));
  S360_API_END
}

int s360_flow_compute_batch(s360_ctx* c, const char* alg, int n, const uint8_t* const* i0, const uint8_t* const* i1, int w, int h,
                            float* const* flowOut, const s360_hint* hints) {
  S360_API_BEGIN
  useDevice(c);
  FlowParams fp;
  if (!flowParamsByName(alg, &fp)) throw std::runtime_error(std::string("unrecognized flow algorithm name: ") + (alg ? alg : "(null)"));
  require(n >= 1 && n <= kMaxFlowBatch && w > 0 && h > 0 && i0 && i1 && flowOut, "bad flow arguments");
  cudaStream_t st = c->stream;
  const size_t px = (size_t)w * h;
  DeviceBuffer d0(px * 4 * n), d1(px * 4 * n), dF(px * 8 * n);
  FlowImagePtrs ip{};
  FlowPrevPtrs pv{};
  for (int k = 0; k < n; ++k) {
    S360_CUDA_CHECK(cudaMemcpyAsync(d0.as<uchar4>() + k * px, i0[k], px * 4, cudaMemcpyHostToDevice, st));
    S360_CUDA_CHECK(cudaMemcpyAsync(d1.as<uchar4>() + k * px, i1[k], px * 4, cudaMemcpyHostToDevice, st));
    ip.i0[k] = d0.as<uchar4>() + k * px; ip.i1[k] = d1.as<uchar4>() + k * px; ip.pitch0[k] = ip.pitch1[k] = w;
    ip.hint[k] = hints ? (int)hints[k] : 0;
  }
  FlowBatch fb(n, w, h, fp);
  fb.setLatencyMode(c->latencyMode);
  fb.run(st, ip, pv, dF.as<float2>());
  for (int k = 0; k < n; ++k) S360_CUDA_CHECK(cudaMemcpyAsync(flowOut[k], dF.as<float2>() + k * px, px * 8, cudaMemcpyDeviceToHost, st));
  S360_CUDA_CHECK(cudaStreamSynchronize(st));
  fb.checkSweepError();
  S360_API_END
}

int s360_flow_compute(s360_ctx* c, const char* alg, const uint8_t* i0, const uint8_t* i1, int w, int h, const float* prevFlow,
                      const uint8_t* prevI0, const uint8_t* prevI1, float* flowOut, s360_hint hint) {
  S360_API_BEGIN
  useDevice(c);
  FlowParams fp;
  if (!flowParamsByName(alg, &fp)) throw std::runtime_error(std::string("unrecognized flow algorithm name: ") + (alg ? alg : "(null)"));
  require(w > 0 && h > 0 && i0 && i1 && flowOut, "bad flow arguments");
  (void)prevI0;  // resized by the reference but never read (PixFlow.h:106, 109-117 use prevI1 only)
  cudaStream_t st = c->stream;
  const size_t px = (size_t)w * h;
  DeviceBuffer d0(px * 4), d1(px * 4), dF(px * 8), dPF, dP1;
  S360_CUDA_CHECK(cudaMemcpyAsync(d0.p, i0, px * 4, cudaMemcpyHostToDevice, st));
  S360_CUDA_CHECK(cudaMemcpyAsync(d1.p, i1, px * 4, cudaMemcpyHostToDevice, st));
  FlowImagePtrs ip{};
  FlowPrevPtrs pv{};
  ip.i0[0] = d0.as<uchar4>(); ip.i1[0] = d1.as<uchar4>(); ip.pitch0[0] = ip.pitch1[0] = w; ip.hint[0] = (int)hint;
  if (prevFlow) {
    require(prevI1 != nullptr, "prev_flow given without prev images");
    dPF.alloc(px * 8); dP1.alloc(px * 4);
    S360_CUDA_CHECK(cudaMemcpyAsync(dPF.p, prevFlow, px * 8, cudaMemcpyHostToDevice, st));
    S360_CUDA_CHECK(cudaMemcpyAsync(dP1.p, prevI1, px * 4, cudaMemcpyHostToDevice, st));
    pv.enabled = true; pv.prevFlow[0] = dPF.as<float2>(); pv.prevI1[0] = dP1.as<uchar4>(); pv.prevPitch[0] = w;
  }
  FlowBatch fb(1, w, h, fp, pv.enabled);
  fb.setLatencyMode(c->latencyMode);
  fb.run(st, ip, pv, dF.as<float2>());
  S360_CUDA_CHECK(cudaMemcpyAsync(flowOut, dF.p, px * 8, cudaMemcpyDeviceToHost, st));
  S360_CUDA_CHECK(cudaStreamSynchronize(st));
  fb.checkSweepError();
  S360_API_END
}

// ---- stride-aware and device-pointer variants (cv::Mat views are not contiguous; the reference crops its overlap
// images as ROI views, TRSP:196-198) ----
static void copyIn2D(void* dst, size_t dstRowBytes, const void* src, size_t srcStrideBytes, size_t rowBytes, int rows, cudaStream_t st) {
  S360_CUDA_CHECK(cudaMemcpy2DAsync(dst, dstRowBytes, src, srcStrideBytes ? srcStrideBytes : rowBytes, rowBytes, rows, cudaMemcpyHostToDevice, st));
}
int s360_flow_compute_strided(s360_ctx* c, const char* alg, const uint8_t* i0, size_t i0Stride, const uint8_t* i1, size_t i1Stride, int w, int h,
                              const float* prevFlow, size_t prevFlowStride, const uint8_t* prevI1, size_t prevI1Stride, float* flowOut,
                              size_t flowOutStride, s360_hint hint) {
  S360_API_BEGIN
  useDevice(c);
  FlowParams fp;
  if (!flowParamsByName(alg, &fp)) throw std::runtime_error(std::string("unrecognized flow algorithm name: ") + (alg ? alg : "(null)"));
  require(w > 0 && h > 0 && i0 && i1 && flowOut, "bad flow arguments");
  const size_t rowB = (size_t)w * 4, rowF = (size_t)w * 8;
  require((!i0Stride || i0Stride >= rowB) && (!i1Stride || i1Stride >= rowB) && (!flowOutStride || flowOutStride >= rowF), "stride smaller than a row");
  cudaStream_t st = c->stream;
  const size_t px = (size_t)w * h;
  DeviceBuffer d0(px * 4), d1(px * 4), dF(px * 8), dPF, dP1;
  copyIn2D(d0.p, rowB, i0, i0Stride, rowB, h, st);
  copyIn2D(d1.p, rowB, i1, i1Stride, rowB, h, st);
  FlowImagePtrs ip{};
  FlowPrevPtrs pv{};
  ip.i0[0] = d0.as<uchar4>(); ip.i1[0] = d1.as<uchar4>(); ip.pitch0[0] = ip.pitch1[0] = w; ip.hint[0] = (int)hint;
  if (prevFlow) {
    require(prevI1 != nullptr, "prev_flow given without prev images");
    dPF.alloc(px * 8); dP1.alloc(px * 4);
    copyIn2D(dPF.p, rowF, prevFlow, prevFlowStride, rowF, h, st);
    copyIn2D(dP1.p, rowB, prevI1, prevI1Stride, rowB, h, st);
    pv.enabled = true; pv.prevFlow[0] = dPF.as<float2>(); pv.prevI1[0] = dP1.as<uchar4>(); pv.prevPitch[0] = w;
  }
  FlowBatch fb(1, w, h, fp, pv.enabled);
  fb.setLatencyMode(c->latencyMode);
  fb.run(st, ip, pv, dF.as<float2>());
  S360_CUDA_CHECK(cudaMemcpy2DAsync(flowOut, flowOutStride ? flowOutStride : rowF, dF.p, rowF, rowF, h, cudaMemcpyDeviceToHost, st));
  S360_CUDA_CHECK(cudaStreamSynchronize(st));
  fb.checkSweepError();
  S360_API_END
}
// Device pointers: images BGRA8 with a row stride in bytes (multiple of 4), flow output dense w*h float2. Enqueued on the context
// stream; the solver object for (algorithm, size) is created on first use (that first call allocates and synchronises) and kept.
int s360_flow_compute_dev(s360_ctx* c, const char* alg, const uint8_t* i0Dev, size_t i0Stride, const uint8_t* i1Dev, size_t i1Stride, int w, int h,
                          float* flowOutDev, s360_hint hint) {
  S360_API_BEGIN
  useDevice(c);
  FlowParams fp;
  if (!flowParamsByName(alg, &fp)) throw std::runtime_error(std::string("unrecognized flow algorithm name: ") + (alg ? alg : "(null)"));
  require(w > 0 && h > 0 && i0Dev && i1Dev && flowOutDev, "bad flow arguments");
  if (!i0Stride) i0Stride = (size_t)w * 4;
  if (!i1Stride) i1Stride = (size_t)w * 4;
  require(i0Stride % 4 == 0 && i1Stride % 4 == 0 && i0Stride >= (size_t)w * 4 && i1Stride >= (size_t)w * 4, "row strides must be multiples of 4 bytes and at least a row");
  const std::string key = std::string(alg) + ":" + std::to_string(w) + "x" + std::to_string(h);
  auto& slot = c->flowOps[key];
  if (!slot) slot = std::make_unique<FlowBatch>(1, w, h, fp);
  slot->setLatencyMode(c->latencyMode);
  FlowImagePtrs ip{};
  FlowPrevPtrs pv{};
  ip.i0[0] = reinterpret_cast<const uchar4*>(i0Dev); ip.i1[0] = reinterpret_cast<const uchar4*>(i1Dev);
  ip.pitch0[0] = (int)(i0Stride / 4); ip.pitch1[0] = (int)(i1Stride / 4); ip.hint[0] = (int)hint;
  slot->run(c->stream, ip, pv, reinterpret_cast<float2*>(flowOutDev));
  S360_API_END
}
int s360_novel_view_chunk_dev(s360_ctx* c, const uint8_t* ovLDev, const uint8_t* ovRDev, size_t ovStride, const float* fLRDev, const float* fRLDev, int w, int h,
                              int bufW, int camW, int numNovel, float slabDisp, uint8_t* chunkLDev, uint8_t* chunkRDev) {
  S360_API_BEGIN
  useDevice(c);
  require(ovLDev && ovRDev && fLRDev && fRLDev && chunkLDev && chunkRDev && w > 0 && h > 0 && bufW > 0, "bad novel view arguments");
  if (!ovStride) ovStride = (size_t)w * 4;
  require(ovStride % 4 == 0 && ovStride >= (size_t)w * 4, "row stride must be a multiple of 4 bytes and at least a row");
  launchNovelViewChunk(c->stream, reinterpret_cast<const uchar4*>(ovLDev), reinterpret_cast<const uchar4*>(ovRDev), (int)(ovStride / 4),
                       reinterpret_cast<const float2*>(fLRDev), reinterpret_cast<const float2*>(fRLDev), w, h, bufW, camW, numNovel, slabDisp,
                       reinterpret_cast<uchar4*>(chunkLDev), reinterpret_cast<uchar4*>(chunkRDev), c->rt);
  S360_API_END
}
int s360_project_spherical_strided(s360_ctx* c, int cam, const uint8_t* src, size_t srcStride, int sw, int sh, int sc, float l, float r, float t, float b,
                                   int feather, uint8_t* dst, size_t dstStride, int dw, int dh) {
  S360_API_BEGIN
  require(src && dst && (sc == 3 || sc == 4) && sw > 0 && sh > 0 && dw > 0 && dh > 0, "bad image arguments");
  const size_t srcRow = (size_t)sw * sc, dstRow = (size_t)dw * 4;
  require((!srcStride || srcStride >= srcRow) && (!dstStride || dstStride >= dstRow), "stride smaller than a row");
  // gather the rows, run the dense operator, scatter the rows (host side: the operator's copies dominate either way)
  std::vector<uint8_t> s((size_t)sw * sh * sc), d((size_t)dw * dh * 4);
  for (int y = 0; y < sh; ++y) std::memcpy(s.data() + y * srcRow, src + y * (srcStride ? srcStride : srcRow), srcRow);
  if (s360_project_spherical(c, cam, s.data(), sw, sh, sc, l, r, t, b, feather, d.data(), dw, dh)) throw std::runtime_error(s360_last_error());
  for (int y = 0; y < dh; ++y) std::memcpy(dst + y * (dstStride ? dstStride : dstRow), d.data() + y * dstRow, dstRow);
  S360_API_END
}

int s360_novel_view_chunk(s360_ctx* c, const uint8_t* ovL, const uint8_t* ovR, const float* fLR, const float* fRL, int w, int h, int bufW,
                          int camW, int numNovel, float slabDisp, uint8_t* chunkL, uint8_t* chunkR) {
  S360_API_BEGIN
  useDevice(c);
  require(ovL && ovR && fLR && fRL && chunkL && chunkR && w > 0 && h > 0 && bufW > 0, "bad novel view arguments");
  cudaStream_t st = c->stream;
  const size_t px = (size_t)w * h, opx = (size_t)bufW * h;
  DeviceBuffer dL(px * 4), dR(px * 4), dFLR(px * 8), dFRL(px * 8), dOL(opx * 4), dOR(opx * 4);
  S360_CUDA_CHECK(cudaMemcpyAsync(dL.p, ovL, px * 4, cudaMemcpyHostToDevice, st));
  S360_CUDA_CHECK(cudaMemcpyAsync(dR.p, ovR, px * 4, cudaMemcpyHostToDevice, st));
  S360_CUDA_CHECK(cudaMemcpyAsync(dFLR.p, fLR, px * 8, cudaMemcpyHostToDevice, st));
  S360_CUDA_CHECK(cudaMemcpyAsync(dFRL.p, fRL, px * 8, cudaMemcpyHostToDevice, st));
  launchNovelViewChunk(st, dL.as<uchar4>(), dR.as<uchar4>(), w, dFLR.as<float2>(), dFRL.as<float2>(), w, h, bufW, camW, numNovel, slabDisp,
                       dOL.as<uchar4>(), dOR.as<uchar4>(), c->rt);
  S360_CUDA_CHECK(cudaMemcpyAsync(chunkL, dOL.p, opx * 4, cudaMemcpyDeviceToHost, st));
  S360_CUDA_CHECK(cudaMemcpyAsync(chunkR, dOR.p, opx * 4, cudaMemcpyDeviceToHost, st));
  S360_CUDA_CHECK(cudaStreamSynchronize(st));
  S360_API_END
}

// TestHyperPreview (test/TestHyperPreview.cpp:77-185): raw Bayer frames of the top / bottom / secondary-bottom cameras -> mono equirect
int s360_hyper_preview(s360_ctx* c, const uint16_t* rawTop, const uint16_t* rawBottom, const uint16_t* rawBottom2, int imgW, int imgH, int eqrW, int eqrH,
                       double gamma, double softmaxCoef, float* outBgr) {
  S360_API_BEGIN
  useDevice(c);
  require(c->rigLoaded, "no rig loaded");
  require(rawTop && rawBottom && rawBottom2 && outBgr && imgW > 1 && imgH > 1 && eqrW > 0 && eqrH > 0, "bad preview arguments");
  cudaStream_t st = c->stream;
  if (c->previewWarpW != eqrW || c->previewWarpH != eqrH || !c->previewWarp.p) {   // the precomputed projection: once per (rig, size)
    std::vector<float2> warp;
    buildPreviewWarp(c->rig, eqrW, eqrH, warp);
    c->previewWarp.ensure(warp.size() * sizeof(float2));
    S360_CUDA_CHECK(cudaMemcpy(c->previewWarp.p, warp.data(), warp.size() * sizeof(float2), cudaMemcpyHostToDevice));
    c->previewWarpW = eqrW; c->previewWarpH = eqrH;
  }
  const size_t rawBytes = (size_t)imgW * imgH * 2, layerPx = (size_t)(imgW / 2) * (imgH / 2);
  DeviceBuffer dRaw(rawBytes * 3), dLayers(layerPx * 16 * 3), dOut((size_t)eqrW * eqrH * 12);
  const uint16_t* raws[3] = {rawTop, rawBottom, rawBottom2};
  const uint16_t* rawDev[3];
  for (int i = 0; i < 3; ++i) {
    S360_CUDA_CHECK(cudaMemcpyAsync(dRaw.as<uint8_t>() + i * rawBytes, raws[i], rawBytes, cudaMemcpyHostToDevice, st));
    rawDev[i] = reinterpret_cast<const uint16_t*>(dRaw.as<uint8_t>() + i * rawBytes);
  }
  launchPreview(st, rawDev, imgW, imgH, (float)gamma, dLayers.as<float4>(), c->previewWarp.as<float2>(), eqrW, eqrH, (float)softmaxCoef, c->rt, dOut.as<float>());
  S360_CUDA_CHECK(cudaMemcpyAsync(outBgr, dOut.p, (size_t)eqrW * eqrH * 12, cudaMemcpyDeviceToHost, st));
  S360_CUDA_CHECK(cudaStreamSynchronize(st));
  S360_API_END
}

// ---- soft ISP (isp.h) ----
namespace {
struct ProfActivate {   // events around the launches made by this thread while s360_profile_enable(ctx, 2) is on
  explicit ProfActivate(s360_ctx* c) { c->kernelProf.on = c->profileKernels; s360::t_kernelProf = c->profileKernels ? &c->kernelProf : nullptr; }
  ~ProfActivate() { s360::t_kernelProf = nullptr; }
};
}  // namespace
struct s360_isp { s360::IspPipeline pipe; s360_isp(const char* j, int bpp) : pipe(j, bpp) {} };
int s360_isp_create(const char* ispJson, int outputBpp, s360_isp** out) {
  S360_API_BEGIN
  require(ispJson && out, "null argument");
  *out = new s360_isp(ispJson, outputBpp);
  S360_API_END
}
void s360_isp_destroy(s360_isp* isp) { delete isp; }
int s360_isp_set_bits_per_pixel(s360_isp* isp, int bits) { S360_API_BEGIN require(isp != nullptr, "null isp"); isp->pipe.setBitsPerPixel(bits); S360_API_END }
int s360_isp_set_demosaic_filter(s360_isp* isp, int f) { S360_API_BEGIN require(isp != nullptr, "null isp"); isp->pipe.setDemosaicFilter(f); S360_API_END }
int s360_isp_set_resize(s360_isp* isp, int r) { S360_API_BEGIN require(isp != nullptr, "null isp"); isp->pipe.setResize(r); S360_API_END }
int s360_isp_set_tone_map(s360_isp* isp, int enable) {
  S360_API_BEGIN
  require(isp != nullptr, "null isp");
  if (enable) isp->pipe.enableToneMap(); else isp->pipe.disableToneMap();
  S360_API_END
}
int s360_isp_add_black_level_offset(s360_isp* isp, int o) { S360_API_BEGIN require(isp != nullptr, "null isp"); isp->pipe.addBlackLevelOffset(o); S360_API_END }
int s360_isp_tables(s360_isp* isp, int w, int h, float* ccm9, float* lut, float* vigH, float* vigV) {
  S360_API_BEGIN
  require(isp != nullptr, "null isp");
  isp->pipe.tables(w, h, ccm9, lut, vigH, vigV);
  S360_API_END
}
int s360_isp_process_dev(s360_ctx* c, s360_isp* isp, const void* rawDev, int inW, int inH, int depthBits, size_t strideBytes, void* outDev, int swizzle) {
  S360_API_BEGIN
  useDevice(c);
  require(isp != nullptr, "null isp");
  require(inW > 0 && inH > 0, "bad image size");
  ProfActivate prof(c);
  isp->pipe.process(c->stream, rawDev, inW, inH, depthBits, strideBytes ? strideBytes : (size_t)inW * (depthBits / 8), outDev, swizzle != 0);
  S360_API_END
}
int s360_isp_process_frame_dev(s360_ctx* c, s360_isp* isp, const void* rawDev, int inW, int inH, int depthBits, size_t strideBytes, uint8_t* outBgr8Dev) {
  S360_API_BEGIN
  useDevice(c);
  require(isp != nullptr, "null isp");
  require(inW > 0 && inH > 0, "bad image size");
  ProfActivate prof(c);
  isp->pipe.process(c->stream, rawDev, inW, inH, depthBits, strideBytes ? strideBytes : (size_t)inW * (depthBits / 8), outBgr8Dev, true, true);
  S360_API_END
}
int s360_isp_process_frames_dev(s360_ctx* c, s360_isp* isp, int n, const void* const* rawDev, int inW, int inH, int depthBits, size_t strideBytes,
                                uint8_t* const* outBgr8Dev) {
  S360_API_BEGIN
  useDevice(c);
  require(isp != nullptr && rawDev && outBgr8Dev, "null argument");
  require(inW > 0 && inH > 0, "bad image size");
  ProfActivate prof(c);
  isp->pipe.processBatch(c->stream, n, rawDev, inW, inH, depthBits, strideBytes ? strideBytes : (size_t)inW * (depthBits / 8),
                         reinterpret_cast<void* const*>(outBgr8Dev), true, true);
  S360_API_END
}
int s360_isp_process(s360_ctx* c, s360_isp* isp, const void* raw, int inW, int inH, int depthBits, size_t strideBytes, void* out, int swizzle) {
  S360_API_BEGIN
  useDevice(c);
  require(isp != nullptr && raw && out, "null argument");
  require(inW > 0 && inH > 0 && (depthBits == 8 || depthBits == 16), "bad image size or depth");
  cudaStream_t st = c->stream;
  const size_t rowBytes = (size_t)inW * (depthBits / 8);
  if (!strideBytes) strideBytes = rowBytes;
  require(strideBytes >= rowBytes, "row stride smaller than a row");
  const int r = isp->pipe.resize();
  const size_t outBytes = (size_t)(inW / r) * (inH / r) * 3 * (isp->pipe.outputBpp() / 8);
  DeviceBuffer dRaw(rowBytes * inH), dOut(outBytes);
  S360_CUDA_CHECK(cudaMemcpy2DAsync(dRaw.p, rowBytes, raw, strideBytes, rowBytes, inH, cudaMemcpyHostToDevice, st));
  isp->pipe.process(st, dRaw.p, inW, inH, depthBits, rowBytes, dOut.p, swizzle != 0);
  S360_CUDA_CHECK(cudaMemcpyAsync(out, dOut.p, outBytes, cudaMemcpyDeviceToHost, st));
  S360_CUDA_CHECK(cudaStreamSynchronize(st));
  S360_API_END
}

// ---- footage container + raw-sample conversion (footage.h, isp.h) ----
struct s360_footage { s360::FootageFile file; explicit s360_footage(const char* p) : file(p) {} };
int s360_footage_open(const char* path, s360_footage** out) {
  S360_API_BEGIN
  require(path && out, "null argument");
  *out = new s360_footage(path);
  S360_API_END
}
void s360_footage_close(s360_footage* f) { delete f; }
int s360_footage_info(s360_footage* f, s360_footage_header* header, uint64_t* frames, uint64_t* frameBytes) {
  S360_API_BEGIN
  require(f != nullptr, "null footage file");
  static_assert(sizeof(s360_footage_header) == sizeof(s360::FootageHeader), "header layout");
  if (header) std::memcpy(header, &f->file.header(), sizeof *header);
  if (frames) *frames = f->file.numberOfFrames();
  if (frameBytes) *frameBytes = f->file.frameSize();
  S360_API_END
}
int s360_footage_frame(s360_footage* f, uint32_t frame, uint32_t camera, const uint8_t** data) {
  S360_API_BEGIN
  require(f != nullptr && data != nullptr, "null argument");
  *data = f->file.frame(frame, camera);
  S360_API_END
}
int s360_raw_convert_dev(s360_ctx* c, int n, const void* const* packedDev, int bits, int w, int h, uint16_t* const* outDev) {
  S360_API_BEGIN
  useDevice(c);
  ProfActivate prof(c);
  s360::rawConvertBatch(c->stream, n, packedDev, bits, w, h, reinterpret_cast<void* const*>(outDev));
  S360_API_END
}
int s360_raw_convert(s360_ctx* c, const void* packed, int bits, int w, int h, uint16_t* out) {
  S360_API_BEGIN
  useDevice(c);
  require(packed && out, "null argument");
  require((bits == 8 || bits == 12) && w > 0 && h > 0, "footage samples are 8 or 12 bits per pixel");
  const size_t inBytes = (size_t)w * h * bits / 8, outBytes = (size_t)w * h * 2;
  DeviceBuffer dIn(inBytes), dOut(outBytes);
  S360_CUDA_CHECK(cudaMemcpyAsync(dIn.p, packed, inBytes, cudaMemcpyHostToDevice, c->stream));
  const void* src = dIn.p; void* dst = dOut.p;
  s360::rawConvertBatch(c->stream, 1, &src, bits, w, h, &dst);
  S360_CUDA_CHECK(cudaMemcpyAsync(out, dOut.p, outBytes, cudaMemcpyDeviceToHost, c->stream));
  S360_CUDA_CHECK(cudaStreamSynchronize(c->stream));
  S360_API_END
}
int s360_isp_process_footage_dev(s360_ctx* c, s360_isp* isp, int n, const void* const* packedDev, int bits, int w, int h, void* const* outDev, int highByte) {
  S360_API_BEGIN
  useDevice(c);
  require(isp != nullptr && packedDev && outDev, "null argument");
  require(w > 0 && h > 0, "bad image size");
  require(bits == 8 || bits == 12, "footage samples are 8 or 12 bits per pixel");
  ProfActivate prof(c);
  isp->pipe.processBatch(c->stream, n, packedDev, w, h, 16, (size_t)w * bits / 8, outDev, true, highByte != 0, bits);
  S360_API_END
}

int s360_isp_process_footage(s360_ctx* c, s360_isp* isp, const void* packed, int bits, int w, int h, uint16_t* out16, uint16_t* raw16Out) {
  S360_API_BEGIN
  useDevice(c);
  require(packed != nullptr, "null argument");
  require(isp == nullptr || out16 != nullptr, "null output image");
  require(w > 0 && h > 0, "bad image size");
  require(bits == 8 || bits == 12, "footage samples are 8 or 12 bits per pixel");
  require(isp == nullptr || isp->pipe.outputBpp() == 16, "the Unpacker's ISP has 16 output bits");
  const size_t px = (size_t)w * h, inBytes = px * bits / 8;
  cudaStream_t st = c->stream;
  DeviceBuffer dIn(inBytes), dRaw(raw16Out ? px * 2 : 0), dOut(isp ? px * 6 : 0);
  S360_CUDA_CHECK(cudaMemcpyAsync(dIn.p, packed, inBytes, cudaMemcpyHostToDevice, st));
  const void* src = dIn.p;
  if (raw16Out) {
    void* dst = dRaw.p;
    s360::rawConvertBatch(st, 1, &src, bits, w, h, &dst);
    S360_CUDA_CHECK(cudaMemcpyAsync(raw16Out, dRaw.p, px * 2, cudaMemcpyDeviceToHost, st));
  }
  if (isp) {
    void* dst = dOut.p;
    isp->pipe.processBatch(st, 1, &src, w, h, 16, (size_t)w * bits / 8, &dst, true, false, bits);
    S360_CUDA_CHECK(cudaMemcpyAsync(out16, dOut.p, px * 6, cudaMemcpyDeviceToHost, st));
  }
  S360_CUDA_CHECK(cudaStreamSynchronize(st));
  S360_API_END
}

int s360_flatten_deghost(s360_ctx* c, const uint8_t* base, const uint8_t* top, int w, int h, uint8_t* out) {
  S360_API_BEGIN
  useDevice(c);
  require(base && top && out && w > 0 && h > 0, "bad arguments");
  cudaStream_t st = c->stream;
  const size_t px = (size_t)w * h;
  DeviceBuffer dB(px * 4), dT(px * 4), dO(px * 4);
  S360_CUDA_CHECK(cudaMemcpyAsync(dB.p, base, px * 4, cudaMemcpyHostToDevice, st));
  S360_CUDA_CHECK(cudaMemcpyAsync(dT.p, top, px * 4, cudaMemcpyHostToDevice, st));
  launchFlatten(st, dB.as<uchar4>(), dT.as<uchar4>(), w, h, dO.as<uchar4>(), c->rt);
  S360_CUDA_CHECK(cudaMemcpyAsync(out, dO.p, px * 4, cudaMemcpyDeviceToHost, st));
  S360_CUDA_CHECK(cudaStreamSynchronize(st));
  S360_API_END
}

int s360_feather_alpha(s360_ctx* c, const uint8_t* src, int w, int h, int erodeSize, uint8_t* out) {
  S360_API_BEGIN
  useDevice(c);
  require(src && out && w > 0 && h > 0, "bad arguments");
  cudaStream_t st = c->stream;
  const size_t px = (size_t)w * h;
  DeviceBuffer dS(px * 4), dO(px * 4), dScr(featherScratchBytes(w, h));
  S360_CUDA_CHECK(cudaMemcpyAsync(dS.p, src, px * 4, cudaMemcpyHostToDevice, st));
  launchFeatherAlpha(st, dS.as<uchar4>(), w, h, erodeSize, c->blurKernelFor(erodeSize), dO.as<uchar4>(), dScr.as<uint8_t>());
  S360_CUDA_CHECK(cudaMemcpyAsync(out, dO.p, px * 4, cudaMemcpyDeviceToHost, st));
  S360_CUDA_CHECK(cudaStreamSynchronize(st));
  S360_API_END
}

int s360_offset_wrap(s360_ctx* c, const uint8_t* src, int w, int h, int channels, float offset, uint8_t* out) {
  S360_API_BEGIN
  useDevice(c);
  require(src && out && w > 0 && h > 0 && (channels == 3 || channels == 4), "bad arguments");
  cudaStream_t st = c->stream;
  const size_t px = (size_t)w * h;
  DeviceBuffer dRaw(px * channels), dS(px * 4), dO(px * 4), dOut(px * channels);
  S360_CUDA_CHECK(cudaMemcpyAsync(dRaw.p, src, px * channels, cudaMemcpyHostToDevice, st));
  launchUnpackBGRA(st, dRaw.as<uint8_t>(), channels, w, h, dS.as<uchar4>());
  launchOffsetWrap(st, dS.as<uchar4>(), w, h, offset, dO.as<uchar4>());
  if (channels == 3) launchPackBGR(st, dO.as<uchar4>(), w, h, w, dOut.as<uint8_t>());
  S360_CUDA_CHECK(cudaMemcpyAsync(out, channels == 3 ? dOut.p : dO.p, px * channels, cudaMemcpyDeviceToHost, st));
  S360_CUDA_CHECK(cudaStreamSynchronize(st));
  S360_API_END
}

int s360_sharpen(s360_ctx* c, uint8_t* bgr, int w, int h, double sharpening) {
  S360_API_BEGIN
  useDevice(c);
  require(bgr && w > 1 && h > 1, "bad arguments");
  cudaStream_t st = c->stream;
  const size_t px = (size_t)w * h;
  DeviceBuffer dRaw(px * 3), dS(px * 4), dF(px * 12), dLp(px * 8);
  S360_CUDA_CHECK(cudaMemcpyAsync(dRaw.p, bgr, px * 3, cudaMemcpyHostToDevice, st));
  launchUnpackBGRA(st, dRaw.as<uint8_t>(), 3, w, h, dS.as<uchar4>());
  launchSharpen(st, dS.as<uchar4>(), w, h, 1, (float)(1.0f + sharpening), dF.as<float>(), dLp.as<uchar4>());
  launchPackBGR(st, dS.as<uchar4>(), w, h, w, dRaw.as<uint8_t>());
  S360_CUDA_CHECK(cudaMemcpyAsync(bgr, dRaw.p, px * 3, cudaMemcpyDeviceToHost, st));
  S360_CUDA_CHECK(cudaStreamSynchronize(st));
  S360_API_END
}

int s360_cubemap(s360_ctx* c, const uint8_t* eye, int w, int h, int fw, int fh, int photo, uint8_t* out) {
  S360_API_BEGIN
  useDevice(c);
  require(eye && out && w > 0 && h > 0 && fw > 0 && fh > 0, "bad arguments");
  cudaStream_t st = c->stream;
  const size_t px = (size_t)w * h, opx = (size_t)6 * fw * fh;
  DeviceBuffer dRaw(px * 3), dS(px * 4), dO(opx * 3);
  S360_CUDA_CHECK(cudaMemcpyAsync(dRaw.p, eye, px * 3, cudaMemcpyHostToDevice, st));
  launchUnpackBGRA(st, dRaw.as<uint8_t>(), 3, w, h, dS.as<uchar4>());
  launchCubemap(st, dS.as<uchar4>(), w, h, fw, fh, photo, dO.as<uint8_t>(), c->rt);
  S360_CUDA_CHECK(cudaMemcpyAsync(out, dO.p, opx * 3, cudaMemcpyDeviceToHost, st));
  S360_CUDA_CHECK(cudaStreamSynchronize(st));
  S360_API_END
}

int s360_resize_cubic(s360_ctx* c, const uint8_t* src, int w, int h, int channels, int dw, int dh, uint8_t* dst) {
  S360_API_BEGIN
  useDevice(c);
  require(src && dst && w > 0 && h > 0 && dw > 0 && dh > 0 && (channels == 3 || channels == 4), "bad arguments");
  cudaStream_t st = c->stream;
  const size_t px = (size_t)w * h, opx = (size_t)dw * dh;
  DeviceBuffer dRaw(px * channels), dS(px * 4), dO(opx * 4), dOut(opx * 3);
  S360_CUDA_CHECK(cudaMemcpyAsync(dRaw.p, src, px * channels, cudaMemcpyHostToDevice, st));
  launchUnpackBGRA(st, dRaw.as<uint8_t>(), channels, w, h, dS.as<uchar4>());
  if (channels == 3) {
    launchResizeCubic8(st, dS.as<uchar4>(), w, h, w, 3, nullptr, dw, dh, dw, dOut.as<uint8_t>());
    S360_CUDA_CHECK(cudaMemcpyAsync(dst, dOut.p, opx * 3, cudaMemcpyDeviceToHost, st));
  } else {
    launchResizeCubic8(st, dS.as<uchar4>(), w, h, w, 4, dO.as<uchar4>(), dw, dh, dw, nullptr);
    S360_CUDA_CHECK(cudaMemcpyAsync(dst, dO.p, opx * 4, cudaMemcpyDeviceToHost, st));
  }
  S360_CUDA_CHECK(cudaStreamSynchronize(st));
  S360_API_END
}

int s360_frame_to_nv12_dev(s360_ctx* c, const uint8_t* bgrDev, size_t bgrStride, int w, int h, int matrix, uint8_t* yDev, size_t yPitch, uint8_t* uvDev,
                           size_t uvPitch) {
  S360_API_BEGIN
  useDevice(c);
  require(bgrDev && yDev && uvDev && w > 0 && h > 0, "bad arguments");
  require(matrix == 0 || matrix == 1, "matrix: 0 (BT.601) or 1 (BT.709)");
  if (!bgrStride) bgrStride = (size_t)w * 3;
  if (!yPitch) yPitch = (size_t)w;
  if (!uvPitch) uvPitch = (size_t)((w + 1) / 2) * 2;
  require(bgrStride >= (size_t)w * 3 && yPitch >= (size_t)w && uvPitch >= (size_t)((w + 1) / 2) * 2, "row pitch smaller than a row");
  ProfActivate prof(c);
  launchBgrToNv12(c->stream, bgrDev, bgrStride, w, h, matrix, yDev, yPitch, uvDev, uvPitch);
  S360_API_END
}
int s360_frame_to_nv12(s360_ctx* c, const uint8_t* bgr, int w, int h, int matrix, uint8_t* yOut, uint8_t* uvOut) {
  S360_API_BEGIN
  useDevice(c);
  require(bgr && yOut && uvOut && w > 0 && h > 0, "bad arguments");
  require(matrix == 0 || matrix == 1, "matrix: 0 (BT.601) or 1 (BT.709)");
  const size_t px = (size_t)w * h, cw = (size_t)((w + 1) / 2) * 2, chh = (size_t)(h + 1) / 2;
  cudaStream_t st = c->stream;
  DeviceBuffer dIn(px * 3), dY(px), dUV(cw * chh);
  S360_CUDA_CHECK(cudaMemcpyAsync(dIn.p, bgr, px * 3, cudaMemcpyHostToDevice, st));
  launchBgrToNv12(st, dIn.as<uint8_t>(), (size_t)w * 3, w, h, matrix, dY.as<uint8_t>(), (size_t)w, dUV.as<uint8_t>(), cw);
  S360_CUDA_CHECK(cudaMemcpyAsync(yOut, dY.p, px, cudaMemcpyDeviceToHost, st));
  S360_CUDA_CHECK(cudaMemcpyAsync(uvOut, dUV.p, cw * chh, cudaMemcpyDeviceToHost, st));
  S360_CUDA_CHECK(cudaStreamSynchronize(st));
  S360_API_END
}

int s360_pole_to_side(s360_ctx* c, const s360_params* p, const uint8_t* side, int sw, int sh, const uint8_t* fish, int fw, int fh,
                      uint8_t* warpedOut, float* flowOut) {
  S360_API_BEGIN
  useDevice(c);
  require(p && side && fish && warpedOut, "null argument");
  poleToSide(c, *p, side, sw, sh, fish, fw, fh, warpedOut, flowOut);
  S360_API_END
}

// ---- whole frame ----
int s360_frame_output_size(s360_ctx* c, const s360_params* p, int* eqW, int* eqH, int* cubeW, int* cubeH) {
  S360_API_BEGIN
  require(c && p && eqW && eqH && cubeW && cubeH, "null argument");
  frameOutputSize(c, *p, eqW, eqH, cubeW, cubeH);
  S360_API_END
}
int s360_render_frame(s360_ctx* c, const s360_params* p, const uint8_t* const* images, int w, int h, uint8_t* eq, uint8_t* cube) {
  S360_API_BEGIN
  useDevice(c);
  require(p && images && eq && w > 0 && h > 0, "bad arguments");
  renderFrameHost(c, *p, images, w, h, eq, cube);
  S360_API_END
}
int s360_render_frame_dev(s360_ctx* c, const s360_params* p, const uint8_t* const* images, int w, int h, uint8_t* eq, uint8_t* cube) {
  S360_API_BEGIN
  useDevice(c);
  require(p && images && eq && w > 0 && h > 0, "bad arguments");
  renderFrameDevice(c, *p, images, w, h, eq, cube);
  S360_API_END
}
int s360_frame_intermediate(s360_ctx* c, const char* name, int idx, void* out, int* w, int* h, int* ch, int* isFloat) {
  S360_API_BEGIN
  useDevice(c);
  require(name && w && h && ch && isFloat, "null argument");
  frameIntermediate(c, name, idx, out, w, h, ch, isFloat);
  S360_API_END
}
int s360_frame_stage_ms(s360_ctx* c, float out[7]) {
  S360_API_BEGIN
  require(c && out, "null argument");
  for (int i = 0; i < 7; ++i) out[i] = c->stageMs[i];
  S360_API_END
}

// ---- flow files (CvUtil.cpp:159-199): int32 rows, int32 cols, then (f32 x, f32 y) row-major ----
int s360_flow_save(const char* path, const float* flow, int w, int h) {
  S360_API_BEGIN
  require(path && flow && w > 0 && h > 0, "bad arguments");
  FILE* f = std::fopen(path, "wb");
  if (!f) throw std::runtime_error(std::string("file not found: ") + path);
  const int32_t rows = h, cols = w;
  const bool ok = std::fwrite(&rows, 4, 1, f) == 1 && std::fwrite(&cols, 4, 1, f) == 1 && std::fwrite(flow, 8, (size_t)w * h, f) == (size_t)w * h;
  if (std::fclose(f) != 0 || !ok) throw std::runtime_error(std::string("failed to write flow file: ") + path);
  S360_API_END
}
int s360_flow_load(const char* path, float** out, int* w, int* h) {
  S360_API_BEGIN
  require(path && out && w && h, "bad arguments");
  FILE* f = std::fopen(path, "rb");
  if (!f) throw std::runtime_error(std::string("file not found: ") + path);
  int32_t rows = 0, cols = 0;
  if (std::fread(&rows, 4, 1, f) != 1 || std::fread(&cols, 4, 1, f) != 1 || rows <= 0 || cols <= 0) {
    std::fclose(f);
    throw std::runtime_error(std::string("bad flow file: ") + path);
  }
  float* buf = (float*)std::malloc((size_t)rows * cols * 8);
  if (!buf) { std::fclose(f); throw std::runtime_error(std::string("out of memory reading flow file: ") + path); }
  const size_t got = std::fread(buf, 8, (size_t)rows * cols, f);
  std::fclose(f);
  if (got != (size_t)rows * cols) { std::free(buf); throw std::runtime_error(std::string("truncated flow file: ") + path); }
  *out = buf; *w = cols; *h = rows;
  S360_API_END
}
void s360_free(void* p) { std::free(p); }

}  // extern "C"
