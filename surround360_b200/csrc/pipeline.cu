// Per-frame orchestration of the render path on one B200. See pipeline.h.
#include "pipeline.h"
#include <algorithm>
#include <cmath>
#include <cstring>

namespace s360 {

void DeviceBuffer::alloc(size_t n) {
  if (n == 0) n = 256;
  S360_CUDA_CHECK(cudaMalloc(&p, n));
  bytes = n;
}
void DeviceBuffer::release() {
  if (p) cudaFree(p);
  p = nullptr;
  bytes = 0;
}

// 8.8 fixed-point Gaussian taps of OpenCV's bit-exact 8U GaussianBlur: error-diffusion
// rounding from the edge inwards, symmetric, centre takes the residual so the taps sum to 256.
std::vector<int> gaussKernelFixed8(int n, double sigma) {
  std::vector<double> t(n);
  double s2 = -0.5 / (sigma * sigma), sum = 0;
  for (int i = 0; i < n; ++i) { double x = i - (n - 1) * 0.5; t[i] = std::exp(s2 * x * x); sum += t[i]; }
  std::vector<int> k(n);
  int isum = 0;
  double err = 0;
  for (int i = 0; i < n / 2; ++i) {
    const double v = t[i] / sum * 256.0 + err;
    const int q = (int)std::floor(v + 0.5);
    err = v - q;
    k[i] = k[n - 1 - i] = q;
    isum += 2 * q;
  }
  k[n / 2] = 256 - isum;
  return k;
}

// cv::remap bicubic weight tables (imgwarp.cpp initInterTab2D): 32x32 sub-pixel positions,
// float products and 15-bit integers whose sum is forced to 2^15.
static void buildRemapTables(std::vector<short>& wi, std::vector<float>& wf) {
  wi.assign(1024 * 16, 0);
  wf.assign(1024 * 16, 0.f);
  float t1[32][4];
  for (int k = 0; k < 32; ++k) cubicCoeffs(k * (1.f / 32), t1[k]);
  for (int yi = 0; yi < 32; ++yi)
    for (int xi = 0; xi < 32; ++xi) {
      float* tf = &wf[(size_t)(yi * 32 + xi) * 16];
      short* ti = &wi[(size_t)(yi * 32 + xi) * 16];
      int isum = 0;
      for (int k1 = 0; k1 < 4; ++k1)
        for (int k2 = 0; k2 < 4; ++k2) {
          const float v = t1[yi][k1] * t1[xi][k2];
          tf[k1 * 4 + k2] = v;
          int q = (int)std::nearbyint((double)(v * 32768.f));
          q = std::min(std::max(q, -32768), 32767);
          ti[k1 * 4 + k2] = (short)q;
          isum += q;
        }
      if (isum != 32768) {
        const int diff = isum - 32768;
        int Mk1 = 2, Mk2 = 2, mk1 = 2, mk2 = 2;
        for (int k1 = 2; k1 < 4; ++k1)
          for (int k2 = 2; k2 < 4; ++k2) {
            if (ti[k1 * 4 + k2] < ti[mk1 * 4 + mk2]) { mk1 = k1; mk2 = k2; }
            else if (ti[k1 * 4 + k2] > ti[Mk1 * 4 + Mk2]) { Mk1 = k1; Mk2 = k2; }
          }
        if (diff < 0) ti[Mk1 * 4 + Mk2] = (short)(ti[Mk1 * 4 + Mk2] - diff);
        else ti[mk1 * 4 + mk2] = (short)(ti[mk1 * 4 + mk2] - diff);
      }
    }
}
void initContextTables(s360_ctx* c) {
  std::vector<short> wi;
  std::vector<float> wf;
  buildRemapTables(wi, wf);
  c->tabI.alloc(wi.size() * 2);
  c->tabF.alloc(wf.size() * 4);
  S360_CUDA_CHECK(cudaMemcpy(c->tabI.p, wi.data(), wi.size() * 2, cudaMemcpyHostToDevice));
  S360_CUDA_CHECK(cudaMemcpy(c->tabF.p, wf.data(), wf.size() * 4, cudaMemcpyHostToDevice));
  c->rt.wi = c->tabI.as<short>();
  c->rt.wf = c->tabF.as<float>();
  // host libm tables (see RemapTables)
  std::vector<float> t(766 + 766 + 256);
  for (int n = 0; n < 766; ++n) {
    const float colorDiff = (float)n / 255.0f;
    t[n] = tanhf(colorDiff * 10.0f);
    t[766 + n] = tanhf(colorDiff * 5.0f);
  }
  for (int a = 0; a < 256; ++a) {
    const float alphaR = (float)a / 255.0f;
    const float alphaL = 1.0f - alphaR;
    const double expL = expf(5.0f * alphaL * 2.0f);  // float overload in the reference (CvUtil.cpp:245-246)
    const double expR = expf(5.0f * alphaR);
    const double sumExp = expL + expR + 0.00001;
    t[1532 + a] = float(expL / sumExp);
  }
  c->tabT.alloc(t.size() * 4);
  S360_CUDA_CHECK(cudaMemcpy(c->tabT.p, t.data(), t.size() * 4, cudaMemcpyHostToDevice));
  c->rt.tanh10 = c->tabT.as<float>();
  c->rt.tanh5 = c->tabT.as<float>() + 766;
  c->rt.flatSoftL = c->tabT.as<float>() + 1532;
}

// ------------------------------------------------------------------------------------
// geometry: the float/double expression order follows the reference line by line
// ------------------------------------------------------------------------------------
Geometry computeGeometry(const Rig& rig, const s360_params& P) {
  Geometry g{};
  const int numCams = (int)rig.side.size();
  const float fovH = rig.approximateFov(false), fovV = rig.approximateFov(true);
  g.hRadians = 2 * fovH;                                                     // TRSP:155-156
  g.vRadians = 2 * fovV;
  g.camH = (int)(P.eqr_height * g.vRadians / M_PI);                           // TRSP:159-162
  g.camW = (int)(P.eqr_width * g.hRadians / (2 * M_PI));
  const float camFovDeg = (float)(2 * fovH * (180 / M_PI));                   // TRSP:781-782 -> float param
  g.fovHRad = (float)(camFovDeg * M_PI / 180.0f);                             // MathUtil.h:32
  const float overlapDeg = (float)((camFovDeg * float(numCams) - 360.0) / float(numCams));  // TRSP:310-311
  g.overlapW = (int)(float(g.camW) * (overlapDeg / camFovDeg));               // TRSP:314-315
  g.numNovel = g.camW - g.overlapW;
  const float v = atanf((float)(P.zero_parallax_dist / (P.interpupilary_dist / 2.0f)));  // TRSP:340-348
  const float psi = asinf((float)(sinf(v) * (P.interpupilary_dist / 2.0f) / rig.getRingRadius()));
  g.slabDisp = psi * (float(g.camW) / g.fovHRad);
  const float theta = (float)(-M_PI / 2.0f + v + psi);
  g.shiftPx = (float)(float(P.eqr_width) * (theta / (2.0f * M_PI)));
  // poles
  const Camera& bottom = rig.cameras[rig.bottomCamera()];
  g.poleH = (int)(P.eqr_height * bottom.getFov() / M_PI);                     // TRSP:607-610
  g.extW = (int)(float(P.eqr_width) * 1.2f);                                  // TRSP:400-401
  float poleCameraRadius = (float)bottom.getFov();                            // TRSP:461
  float sideCameraRadius = fovV;                                              // TRSP:464
  float poleCameraCropRadius = (float)(0.5f * (M_PI / 2 - sideCameraRadius) + 0.5f * (std::min(float(M_PI / 2), poleCameraRadius)));
  poleCameraCropRadius = (float)(poleCameraCropRadius * (180 / M_PI));
  poleCameraRadius = (float)(poleCameraRadius * (180 / M_PI));
  sideCameraRadius = (float)(sideCameraRadius * (180 / M_PI));
  const float phiFromPole = poleCameraCropRadius, phiFromSide = 90.0f - sideCameraRadius;
  g.phiMid = (phiFromPole + phiFromSide) / 2.0f;
  const float phiDiff = fabsf(phiFromPole - phiFromSide);
  g.phiRampStart = g.phiMid - 1.0f * phiDiff / 2.0f;
  g.phiRampEnd = g.phiMid + 1.0f * phiDiff / 2.0f;
  g.poleCameraRadiusDeg = poleCameraRadius;
  return g;
}

// trig tables of bicubicRemapToSpherical (ImageWarper.cpp:150-163): float fractions and
// angles, cosf/sinf on the host -- the per-column / per-row "rig tables" every rank needs.
static void sphericalTablesHost(int w, int h, float left, float right, float top, float bottom, std::vector<float>& out) {
  out.resize((size_t)2 * w + 2 * h);
  for (int x = 0; x < w; ++x) {
    const float xFrac = (x + 0.5f) / w;
    const float xAngle = (1 - xFrac) * left + xFrac * right;
    out[x] = cosf(xAngle);
    out[w + x] = sinf(xAngle);
  }
  for (int y = 0; y < h; ++y) {
    const float yFrac = (y + 0.5f) / h;
    const float yAngle = (1 - yFrac) * top + yFrac * bottom;
    out[2 * w + y] = cosf(yAngle);
    out[2 * w + h + y] = sinf(yAngle);
  }
}
struct DevSphericalTables {
  DeviceBuffer buf;
  SphericalTables t{};
  void upload(int w, int h, float l, float r, float tp, float b, cudaStream_t st) {
    std::vector<float> hst;
    sphericalTablesHost(w, h, l, r, tp, b, hst);
    buf.ensure(hst.size() * 4);
    S360_CUDA_CHECK(cudaMemcpyAsync(buf.p, hst.data(), hst.size() * 4, cudaMemcpyHostToDevice, st));
    S360_CUDA_CHECK(cudaStreamSynchronize(st));
    float* d = buf.as<float>();
    t.cosX = d; t.sinX = d + w; t.cosY = d + 2 * w; t.sinY = d + 2 * w + h;
  }
};

// ------------------------------------------------------------------------------------
// frame plan
// ------------------------------------------------------------------------------------
struct FramePlan {
  s360_params P;
  int imgW = 0, imgH = 0, numCams = 0, nRig = 0;
  Geometry g;
  int padAbove = 0;
  bool top = false, bottom = false;
  int nPole = 0;  // number of pole flows (2 per enabled pole)
  std::vector<DevSphericalTables> sideTables;
  DevSphericalTables topTables, botTables;
  DeviceBuffer srcRaw;     // staging for host inputs: nRig * imgW*imgH*3
  DeviceBuffer srcBGRA;    // one prepared source at a time (reused)
  DeviceBuffer proj;       // [numCams][camH][camW] uchar4
  DeviceBuffer flowSide;   // [2][numCams][camH][overlapW] float2: LtoR then RtoL
  DeviceBuffer panoL, panoR;        // eqrW x eqrH uchar4 (pad rows stay zero)
  DeviceBuffer topSph, botSph;      // eqrW x poleH
  DeviceBuffer prA, prB, prC, prFlow, prScratch;   // pole removal work images (imgW x imgH)
  std::unique_ptr<FlowBatch> prFlowBatch;
  bool poleRemoval = false;
  // temporal regularisation across consecutive frames (video mode): previous projections, flows and
  // extended fisheye images stay on the device instead of going through .bin / PNG files
  bool temporal = false, prevValid = false;
  DeviceBuffer prevProj, prevFlowSide, prevExtFish, prevPoleFlow;
  DeviceBuffer extSide, extFish;    // [nPole][poleH][extW]
  DeviceBuffer featherScratch;
  DeviceBuffer poleFlow;            // [nPole][poleH][extW] float2
  DeviceBuffer poleWarped;          // [nPole][poleH][eqrW]
  DeviceBuffer eyes;                // 2 x (eqrW x eqrH) uchar4, left then right
  uchar4* eye(int e, size_t eqPx) const { return eyes.as<uchar4>() + (size_t)e * eqPx; }
  DeviceBuffer eyePre[2];           // copy before sharpening (keepIntermediates)
  DeviceBuffer workspace;           // shared by the flow batches and the sharpen scratch (never concurrent)
  float* sharpF = nullptr;
  uchar4* sharpLp = nullptr;
  DeviceBuffer outEq, outCube;      // device outputs for the host variant
  std::unique_ptr<FlowBatch> sideFlow, poleFlowBatch;
  cudaEvent_t ev[8];
  cudaGraphExec_t graphExec = nullptr;
  bool graphHasCube = false, graphUsesPrev = false, lastWasTimed = false;
  unsigned long long graphLaunches = 0;
  bool valid = false;
  bool evCreated = false;
  ~FramePlan() {
    if (graphExec) cudaGraphExecDestroy(graphExec);
    if (evCreated) for (auto& e : ev) cudaEventDestroy(e);
  }
};

static bool sameParams(const s360_params& a, const s360_params& b) { return std::memcmp(&a, &b, sizeof(a)) == 0; }
static bool planMatches(s360_ctx* c, const s360_params& P, int imgW, int imgH) {
  return c->plan && sameParams(c->plan->P, P) && c->plan->imgW == imgW && c->plan->imgH == imgH && c->plan->temporal == c->temporal;
}

static void buildPlan(s360_ctx* c, const s360_params& P, int imgW, int imgH) {
  if (!c->rigLoaded) throw std::runtime_error("no rig loaded");
  auto plan = std::make_unique<FramePlan>();
  FramePlan& F = *plan;
  F.P = P; F.imgW = imgW; F.imgH = imgH;
  const Rig& rig = c->rig;
  F.numCams = (int)rig.side.size();
  F.nRig = (int)rig.cameras.size();
  if (P.eqr_width % F.numCams != 0) throw std::runtime_error("eqr_width must be evenly divisible by the number of cameras");
  F.g = computeGeometry(rig, P);
  const Geometry& g = F.g;
  if (g.camH > P.eqr_height || g.camW <= 0 || g.overlapW <= 0 || g.numNovel <= 0) throw std::runtime_error("degenerate side geometry");
  FlowParams sideFP, poleFP;
  if (!flowParamsByName(P.side_flow_alg, &sideFP)) throw std::runtime_error(std::string("unrecognized flow algorithm name: ") + P.side_flow_alg);
  if (!flowParamsByName(P.polar_flow_alg, &poleFP)) throw std::runtime_error(std::string("unrecognized flow algorithm name: ") + P.polar_flow_alg);
  F.padAbove = (P.eqr_height - g.camH) / 2;  // TRSP:702
  F.top = P.enable_top != 0; F.bottom = P.enable_bottom != 0;
  F.nPole = (F.top ? 2 : 0) + (F.bottom ? 2 : 0);
  cudaStream_t st = c->stream;
  // tables
  F.sideTables.resize(F.numCams);
  for (int i = 0; i < F.numCams; ++i) {
    const float direction = (float)(-float(i) / float(F.numCams) * 2.0f * M_PI);  // TRSP:164
    F.sideTables[i].upload(g.camW, g.camH, direction + g.hRadians / 2, direction - g.hRadians / 2, g.vRadians / 2, -g.vRadians / 2, st);
  }
  if (F.top) {
    const Camera& cam = rig.cameras[rig.topCamera()];
    F.topTables.upload(P.eqr_width, g.poleH, (float)(2.0f * M_PI), 0.f, (float)(M_PI / 2.0f), (float)(M_PI / 2.0f - cam.getFov()), st);  // TRSP:660-667
  }
  if (F.bottom) {
    const Camera& cam = rig.cameras[rig.bottomCamera()];
    F.botTables.upload(P.eqr_width, g.poleH, 0.f, (float)(2.0f * M_PI), (float)(-(M_PI / 2.0f)), (float)(-(M_PI / 2.0f - cam.getFov())), st);  // TRSP:611-618
  }
  const size_t eqPx = (size_t)P.eqr_width * P.eqr_height;
  F.srcRaw.alloc((size_t)F.nRig * imgW * imgH * 3);
  F.srcBGRA.alloc((size_t)imgW * imgH * 4);
  F.proj.alloc((size_t)F.numCams * g.camW * g.camH * 4);
  F.flowSide.alloc((size_t)2 * F.numCams * g.overlapW * g.camH * 8);
  F.panoL.alloc(eqPx * 4); F.panoR.alloc(eqPx * 4);
  S360_CUDA_CHECK(cudaMemsetAsync(F.panoL.p, 0, eqPx * 4, st));
  S360_CUDA_CHECK(cudaMemsetAsync(F.panoR.p, 0, eqPx * 4, st));
  F.temporal = c->temporal;
  F.sideFlow = std::make_unique<FlowBatch>(2 * F.numCams, g.overlapW, g.camH, sideFP, F.temporal, false, true);
  F.sideFlow->setLatencyMode(c->latencyMode);
  if (F.temporal) { F.prevProj.alloc(F.proj.bytes); F.prevFlowSide.alloc(F.flowSide.bytes); }
  if (F.nPole) {
    if (g.poleH > P.eqr_height) throw std::runtime_error("pole image taller than the equirect");
    const size_t polePx = (size_t)P.eqr_width * g.poleH, extPx = (size_t)g.extW * g.poleH;
    F.topSph.alloc(polePx * 4); F.botSph.alloc(polePx * 4);
    F.extSide.alloc(extPx * 4 * F.nPole); F.extFish.alloc(extPx * 4 * F.nPole);
    F.featherScratch.alloc(featherScratchBytes(P.eqr_width, g.poleH));
    F.poleFlow.alloc(extPx * 8 * F.nPole);
    F.poleWarped.alloc(polePx * 4 * F.nPole);
    F.poleFlowBatch = std::make_unique<FlowBatch>(F.nPole, g.extW, g.poleH, poleFP, F.temporal, false, true);
    F.poleFlowBatch->setLatencyMode(c->latencyMode);
    if (F.temporal) { F.prevExtFish.alloc(F.extFish.bytes); F.prevPoleFlow.alloc(F.poleFlow.bytes); }
    (void)c->blurKernelFor(P.std_alpha_feather_size);  // allocate now: nothing may allocate while a frame is captured
    F.poleRemoval = F.bottom && P.enable_pole_removal != 0;
    if (F.poleRemoval) {
      FlowParams prFP;
      if (!flowParamsByName(P.poleremoval_flow_alg, &prFP)) throw std::runtime_error(std::string("unrecognized flow algorithm name: ") + P.poleremoval_flow_alg);
      if (!c->poleMasks.p || c->poleMaskW != imgW || c->poleMaskH != imgH) throw std::runtime_error("missing or bad pole mask: call s360_set_pole_masks with masks of the camera image size");
      const size_t ipx = (size_t)imgW * imgH;
      F.prA.alloc(ipx * 4); F.prB.alloc(ipx * 4); F.prC.alloc(ipx * 4); F.prFlow.alloc(ipx * 8);
      F.prScratch.alloc(featherScratchBytes(imgW, imgH));
      F.prFlowBatch = std::make_unique<FlowBatch>(1, imgW, imgH, prFP, false, false, true);
      F.prFlowBatch->setLatencyMode(c->latencyMode);
    }
  }
  F.eyes.alloc(eqPx * 4 * 2);   // both eyes back to back so one launch sharpens them together
  // One workspace for everything that never runs at the same time within a frame (one stream): the side, pole
  // and pole-removal flow batches and the sharpen scratch.
  const size_t sharpBytes = P.sharpening > 0.0f ? eqPx * 3 * 4 * 2 + eqPx * 4 * 2 * 2 : 0;
  size_t wsBytes = std::max(F.sideFlow->bytes(), sharpBytes);
  if (F.poleFlowBatch) wsBytes = std::max(wsBytes, F.poleFlowBatch->bytes());
  if (F.prFlowBatch) wsBytes = std::max(wsBytes, F.prFlowBatch->bytes());
  F.workspace.alloc(wsBytes);
  F.sideFlow->bindWorkspace(F.workspace.p, wsBytes);
  if (F.poleFlowBatch) F.poleFlowBatch->bindWorkspace(F.workspace.p, wsBytes);
  if (F.prFlowBatch) F.prFlowBatch->bindWorkspace(F.workspace.p, wsBytes);
  F.sharpF = F.workspace.as<float>();
  F.sharpLp = reinterpret_cast<uchar4*>(F.workspace.as<uint8_t>() + eqPx * 3 * 4 * 2);
  for (auto& e : F.ev) S360_CUDA_CHECK(cudaEventCreate(&e));
  F.evCreated = true;
  F.valid = true;
  c->plan = std::move(plan);
}

void frameOutputSize(s360_ctx* c, const s360_params& P, int* eqW, int* eqH, int* cubeW, int* cubeH) {
  (void)c;
  const bool resize = P.final_eqr_width != 0 && P.final_eqr_height != 0 && P.final_eqr_width != P.eqr_width &&
                      P.final_eqr_height != P.eqr_height / 2;  // TRSP:938-941
  *eqW = resize ? P.final_eqr_width : P.eqr_width;
  *eqH = resize ? (P.final_eqr_height / 2) * 2 : P.eqr_height * 2;
  const bool cube = P.want_cubemap && P.cubemap_width > 0 && P.cubemap_height > 0;
  *cubeW = !cube ? 0 : (P.cubemap_photo ? P.cubemap_width : 3 * P.cubemap_width);
  *cubeH = !cube ? 0 : 2 * (P.cubemap_photo ? 6 * P.cubemap_height : 2 * P.cubemap_height);
}

// Enqueues every kernel of one frame on the context stream. `timed`: record the stage events and
// the per-sweep profiling events (not possible while the stream is being captured into a graph).
static void enqueueFrame(s360_ctx* c, const uint8_t* const* imagesDev, uint8_t* equirectDev, uint8_t* cubemapDev, bool timed) {
  FramePlan& F = *c->plan;
  const s360_params& P = F.P;
  const int imgW = F.imgW, imgH = F.imgH;
  const Geometry& g = F.g;
  const Rig& rig = c->rig;
  cudaStream_t st = c->stream;
  const int eqW = P.eqr_width, eqH = P.eqr_height;
  auto stageEvent = [&](int i) { if (timed) S360_CUDA_CHECK(cudaEventRecord(F.ev[i], st)); };
  stageEvent(0);
  const bool usePrev = F.temporal && F.prevValid;
  F.sideFlow->enableProfile(timed && c->profile);
  if (F.poleFlowBatch) F.poleFlowBatch->enableProfile(timed && c->profile);

  // ---- projections (TRSP:138-175, 564-685) ----
  uchar4* proj = F.proj.as<uchar4>();
  for (int i = 0; i < F.numCams; ++i) {
    const int ci = rig.side[i];
    launchPrepSource(st, imagesDev[ci], 3, imgW, imgH, P.side_alpha_feather_size, F.srcBGRA.as<uchar4>());
    launchProjectSpherical(st, F.srcBGRA.as<uchar4>(), imgW, imgH, rig.cameras[ci].toDevice(), F.sideTables[i].t,
                           proj + (size_t)i * g.camW * g.camH, g.camW, g.camH, g.camW, 0, 0, 1, c->rt);
  }
  const int featherStart = g.poleH - 1 - P.std_alpha_feather_size;
  if (F.top) {
    const int ci = rig.topCamera();
    launchPrepSource(st, imagesDev[ci], 3, imgW, imgH, 0, F.srcBGRA.as<uchar4>());
    launchProjectSpherical(st, F.srcBGRA.as<uchar4>(), imgW, imgH, rig.cameras[ci].toDevice(), F.topTables.t, F.topSph.as<uchar4>(), eqW,
                           g.poleH, eqW, 1, std::max(featherStart, 0), P.std_alpha_feather_size, c->rt);
  }
  if (F.bottom) {
    const int ci = rig.bottomCamera();
    if (F.poleRemoval) {  // TRSP:569-597, PoleRemoval.cpp:32-188
      const int ci2 = rig.findLargestDistCamAxisToRigCenter();
      const Camera &cam = rig.cameras[ci], &cam2 = rig.cameras[ci2];
      const bool flip180 = dot(cam.up(), cam2.up()) < 0;
      const float r1 = cam.approximateUsablePixelsRadius(), r2 = cam2.approximateUsablePixelsRadius();
      const int* fk = c->blurKernelFor(P.std_alpha_feather_size);
      const uint8_t* m1 = c->poleMasks.as<uint8_t>();
      const uint8_t* m2 = m1 + (size_t)imgW * imgH * 3;
      uchar4 *A = F.prA.as<uchar4>(), *B = F.prB.as<uchar4>(), *C = F.prC.as<uchar4>();
      launchPoleRemovalPrepare(st, imagesDev[ci], m1, imgW, imgH, r1, C);
      launchFeatherAlpha(st, C, imgW, imgH, P.std_alpha_feather_size, fk, A, F.prScratch.as<uint8_t>());        // A = bottomImage
      launchPoleRemovalPrepare(st, imagesDev[ci2], m2, imgW, imgH, r2, C);
      launchFeatherAlpha(st, C, imgW, imgH, P.std_alpha_feather_size, fk, B, F.prScratch.as<uint8_t>());        // B = bottomImage2
      if (flip180) { launchFlipBoth(st, B, imgW, imgH, C); std::swap(B, C); }
      FlowImagePtrs ip{};
      FlowPrevPtrs pv{};
      ip.i0[0] = A; ip.i1[0] = B; ip.pitch0[0] = ip.pitch1[0] = imgW; ip.hint[0] = S360_HINT_DOWN;
      F.prFlowBatch->run(st, ip, pv, F.prFlow.as<float2>());
      launchPoleRemovalMerge(st, A, B, F.prFlow.as<float2>(), imgW, imgH, r1, C, c->rt);
      launchFeatherAlpha(st, C, imgW, imgH, P.std_alpha_feather_size, fk, F.srcBGRA.as<uchar4>(), F.prScratch.as<uint8_t>());
      if (c->keepIntermediates) S360_CUDA_CHECK(cudaMemcpyAsync(A, F.srcBGRA.p, (size_t)imgW * imgH * 4, cudaMemcpyDeviceToDevice, st));
      launchProjectSpherical(st, F.srcBGRA.as<uchar4>(), imgW, imgH, cam.toDevice(), F.botTables.t, F.botSph.as<uchar4>(), eqW,
                             g.poleH, eqW, 2, std::max(featherStart, 0), P.std_alpha_feather_size, c->rt);
    } else {
      launchPrepSource(st, imagesDev[ci], 3, imgW, imgH, 0, F.srcBGRA.as<uchar4>());
      launchProjectSpherical(st, F.srcBGRA.as<uchar4>(), imgW, imgH, rig.cameras[ci].toDevice(), F.botTables.t, F.botSph.as<uchar4>(), eqW,
                             g.poleH, eqW, 1, std::max(featherStart, 0), P.std_alpha_feather_size, c->rt);
    }
  }
  stageEvent(1);

  // ---- side optical flow (TRSP:318-337, NovelView.cpp:270-299) ----
  {
    FlowImagePtrs ip{};
    FlowPrevPtrs pv{};
    for (int i = 0; i < F.numCams; ++i) {
      const int r = (i + 1) % F.numCams;
      const uchar4* ovL = proj + (size_t)i * g.camW * g.camH + (g.camW - g.overlapW);
      const uchar4* ovR = proj + (size_t)r * g.camW * g.camH;
      ip.i0[i] = ovL; ip.i1[i] = ovR; ip.hint[i] = S360_HINT_LEFT;                           // flowLtoR
      ip.i0[F.numCams + i] = ovR; ip.i1[F.numCams + i] = ovL; ip.hint[F.numCams + i] = S360_HINT_RIGHT;  // flowRtoL
      ip.pitch0[i] = ip.pitch1[i] = ip.pitch0[F.numCams + i] = ip.pitch1[F.numCams + i] = g.camW;
      if (usePrev) {  // TRSP:210-244, NovelView.cpp:282-297: previous flow of the same direction, previous I1 image
        const uchar4* pproj = F.prevProj.as<uchar4>();
        const size_t per = (size_t)g.overlapW * g.camH;
        pv.prevFlow[i] = F.prevFlowSide.as<float2>() + (size_t)i * per;
        pv.prevI1[i] = pproj + (size_t)r * g.camW * g.camH;                                        // previous overlap R
        pv.prevFlow[F.numCams + i] = F.prevFlowSide.as<float2>() + (size_t)(F.numCams + i) * per;
        pv.prevI1[F.numCams + i] = pproj + (size_t)i * g.camW * g.camH + (g.camW - g.overlapW);   // previous overlap L
        pv.prevPitch[i] = pv.prevPitch[F.numCams + i] = g.camW;
      }
    }
    pv.enabled = usePrev;
    F.sideFlow->run(st, ip, pv, F.flowSide.as<float2>());
  }
  stageEvent(2);

  // ---- novel views -> eye panoramas (TRSP:339-385, 806-807) ----
  {
    NovelViewArgs a{};
    a.proj = proj;
    a.flowLR = F.flowSide.as<float2>();
    a.flowRL = F.flowSide.as<float2>() + (size_t)F.numCams * g.overlapW * g.camH;
    a.numCams = F.numCams; a.camW = g.camW; a.camH = g.camH; a.overlapW = g.overlapW; a.numNovel = g.numNovel;
    a.bufW = eqW / F.numCams; a.slabDisp = g.slabDisp; a.shiftPx = g.shiftPx;
    a.panoL = F.panoL.as<uchar4>(); a.panoR = F.panoR.as<uchar4>(); a.eqrW = eqW; a.padAbove = F.padAbove;
    launchNovelView(st, a, c->rt);
  }
  stageEvent(3);

  // ---- poles (TRSP:811-885, 388-561) ----
  const uchar4* topLayer[2] = {nullptr, nullptr};
  const uchar4* botLayer[2] = {nullptr, nullptr};
  if (F.nPole) {
    const size_t extPx = (size_t)g.extW * g.poleH, polePx = (size_t)eqW * g.poleH;
    const int* fk = c->blurKernelFor(P.std_alpha_feather_size);
    FlowImagePtrs ip{};
    FlowPrevPtrs pv{};
    int k = 0;
    struct Slot { bool isTop; int eye; };
    std::vector<Slot> slots;
    if (F.top) { slots.push_back({true, 0}); slots.push_back({true, 1}); }
    if (F.bottom) { slots.push_back({false, 0}); slots.push_back({false, 1}); }
    for (const Slot& s : slots) {
      const uchar4* pano = s.eye == 0 ? F.panoL.as<uchar4>() : F.panoR.as<uchar4>();
      uchar4* extSide = F.extSide.as<uchar4>() + (size_t)k * extPx;
      uchar4* extFish = F.extFish.as<uchar4>() + (size_t)k * extPx;
      // crop to the pole rows (of the flipped panorama for the bottom), feather, wrap-extend
      launchFeatherExtend(st, pano, eqW, eqH, g.poleH, s.isTop ? 0 : 1, P.std_alpha_feather_size, fk, extSide, g.extW,
                          F.featherScratch.as<uint8_t>());
      launchExtendWrap(st, s.isTop ? F.topSph.as<uchar4>() : F.botSph.as<uchar4>(), eqW, g.poleH, extFish, g.extW);
      ip.i0[k] = extSide; ip.i1[k] = extFish; ip.pitch0[k] = ip.pitch1[k] = g.extW; ip.hint[k] = S360_HINT_DOWN;
      if (usePrev) {  // TRSP:418-447
        pv.prevFlow[k] = F.prevPoleFlow.as<float2>() + (size_t)k * extPx;
        pv.prevI1[k] = F.prevExtFish.as<uchar4>() + (size_t)k * extPx;
        pv.prevPitch[k] = g.extW;
      }
      ++k;
    }
    pv.enabled = usePrev;
    F.poleFlowBatch->run(st, ip, pv, F.poleFlow.as<float2>());
    k = 0;
    for (const Slot& s : slots) {
      PoleWarpArgs a{};
      a.extFish = F.extFish.as<uchar4>() + (size_t)k * extPx;
      a.flow = F.poleFlow.as<float2>() + (size_t)k * extPx;
      a.cols = eqW; a.rows = g.poleH; a.extW = g.extW;
      a.poleCameraRadiusDeg = g.poleCameraRadiusDeg; a.phiRampStart = g.phiRampStart; a.phiMid = g.phiMid; a.phiRampEnd = g.phiRampEnd;
      a.warped = F.poleWarped.as<uchar4>() + (size_t)k * polePx;
      launchPoleWarp(st, a, c->rt);
      (s.isTop ? topLayer : botLayer)[s.eye] = a.warped;
      ++k;
    }
  }
  if (F.temporal) {  // what the next frame's flows regularise against
    S360_CUDA_CHECK(cudaMemcpyAsync(F.prevProj.p, F.proj.p, F.proj.bytes, cudaMemcpyDeviceToDevice, st));
    S360_CUDA_CHECK(cudaMemcpyAsync(F.prevFlowSide.p, F.flowSide.p, F.flowSide.bytes, cudaMemcpyDeviceToDevice, st));
    if (F.nPole) {
      S360_CUDA_CHECK(cudaMemcpyAsync(F.prevExtFish.p, F.extFish.p, F.extFish.bytes, cudaMemcpyDeviceToDevice, st));
      S360_CUDA_CHECK(cudaMemcpyAsync(F.prevPoleFlow.p, F.poleFlow.p, F.poleFlow.bytes, cudaMemcpyDeviceToDevice, st));
    }
  }
  // flatten (or plain copy when no poles): eye = BGRA panorama with the pole layers merged
  launchFlattenPoles(st, F.panoL.as<uchar4>(), eqW, eqH, topLayer[0], botLayer[0], g.poleH, F.eye(0, (size_t)eqW * eqH), c->rt);
  launchFlattenPoles(st, F.panoR.as<uchar4>(), eqW, eqH, topLayer[1], botLayer[1], g.poleH, F.eye(1, (size_t)eqW * eqH), c->rt);
  stageEvent(4);

  // ---- sharpen (TRSP:901-913) ----
  if (c->keepIntermediates)
    for (int e = 0; e < 2; ++e) {
      F.eyePre[e].ensure((size_t)eqW * eqH * 4);
      S360_CUDA_CHECK(cudaMemcpyAsync(F.eyePre[e].p, F.eye(e, (size_t)eqW * eqH), (size_t)eqW * eqH * 4, cudaMemcpyDeviceToDevice, st));
    }
  if (P.sharpening > 0.0f)
    launchSharpen(st, F.eyes.as<uchar4>(), eqW, eqH, 2, (float)(1.0f + P.sharpening), F.sharpF, F.sharpLp);
  stageEvent(5);

  // ---- cubemap + final equirect (TRSP:915-961) ----
  int oW, oH, cW, cH;
  frameOutputSize(c, P, &oW, &oH, &cW, &cH);
  if (cW > 0 && cubemapDev)
    for (int e = 0; e < 2; ++e)
      launchCubemap(st, F.eye(e, (size_t)eqW * eqH), eqW, eqH, P.cubemap_width, P.cubemap_height, P.cubemap_photo,
                    cubemapDev + (size_t)e * cW * (cH / 2) * 3, c->rt);
  for (int e = 0; e < 2; ++e) {
    uint8_t* dst = equirectDev + (size_t)e * oW * (oH / 2) * 3;
    if (oW != eqW || oH / 2 != eqH) launchResizeCubic8(st, F.eye(e, (size_t)eqW * eqH), eqW, eqH, eqW, 3, nullptr, oW, oH / 2, oW, dst);
    else launchPackBGR(st, F.eye(e, (size_t)eqW * eqH), eqW, eqH, eqW, dst);
  }
  stageEvent(6);
  // a rejected launch (bad configuration) does not stick to the stream: fail here instead of returning a stale frame
  S360_CUDA_CHECK(cudaGetLastError());
}

// A frame is ~1000 small launches; replaying it as one CUDA graph removes the host launch cost and
// the launch-queue back-pressure that otherwise throttles several frames in flight. The graph
// reads the plan-owned input staging buffer and writes the plan-owned output buffers, so it is
// captured once per plan; per call only the D2D copies in and out are enqueued around it.
static void runFrameFromStaging(s360_ctx* c, bool wantCube) {
  FramePlan& F = *c->plan;
  cudaStream_t st = c->stream;
  const size_t imgBytes = (size_t)F.imgW * F.imgH * 3;
  std::vector<const uint8_t*> dev(F.nRig);
  for (int i = 0; i < F.nRig; ++i) dev[i] = F.srcRaw.as<uint8_t>() + (size_t)i * imgBytes;
  uint8_t* cube = wantCube ? F.outCube.as<uint8_t>() : nullptr;
  const bool useGraph = c->useGraphs && !c->profile && !c->keepIntermediates;
  if (!useGraph) {
    c->kernelProf.on = c->profileKernels;
    t_kernelProf = c->profileKernels ? &c->kernelProf : nullptr;   // launches of this frame are made by this thread
    try {
      enqueueFrame(c, dev.data(), F.outEq.as<uint8_t>(), cube, true);
    } catch (...) {
      t_kernelProf = nullptr;
      throw;
    }
    t_kernelProf = nullptr;
    F.lastWasTimed = true;
    if (F.temporal) F.prevValid = true;
    return;
  }
  const bool usePrevNow = F.temporal && F.prevValid;
  if (!F.graphExec || F.graphHasCube != wantCube || F.graphUsesPrev != usePrevNow) {
    if (F.graphExec) { cudaGraphExecDestroy(F.graphExec); F.graphExec = nullptr; }
    const unsigned long long before = g_launchCount.load();
    cudaGraph_t graph = nullptr;
    S360_CUDA_CHECK(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
    try {
      enqueueFrame(c, dev.data(), F.outEq.as<uint8_t>(), cube, false);
    } catch (...) {
      cudaStreamEndCapture(st, &graph);
      if (graph) cudaGraphDestroy(graph);
      throw;
    }
    S360_CUDA_CHECK(cudaStreamEndCapture(st, &graph));
    F.graphLaunches = g_launchCount.load() - before;
    g_launchCount -= F.graphLaunches;  // nothing has run yet; replays are counted below
    S360_CUDA_CHECK(cudaGraphInstantiate(&F.graphExec, graph, 0));
    cudaGraphDestroy(graph);
    F.graphHasCube = wantCube;
    F.graphUsesPrev = usePrevNow;
  }
  S360_CUDA_CHECK(cudaGraphLaunch(F.graphExec, st));
  g_launchCount += F.graphLaunches;
  F.lastWasTimed = false;
  if (F.temporal) F.prevValid = true;
}

void renderFrameDevice(s360_ctx* c, const s360_params& P, const uint8_t* const* imagesDev, int imgW, int imgH, uint8_t* equirectDev,
                       uint8_t* cubemapDev) {
  if (!planMatches(c, P, imgW, imgH)) buildPlan(c, P, imgW, imgH);
  FramePlan& F = *c->plan;
  cudaStream_t st = c->stream;
  const size_t imgBytes = (size_t)imgW * imgH * 3;
  for (int i = 0; i < F.nRig; ++i)
    if (imagesDev[i]) S360_CUDA_CHECK(cudaMemcpyAsync(F.srcRaw.as<uint8_t>() + (size_t)i * imgBytes, imagesDev[i], imgBytes, cudaMemcpyDeviceToDevice, st));
  int oW, oH, cW, cH;
  frameOutputSize(c, P, &oW, &oH, &cW, &cH);
  F.outEq.ensure((size_t)oW * oH * 3);
  const bool wantCube = cW > 0 && cubemapDev;
  if (wantCube) F.outCube.ensure((size_t)cW * cH * 3);
  runFrameFromStaging(c, wantCube);
  S360_CUDA_CHECK(cudaMemcpyAsync(equirectDev, F.outEq.p, (size_t)oW * oH * 3, cudaMemcpyDeviceToDevice, st));
  if (wantCube) S360_CUDA_CHECK(cudaMemcpyAsync(cubemapDev, F.outCube.p, (size_t)cW * cH * 3, cudaMemcpyDeviceToDevice, st));
}

void renderFrameHost(s360_ctx* c, const s360_params& P, const uint8_t* const* images, int imgW, int imgH, uint8_t* equirect,
                     uint8_t* cubemap) {
  if (!planMatches(c, P, imgW, imgH)) buildPlan(c, P, imgW, imgH);
  FramePlan& F = *c->plan;
  cudaStream_t st = c->stream;
  const size_t imgBytes = (size_t)imgW * imgH * 3;
  for (int i = 0; i < F.nRig; ++i)
    if (images[i]) S360_CUDA_CHECK(cudaMemcpyAsync(F.srcRaw.as<uint8_t>() + (size_t)i * imgBytes, images[i], imgBytes, cudaMemcpyHostToDevice, st));
  int oW, oH, cW, cH;
  frameOutputSize(c, P, &oW, &oH, &cW, &cH);
  F.outEq.ensure((size_t)oW * oH * 3);
  const bool wantCube = cW > 0 && cubemap;
  if (wantCube) F.outCube.ensure((size_t)cW * cH * 3);
  runFrameFromStaging(c, wantCube);
  S360_CUDA_CHECK(cudaMemcpyAsync(equirect, F.outEq.p, (size_t)oW * oH * 3, cudaMemcpyDeviceToHost, st));
  if (wantCube) S360_CUDA_CHECK(cudaMemcpyAsync(cubemap, F.outCube.p, (size_t)cW * cH * 3, cudaMemcpyDeviceToHost, st));
  S360_CUDA_CHECK(cudaStreamSynchronize(st));
  checkFlowErrors(c);
  if (!F.lastWasTimed) return;   // graph replay: no per-stage events
  // stage times of this frame
  FramePlan& G = *c->plan;
  float ms;
  S360_CUDA_CHECK(cudaEventElapsedTime(&ms, G.ev[0], G.ev[6])); c->stageMs[0] = ms;
  for (int i = 0; i < 6; ++i) { S360_CUDA_CHECK(cudaEventElapsedTime(&ms, G.ev[i], G.ev[i + 1])); c->stageMs[i + 1] = ms; }
}

// ------------------------------------------------------------------------------------
// intermediates for stage-wise parity tests
// ------------------------------------------------------------------------------------
int frameIntermediate(s360_ctx* c, const char* name, int idx, void* out, int* w, int* h, int* ch, int* isFloat) {
  if (!c->plan || !c->plan->valid) throw std::runtime_error("no frame has been rendered on this context");
  FramePlan& F = *c->plan;
  const Geometry& g = F.g;
  const std::string n = name;
  cudaStream_t st = c->stream;
  S360_CUDA_CHECK(cudaStreamSynchronize(st));
  const int eqW = F.P.eqr_width, eqH = F.P.eqr_height;
  auto copy2d = [&](const void* src, size_t rowBytes, int rows, size_t srcPitchBytes) {
    if (out) S360_CUDA_CHECK(cudaMemcpy2D(out, rowBytes, src, srcPitchBytes, rowBytes, rows, cudaMemcpyDeviceToHost));
  };
  if (n == "projection") {
    if (idx < 0 || idx >= F.numCams) throw std::runtime_error("projection index out of range");
    *w = g.camW; *h = g.camH; *ch = 4; *isFloat = 0;
    copy2d(F.proj.as<uchar4>() + (size_t)idx * g.camW * g.camH, (size_t)g.camW * 4, g.camH, (size_t)g.camW * 4);
  } else if (n == "flowLtoR" || n == "flowRtoL") {
    if (idx < 0 || idx >= F.numCams) throw std::runtime_error("flow index out of range");
    *w = g.overlapW; *h = g.camH; *ch = 2; *isFloat = 1;
    const size_t per = (size_t)g.overlapW * g.camH;
    copy2d(F.flowSide.as<float2>() + ((n == "flowRtoL" ? F.numCams : 0) + idx) * per, (size_t)g.overlapW * 8, g.camH, (size_t)g.overlapW * 8);
  } else if (n == "panoL" || n == "panoR") {
    *w = eqW; *h = g.camH; *ch = 4; *isFloat = 0;
    const uchar4* p = (n == "panoL" ? F.panoL : F.panoR).as<uchar4>() + (size_t)F.padAbove * eqW;
    copy2d(p, (size_t)eqW * 4, g.camH, (size_t)eqW * 4);
  } else if (n == "topSpherical" || n == "bottomSpherical") {
    if (!F.nPole) throw std::runtime_error("poles are not enabled");
    *w = eqW; *h = g.poleH; *ch = 4; *isFloat = 0;
    copy2d((n == "topSpherical" ? F.topSph : F.botSph).p, (size_t)eqW * 4, g.poleH, (size_t)eqW * 4);
  } else if (n == "bottomMerged") {
    if (!F.poleRemoval) throw std::runtime_error("pole removal is not enabled");
    *w = F.imgW; *h = F.imgH; *ch = 4; *isFloat = 0;
    copy2d(F.prA.p, (size_t)F.imgW * 4, F.imgH, (size_t)F.imgW * 4);
  } else if (n == "poleFlow" || n == "poleWarped") {
    // reference order: top_left, top_right, bottom_left, bottom_right
    int k = idx;
    if (!F.top) k -= 2;
    if (k < 0 || k >= F.nPole) throw std::runtime_error("pole slot not rendered");
    if (n == "poleFlow") {
      *w = g.extW; *h = g.poleH; *ch = 2; *isFloat = 1;
      copy2d(F.poleFlow.as<float2>() + (size_t)k * g.extW * g.poleH, (size_t)g.extW * 8, g.poleH, (size_t)g.extW * 8);
    } else {  // padded to the equirect height like copyMakeBorder (TRSP:538-546)
      *w = eqW; *h = eqH; *ch = 4; *isFloat = 0;
      if (out) {
        std::memset(out, 0, (size_t)eqW * eqH * 4);
        copy2d(F.poleWarped.as<uchar4>() + (size_t)k * eqW * g.poleH, (size_t)eqW * 4, g.poleH, (size_t)eqW * 4);
      }
    }
  } else if (n == "overlapL" || n == "overlapR") {
    // the flow inputs of pair idx (TRSP:196-198): right overlapW columns of projection idx, left overlapW columns of idx+1
    if (idx < 0 || idx >= F.numCams) throw std::runtime_error("pair index out of range");
    *w = g.overlapW; *h = g.camH; *ch = 4; *isFloat = 0;
    const uchar4* src = n == "overlapL" ? F.proj.as<uchar4>() + (size_t)idx * g.camW * g.camH + (g.camW - g.overlapW)
                                         : F.proj.as<uchar4>() + (size_t)((idx + 1) % F.numCams) * g.camW * g.camH;
    copy2d(src, (size_t)g.overlapW * 4, g.camH, (size_t)g.camW * 4);
  } else if (n == "extSide" || n == "extFish") {
    int k = idx;   // reference order: top_left, top_right, bottom_left, bottom_right
    if (!F.top) k -= 2;
    if (k < 0 || k >= F.nPole) throw std::runtime_error("pole slot not rendered");
    *w = g.extW; *h = g.poleH; *ch = 4; *isFloat = 0;
    copy2d((n == "extSide" ? F.extSide : F.extFish).as<uchar4>() + (size_t)k * g.extW * g.poleH, (size_t)g.extW * 4, g.poleH, (size_t)g.extW * 4);
  } else if (n == "eyeL" || n == "eyeR") {
    const int e = n == "eyeL" ? 0 : 1;
    if (!F.eyePre[e].p) throw std::runtime_error("enable keep_intermediates before rendering");
    *w = eqW; *h = eqH; *ch = 4; *isFloat = 0;  // BGRA storage; alpha is not part of the reference image
    copy2d(F.eyePre[e].p, (size_t)eqW * 4, eqH, (size_t)eqW * 4);
  } else {
    throw std::runtime_error("unknown intermediate: " + n);
  }
  return 0;
}

// ------------------------------------------------------------------------------------
// video mode through files (TRSP:200-255, 413-452): the previous frame's state arrives from the host
// ------------------------------------------------------------------------------------
void prepareFrame(s360_ctx* c, const s360_params& P, int imgW, int imgH) {
  if (!planMatches(c, P, imgW, imgH)) buildPlan(c, P, imgW, imgH);
}
static FramePlan& temporalPlan(s360_ctx* c) {
  if (!c->plan) throw std::runtime_error("no frame plan: call s360_frame_prepare (or render a frame) first");
  if (!c->plan->temporal) throw std::runtime_error("temporal mode is off: call s360_ctx_set_temporal before s360_frame_prepare");
  return *c->plan;
}
void temporalSizes(s360_ctx* c, int* overlapW, int* overlapH, int* extW, int* extH) {
  if (!c->plan) throw std::runtime_error("no frame plan: call s360_frame_prepare (or render a frame) first");
  const Geometry& g = c->plan->g;
  *overlapW = g.overlapW; *overlapH = g.camH;
  *extW = c->plan->nPole ? g.extW : 0; *extH = c->plan->nPole ? g.poleH : 0;
}
void temporalImportSide(s360_ctx* c, int pair, const float* flowLtoR, const float* flowRtoL, const uint8_t* overlapL, const uint8_t* overlapR) {
  FramePlan& F = temporalPlan(c);
  const Geometry& g = F.g;
  if (pair < 0 || pair >= F.numCams) throw std::runtime_error("pair index out of range");
  cudaStream_t st = c->stream;
  const size_t per = (size_t)g.overlapW * g.camH;
  float2* pf = F.prevFlowSide.as<float2>();   // [2][numCams]: LtoR then RtoL, like flowSide
  S360_CUDA_CHECK(cudaMemcpyAsync(pf + (size_t)pair * per, flowLtoR, per * 8, cudaMemcpyHostToDevice, st));
  S360_CUDA_CHECK(cudaMemcpyAsync(pf + (size_t)(F.numCams + pair) * per, flowRtoL, per * 8, cudaMemcpyHostToDevice, st));
  // the overlap images are views into the projections (enqueueFrame): L = right columns of `pair`, R = left columns of pair+1
  uchar4* pp = F.prevProj.as<uchar4>();
  const size_t rowB = (size_t)g.overlapW * 4, pitchB = (size_t)g.camW * 4;
  S360_CUDA_CHECK(cudaMemcpy2DAsync(pp + (size_t)pair * g.camW * g.camH + (g.camW - g.overlapW), pitchB, overlapL, rowB, rowB, g.camH,
                                    cudaMemcpyHostToDevice, st));
  S360_CUDA_CHECK(cudaMemcpy2DAsync(pp + (size_t)((pair + 1) % F.numCams) * g.camW * g.camH, pitchB, overlapR, rowB, rowB, g.camH,
                                    cudaMemcpyHostToDevice, st));
  S360_CUDA_CHECK(cudaStreamSynchronize(st));   // the caller may release its buffers on return
}
void temporalImportPole(s360_ctx* c, int slot, const float* flow, const uint8_t* extFish) {
  FramePlan& F = temporalPlan(c);
  const Geometry& g = F.g;
  int k = slot;   // reference order: top_left, top_right, bottom_left, bottom_right
  if (!F.top) k -= 2;
  if (slot < 0 || slot > 3 || k < 0 || k >= F.nPole) throw std::runtime_error("pole slot not rendered");
  cudaStream_t st = c->stream;
  const size_t extPx = (size_t)g.extW * g.poleH;
  S360_CUDA_CHECK(cudaMemcpyAsync(F.prevPoleFlow.as<float2>() + (size_t)k * extPx, flow, extPx * 8, cudaMemcpyHostToDevice, st));
  S360_CUDA_CHECK(cudaMemcpyAsync(F.prevExtFish.as<uchar4>() + (size_t)k * extPx, extFish, extPx * 4, cudaMemcpyHostToDevice, st));
  S360_CUDA_CHECK(cudaStreamSynchronize(st));
}
void temporalCommit(s360_ctx* c) { temporalPlan(c).prevValid = true; }

// ------------------------------------------------------------------------------------
// poleToSideFlowThread as a stand-alone operator (C-ABI)
// ------------------------------------------------------------------------------------
void poleToSide(s360_ctx* c, const s360_params& P, const uint8_t* sideHost, int sw, int sh, const uint8_t* fishHost, int fw, int fh,
                uint8_t* warpedOut, float* flowOut) {
  if (!c->rigLoaded) throw std::runtime_error("no rig loaded");
  if (fw != sw || fh > sh) throw std::runtime_error("pole_to_side: fisheye must be as wide as the side panorama and not taller");
  FlowParams fp;
  if (!flowParamsByName(P.polar_flow_alg, &fp)) throw std::runtime_error(std::string("unrecognized flow algorithm name: ") + P.polar_flow_alg);
  s360_params Pq = P;
  const Geometry g = computeGeometry(c->rig, Pq);
  cudaStream_t st = c->stream;
  const int extW = (int)(float(fw) * 1.2f);
  const size_t sidePx = (size_t)sw * sh, fishPx = (size_t)fw * fh, extPx = (size_t)extW * fh;
  DeviceBuffer dSide(sidePx * 4), dFish(fishPx * 4), dExtS(extPx * 4), dExtF(extPx * 4), dScr(featherScratchBytes(fw, fh)), dFlow(extPx * 8),
      dWarp(sidePx * 4);
  S360_CUDA_CHECK(cudaMemcpyAsync(dSide.p, sideHost, sidePx * 4, cudaMemcpyHostToDevice, st));
  S360_CUDA_CHECK(cudaMemcpyAsync(dFish.p, fishHost, fishPx * 4, cudaMemcpyHostToDevice, st));
  S360_CUDA_CHECK(cudaMemsetAsync(dWarp.p, 0, sidePx * 4, st));
  launchFeatherExtend(st, dSide.as<uchar4>(), sw, sh, fh, 0, P.std_alpha_feather_size, c->blurKernelFor(P.std_alpha_feather_size),
                      dExtS.as<uchar4>(), extW, dScr.as<uint8_t>());
  launchExtendWrap(st, dFish.as<uchar4>(), fw, fh, dExtF.as<uchar4>(), extW);
  FlowBatch fb(1, extW, fh, fp);
  fb.setLatencyMode(c->latencyMode);
  FlowImagePtrs ip{};
  FlowPrevPtrs pv{};
  ip.i0[0] = dExtS.as<uchar4>(); ip.i1[0] = dExtF.as<uchar4>(); ip.pitch0[0] = ip.pitch1[0] = extW; ip.hint[0] = S360_HINT_DOWN;
  fb.run(st, ip, pv, dFlow.as<float2>());
  PoleWarpArgs a{};
  a.extFish = dExtF.as<uchar4>(); a.flow = dFlow.as<float2>(); a.cols = fw; a.rows = fh; a.extW = extW;
  a.poleCameraRadiusDeg = g.poleCameraRadiusDeg; a.phiRampStart = g.phiRampStart; a.phiMid = g.phiMid; a.phiRampEnd = g.phiRampEnd;
  a.warped = dWarp.as<uchar4>();  // first fh rows of the sw x sh output
  launchPoleWarp(st, a, c->rt);
  S360_CUDA_CHECK(cudaMemcpyAsync(warpedOut, dWarp.p, sidePx * 4, cudaMemcpyDeviceToHost, st));
  if (flowOut) S360_CUDA_CHECK(cudaMemcpyAsync(flowOut, dFlow.p, extPx * 8, cudaMemcpyDeviceToHost, st));
  S360_CUDA_CHECK(cudaStreamSynchronize(st));
}

}  // namespace s360

namespace s360 {
void resetPlan(s360_ctx* c) { c->plan.reset(); }
void resetTemporal(s360_ctx* c) { if (c->plan) c->plan->prevValid = false; }
void checkFlowErrors(s360_ctx* c) {   // call after the context's stream has been synchronised
  if (!c->plan) return;
  if (c->plan->sideFlow) c->plan->sideFlow->checkSweepError();
  if (c->plan->poleFlowBatch) c->plan->poleFlowBatch->checkSweepError();
  if (c->plan->prFlowBatch) c->plan->prFlowBatch->checkSweepError();
}
void collectProfile(s360_ctx* c) {
  S360_CUDA_CHECK(cudaStreamSynchronize(c->stream));
  if (c->plan && c->plan->sideFlow) c->plan->sideFlow->collectProfile(&c->sweepProfile);
  if (c->plan && c->plan->poleFlowBatch) c->plan->poleFlowBatch->collectProfile(&c->sweepProfile);
  c->kernelProf.collect();   // also without a frame plan: the ISP and the stand-alone operators profile through the same scopes
}
}
s360_ctx::s360_ctx() {}
s360_ctx::~s360_ctx() {}

const int* s360_ctx::blurKernelFor(int erodeSize) {
  auto it = blurKernels.find(erodeSize);
  if (it == blurKernels.end()) {
    if (erodeSize < 1 || (erodeSize & 1) == 0) throw std::runtime_error("std_alpha_feather_size must be odd");
    std::vector<int> k = s360::gaussKernelFixed8(erodeSize, erodeSize / 2.0f);
    auto buf = std::make_unique<s360::DeviceBuffer>(k.size() * 4);
    S360_CUDA_CHECK(cudaMemcpy(buf->p, k.data(), k.size() * 4, cudaMemcpyHostToDevice));
    it = blurKernels.emplace(erodeSize, std::move(buf)).first;
  }
  return it->second->as<int>();
}
