// ORACLE — TEST INFRASTRUCTURE ONLY (see ocv.h header). C ABI used by tests/ via ctypes.
#include "ocv.h"
using namespace orc;
extern "C" {

static ImgF wrapF(const float* p, int w, int h, int c) { ImgF m(w, h, c); memcpy(m.d.data(), p, m.d.size() * 4); return m; }
static Img8 wrap8(const uint8_t* p, int w, int h, int c) { Img8 m(w, h, c); memcpy(m.d.data(), p, m.d.size()); return m; }

void o_gauss_kernel(int n, double sigma, float* out) { auto k = gaussKernel(n, sigma); memcpy(out, k.data(), n * 4); }
void o_gaussian_blur_f32(const float* s, int w, int h, int c, int k, double sigma, float* d) {
  ImgF a = wrapF(s, w, h, c), b; gaussianBlurF(a, b, k, sigma); memcpy(d, b.d.data(), b.d.size() * 4);
}
void o_sobel1(const float* s, int w, int h, int dx, float* d) { ImgF a = wrapF(s, w, h, 1), b; sobel1(a, b, dx != 0); memcpy(d, b.d.data(), b.d.size() * 4); }
void o_median5(const float* s, int w, int h, int c, float* d) { ImgF a = wrapF(s, w, h, c), b; median5(a, b); memcpy(d, b.d.data(), b.d.size() * 4); }
void o_resize_linear_f32(const float* s, int w, int h, int c, int dw, int dh, float* d) {
  ImgF a = wrapF(s, w, h, c), b; resizeLinearF(a, b, dw, dh); memcpy(d, b.d.data(), b.d.size() * 4);
}
void o_resize_cubic_f32(const float* s, int w, int h, int c, int dw, int dh, float* d) {
  ImgF a = wrapF(s, w, h, c), b; resizeCubicF(a, b, dw, dh); memcpy(d, b.d.data(), b.d.size() * 4);
}
void o_resize_cubic_u8(const uint8_t* s, int w, int h, int c, int dw, int dh, uint8_t* d) {
  Img8 a = wrap8(s, w, h, c), b; resizeCubic8(a, b, dw, dh); memcpy(d, b.d.data(), b.d.size());
}
void o_remap_cubic_u8(const uint8_t* s, int w, int h, int c, const float* map, int dw, int dh, int border, uint8_t* d) {
  Img8 a = wrap8(s, w, h, c), b; ImgF m = wrapF(map, dw, dh, 2); remapCubic8(a, m, b, border); memcpy(d, b.d.data(), b.d.size());
}
void o_remap_cubic_f32(const float* s, int w, int h, int c, const float* map, int dw, int dh, float* d) {
  ImgF a = wrapF(s, w, h, c), b; ImgF m = wrapF(map, dw, dh, 2); remapCubicF(a, m, b); memcpy(d, b.d.data(), b.d.size() * 4);
}
void o_bgra2gray(const uint8_t* s, int n, int c, int legacy, uint8_t* d) { for (int i = 0; i < n; ++i) d[i] = gray8(s[i * c], s[i * c + 1], s[i * c + 2], legacy != 0); }
void o_erode_cross(const uint8_t* s, int w, int h, int r, uint8_t* d) { Img8 a = wrap8(s, w, h, 1), b; erodeCross8(a, b, r); memcpy(d, b.d.data(), b.d.size()); }
void o_gauss_kernel_fixed8(int n, double sigma, int* out) { auto k = gaussKernelFixed8(n, sigma); memcpy(out, k.data(), n * 4); }
void o_gaussian_blur_u8(const uint8_t* s, int w, int h, int k, double sigma, uint8_t* d) {
  Img8 a = wrap8(s, w, h, 1), b; gaussianBlur8Fixed(a, b, gaussKernelFixed8(k, sigma)); memcpy(d, b.d.data(), b.d.size());
}
int o_offset_wrap_index(int x, int cols, float offset) { return offsetWrapIndex(x, cols, offset); }
}

// ---------------- camera / rig / flow / frame level ----------------
#include "orender.h"
extern "C" {

struct ORenderParams {  // mirrors orc::RenderParams (POD for ctypes)
  double interpupilary_dist, sharpening, zero_parallax_dist;
  int side_alpha_feather_size, std_alpha_feather_size;
  int enable_top, enable_bottom;
  int eqr_width, eqr_height, final_eqr_width, final_eqr_height, cubemap_width, cubemap_height;
  int want_cubemap, cubemap_photo, threads, keep_intermediates;
  int side_search20, polar_search20;
  int enable_pole_removal;
};
static RenderParams toParams(const ORenderParams* p) {
  RenderParams P;
  P.interpupilary_dist = p->interpupilary_dist; P.sharpening = p->sharpening; P.zero_parallax_dist = p->zero_parallax_dist;
  P.side_alpha_feather_size = p->side_alpha_feather_size; P.std_alpha_feather_size = p->std_alpha_feather_size;
  P.enable_top = p->enable_top; P.enable_bottom = p->enable_bottom;
  P.eqr_width = p->eqr_width; P.eqr_height = p->eqr_height; P.final_eqr_width = p->final_eqr_width;
  P.final_eqr_height = p->final_eqr_height; P.cubemap_width = p->cubemap_width; P.cubemap_height = p->cubemap_height;
  P.want_cubemap = p->want_cubemap; P.cubemap_format = p->cubemap_photo ? "photo" : "video";
  P.threads = p->threads; P.keep_intermediates = p->keep_intermediates;
  P.side_flow_alg = p->side_search20 ? "pixflow_search_20" : "pixflow_low";
  P.polar_flow_alg = p->polar_search20 ? "pixflow_search_20" : "pixflow_low";
  P.enable_pole_removal = p->enable_pole_removal != 0;
  return P;
}

static thread_local std::string g_err;
const char* o_last_error() { return g_err.c_str(); }

void* o_rig_from_json(const char* json) {
  try { return new Rig(Rig::fromString(json)); } catch (std::exception& e) { g_err = e.what(); return nullptr; }
}
void o_rig_free(void* r) { delete (Rig*)r; }
int o_rig_num_cameras(void* r) { return (int)((Rig*)r)->rig.size(); }
int o_rig_num_side(void* r) { return (int)((Rig*)r)->side.size(); }
// which: 0 top, 1 bottom, 2 bottom2 ; returns index into rig.rig
int o_rig_special_camera(void* rp, int which) {
  Rig* r = (Rig*)rp;
  const Camera* c = which == 0 ? &r->findByDirection(V3(0, 0, 1)) : which == 1 ? &r->findByDirection(V3(0, 0, -1)) : &r->findLargestAxisDist();
  return (int)(c - r->rig.data());
}
int o_rig_side_index(void* rp, int sideIdx) {  // index into rig.rig of side camera sideIdx
  Rig* r = (Rig*)rp;
  for (size_t i = 0; i < r->rig.size(); ++i) if (r->rig[i].id == r->side[sideIdx].id) return (int)i;
  return -1;
}
const char* o_camera_id(void* rp, int cam) { return ((Rig*)rp)->rig[cam].id.c_str(); }
void o_camera_pixel(void* rp, int cam, const double* xyz, double* out) {
  V2 p = ((Rig*)rp)->rig[cam].pixel(V3(xyz[0], xyz[1], xyz[2])); out[0] = p.x; out[1] = p.y;
}
void o_camera_rig_ray(void* rp, int cam, const double* px, double* out6) {
  Ray r = ((Rig*)rp)->rig[cam].rig(V2{px[0], px[1]});
  out6[0] = r.origin.x; out6[1] = r.origin.y; out6[2] = r.origin.z; out6[3] = r.dir.x; out6[4] = r.dir.y; out6[5] = r.dir.z;
}
// fills: position(3) rotation(9 row-major) resolution(2) principal(2) distortion(2) focal(2) fovThreshold type
void o_camera_params(void* rp, int cam, double* o) {
  const Camera& c = ((Rig*)rp)->rig[cam];
  o[0] = c.position.x; o[1] = c.position.y; o[2] = c.position.z;
  for (int i = 0; i < 9; ++i) o[3 + i] = c.rotation.m[i / 3][i % 3];
  o[12] = c.resolution.x; o[13] = c.resolution.y; o[14] = c.principal.x; o[15] = c.principal.y;
  o[16] = c.distortion.x; o[17] = c.distortion.y; o[18] = c.focal.x; o[19] = c.focal.y; o[20] = c.fovThreshold; o[21] = c.type;
}
double o_camera_get_fov(void* rp, int cam) { return ((Rig*)rp)->rig[cam].getFov(); }
void o_camera_set_fov(void* rp, int cam, double fov) { ((Rig*)rp)->rig[cam].setFov(fov); }
void o_camera_set_default_fov(void* rp, int cam) { ((Rig*)rp)->rig[cam].setDefaultFov(); }
int o_camera_sees(void* rp, int cam, const double* xyz) { return ((Rig*)rp)->rig[cam].sees(V3(xyz[0], xyz[1], xyz[2])); }
void o_camera_set_distortion(void* rp, int cam, double d0, double d1) { ((Rig*)rp)->rig[cam].distortion = V2{d0, d1}; }
double o_camera_distort(void* rp, int cam, double r) { return ((Rig*)rp)->rig[cam].distort(r); }
double o_camera_undistort(void* rp, int cam, double d) { return ((Rig*)rp)->rig[cam].undistort(d); }
float o_approximate_fov(void* rp, int vertical) { return approximateFov(((Rig*)rp)->side, vertical != 0); }
float o_ring_radius(void* rp) { return ((Rig*)rp)->ringRadius(); }
float o_usable_radius(void* rp, int cam) { return Camera::approximateUsablePixelsRadius(((Rig*)rp)->rig[cam]); }

// geometry: fills ints camW camH overlapW numNovel, floats hRad vRad fovHRad slabDisp shiftPx
void o_side_geometry(void* rp, const ORenderParams* p, int* oi, float* of) {
  SideGeometry g = sideGeometry(*(Rig*)rp, toParams(p));
  oi[0] = g.camW; oi[1] = g.camH; oi[2] = g.overlapW; oi[3] = g.numNovel;
  of[0] = g.hRadians; of[1] = g.vRadians; of[2] = g.fovHRad; of[3] = g.slabDisp; of[4] = g.shiftPx;
}

void o_spherical_warp_map(void* rp, int cam, int w, int h, float l, float r, float t, float b, float* out) {
  ImgF m; sphericalWarpMap(m, w, h, ((Rig*)rp)->rig[cam], l, r, t, b); memcpy(out, m.d.data(), m.d.size() * 4);
}
void o_project_side(void* rp, int cam, const uint8_t* src, int sw, int sh, int sc, int dw, int dh, float l, float r, float t,
                    float b, int feather, uint8_t* out) {
  Img8 d = projectSideToSpherical(dw, dh, wrap8(src, sw, sh, sc), ((Rig*)rp)->rig[cam], l, r, t, b, feather);
  memcpy(out, d.d.data(), d.d.size());
}
void o_prepare_pole(void* rp, const ORenderParams* p, const uint8_t* srcBGR, int sw, int sh, int top, uint8_t* out, int* ow, int* oh) {
  Img8 d = preparePoleSpherical(*(Rig*)rp, toParams(p), wrap8(srcBGR, sw, sh, 3), top != 0);
  *ow = d.w; *oh = d.h; if (out) memcpy(out, d.d.data(), d.d.size());
}

// PixFlow: rgba0/1 BGRA8 w*h; prev* optional (null). levels: optional per-level dump via o_pixflow_levels
static std::vector<ImgF> g_levels;
int o_pixflow(const uint8_t* i0, const uint8_t* i1, int w, int h, const float* prevFlow, const uint8_t* prevI0,
              const uint8_t* prevI1, int search20, int hint, int dumpLevels, float* outFlow) {
  try {
    PixFlow pf(flowParamsByName(search20 ? "pixflow_search_20" : "pixflow_low"));
    g_levels.clear();
    if (dumpLevels) pf.levelDump = &g_levels;
    ImgF flow, pfw; Img8 p0, p1;
    if (prevFlow) { pfw = wrapF(prevFlow, w, h, 2); p0 = wrap8(prevI0, w, h, 4); p1 = wrap8(prevI1, w, h, 4); }
    pf.computeOpticalFlow(wrap8(i0, w, h, 4), wrap8(i1, w, h, 4), pfw, p0, p1, flow, hint);
    memcpy(outFlow, flow.d.data(), flow.d.size() * 4);
    return 0;
  } catch (std::exception& e) { g_err = e.what(); return 1; }
}
int o_pixflow_num_levels() { return (int)g_levels.size(); }
void o_pixflow_level(int i, int* w, int* h, float* out) { *w = g_levels[i].w; *h = g_levels[i].h; if (out) memcpy(out, g_levels[i].d.data(), g_levels[i].d.size() * 4); }

// one sweep-level (patchMatchPropagationAndSearch) on given float inputs, for kernel-level parity
void o_pixflow_level_solve(const float* I0, const float* I1, const float* a0, const float* a1, int w, int h, int hasFlow,
                           float* flowInOut, int search20, int hint) {
  PixFlow pf(flowParamsByName(search20 ? "pixflow_search_20" : "pixflow_low"));
  ImgF flow; if (hasFlow) flow = wrapF(flowInOut, w, h, 2);
  pf.patchMatchPropagationAndSearch(wrapF(I0, w, h, 1), wrapF(I1, w, h, 1), wrapF(a0, w, h, 1), wrapF(a1, w, h, 1), flow, hint);
  memcpy(flowInOut, flow.d.data(), flow.d.size() * 4);
}

void o_novel_view_chunk(const uint8_t* ovL, const uint8_t* ovR, const float* fLR, const float* fRL, int ow, int oh,
                        int bufW, int camW, int numNovel, float slabDisp, uint8_t* outL, uint8_t* outR) {
  LazyWarp wl, wr; wl.x.assign(bufW, 0.f); wl.t.assign(bufW, -1.f); wr = wl;
  for (int nv = 0; nv < numNovel && nv < bufW; ++nv) {
    const float shift = float(nv) / float(numNovel);
    const float slabShift = float(camW) * 0.5f - float(numNovel - nv);
    wl.x[nv] = slabShift + slabDisp; wl.t[nv] = shift; wr.x[nv] = slabShift - slabDisp; wr.t[nv] = shift;
  }
  Img8 L = wrap8(ovL, ow, oh, 4), Rr = wrap8(ovR, ow, oh, 4); ImgF flr = wrapF(fLR, ow, oh, 2), frl = wrapF(fRL, ow, oh, 2);
  Img8 a, b, c, d; ImgF ma, mb, mc, md;
  renderLazyNovelView(bufW, oh, wl, L, frl, false, a, ma); renderLazyNovelView(bufW, oh, wl, Rr, flr, true, b, mb);
  renderLazyNovelView(bufW, oh, wr, L, frl, false, c, mc); renderLazyNovelView(bufW, oh, wr, Rr, flr, true, d, md);
  Img8 cl = combineLazyViews(a, b, ma, mb), cr = combineLazyViews(c, d, mc, md);
  memcpy(outL, cl.d.data(), cl.d.size()); memcpy(outR, cr.d.data(), cr.d.size());
}
void o_flatten_deghost(const uint8_t* base, const uint8_t* top, int w, int h, uint8_t* out) {
  Img8 d = flattenLayersDeghostPreferBase(wrap8(base, w, h, 4), wrap8(top, w, h, 4)); memcpy(out, d.d.data(), d.d.size());
}
void o_feather_alpha(const uint8_t* src, int w, int h, int size, uint8_t* out) {
  Img8 d = featherAlphaChannel(wrap8(src, w, h, 4), size); memcpy(out, d.d.data(), d.d.size());
}
void o_offset_wrap(const uint8_t* src, int w, int h, int c, float offset, uint8_t* out) {
  Img8 d = offsetHorizontalWrap(wrap8(src, w, h, c), offset); memcpy(out, d.d.data(), d.d.size());
}
void o_sharpen(uint8_t* img, int w, int h, double sharpening) {
  Img8 m = wrap8(img, w, h, 3); sharpenEye(m, sharpening); memcpy(img, m.d.data(), m.d.size());
}
void o_iir_lowpass(const uint8_t* img, int w, int h, float amount, uint8_t* out) {
  Img8 lp; iirLowPass(wrap8(img, w, h, 3), amount, lp); memcpy(out, lp.d.data(), lp.d.size());
}
void o_cubemap(const uint8_t* eye, int w, int h, int fw, int fh, int photo, uint8_t* out) {
  Img8 c = stackCubemapFaces(photo ? "photo" : "video", sphericalToCubemap(wrap8(eye, w, h, 3), (float)M_PI, fw, fh));
  memcpy(out, c.d.data(), c.d.size());
}
void o_pole_to_side(void* rp, const ORenderParams* p, const uint8_t* side, int sw, int sh, const uint8_t* fish, int fw, int fh,
                    uint8_t* warpedOut, float* flowOut) {
  Img8 w; ImgF f;
  poleToSideFlow(*(Rig*)rp, toParams(p), wrap8(side, sw, sh, 4), wrap8(fish, fw, fh, 4), w, &f);
  memcpy(warpedOut, w.d.data(), w.d.size()); if (flowOut) memcpy(flowOut, f.d.data(), f.d.size() * 4);
}

// full frame. images: array of pointers to BGR frames (w*h*3) in rig order.
void* o_state_new() { return new FrameState(); }
void o_state_free(void* s) { delete (FrameState*)s; }
void* o_render_frame(void* rp, const uint8_t** images, int iw, int ih, const ORenderParams* p, const uint8_t* mask1, const uint8_t* mask2, void* state) {
  try {
    Rig* r = (Rig*)rp;
    std::vector<Img8> im;
    for (size_t i = 0; i < r->rig.size(); ++i) im.push_back(wrap8(images[i], iw, ih, 3));
    Img8 m1, m2;
    if (mask1 && mask2) { m1 = wrap8(mask1, iw, ih, 3); m2 = wrap8(mask2, iw, ih, 3); }
    return new FrameResult(renderFrame(*r, im, toParams(p), mask1 ? &m1 : nullptr, mask2 ? &m2 : nullptr, (FrameState*)state));
  } catch (std::exception& e) { g_err = e.what(); return nullptr; }
}
void o_result_free(void* f) { delete (FrameResult*)f; }
// name -> buffer. returns 0 ok. type: 0 u8, 1 f32
int o_result_get(void* fp, const char* name, int idx, const void** ptr, int* w, int* h, int* c, int* isFloat) {
  FrameResult* f = (FrameResult*)fp; std::string n = name;
  const Img8* a = nullptr; const ImgF* b = nullptr;
  if (n == "equirect") a = &f->equirect; else if (n == "cubemap") a = &f->cubemap;
  else if (n == "projection") a = &f->projections.at(idx);
  else if (n == "flowLtoR") b = &f->flowLtoR.at(idx); else if (n == "flowRtoL") b = &f->flowRtoL.at(idx);
  else if (n == "panoL") a = &f->panoL; else if (n == "panoR") a = &f->panoR;
  else if (n == "bottomMerged") a = &f->bottomMerged; else if (n == "topSpherical") a = &f->topSpherical; else if (n == "bottomSpherical") a = &f->bottomSpherical;
  else if (n == "poleFlow") b = &f->poleFlow[idx]; else if (n == "poleWarped") a = &f->poleWarped[idx];
  else if (n == "eyeL") a = &f->eyeL; else if (n == "eyeR") a = &f->eyeR;
  else return 1;
  if (a) { *ptr = a->d.data(); *w = a->w; *h = a->h; *c = a->c; *isFloat = 0; }
  else { *ptr = b->d.data(); *w = b->w; *h = b->h; *c = b->c; *isFloat = 1; }
  return 0;
}
void o_result_times(void* fp, double* t) {
  FrameResult* f = (FrameResult*)fp; t[0] = f->tTotal; t[1] = f->tProject; t[2] = f->tFlow; t[3] = f->tNovel; t[4] = f->tPole; t[5] = f->tSharpen;
}
}
