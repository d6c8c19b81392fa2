// ORACLE — TEST INFRASTRUCTURE ONLY. Not part of the product path.
//
// C entry points over the REFERENCE'S OWN code (compiled unchanged from /root/reference by oracle/ref_build.sh into
// oracle/_ref/libref.so) so that tests can put the reference, the oracle restatement (oracle/*.h) and the CUDA path
// side by side, operator by operator and for a whole frame. Nothing here restates an algorithm: every function
// converts raw buffers to cv::Mat (the stand-in of oracle/ref_shim/) and calls the reference.
//
// The reference's frame renderer lives in its main program (test/TestRenderStereoPanorama.cpp: static functions and
// gflags globals); that translation unit is included below with `main` renamed, so renderStereoPanorama(),
// poleToSideFlowThread(), sharpenThread() ... are called exactly as the binary would call them.
#include <execinfo.h>
#include <signal.h>
#include <sys/stat.h>
#include <unistd.h>

#include <cstdio>
#include <cstring>
#include <mutex>
#include <string>

#define main ref_trsp_main
#include "test/TestRenderStereoPanorama.cpp"
#undef main

namespace {
std::string gErr;
std::mutex gFrameMutex;   // the reference's flags are process globals: one frame at a time

Mat wrap8(const uint8_t* p, int w, int h, int c) {
  Mat m(h, w, CV_MAKETYPE(CV_8U, c));
  std::memcpy(m.data, p, (size_t)w * h * c);
  return m;
}
Mat wrapF(const float* p, int w, int h, int c) {
  Mat m(h, w, CV_MAKETYPE(CV_32F, c));
  std::memcpy(m.data, p, (size_t)w * h * c * sizeof(float));
  return m;
}
void out8(const Mat& m, uint8_t* p) {
  const size_t rb = (size_t)m.cols * m.elemSize();
  for (int y = 0; y < m.rows; ++y) std::memcpy(p + rb * y, m.ptr(y), rb);
}
void outF(const Mat& m, float* p) { out8(m, (uint8_t*)p); }

template <class F>
int guarded(F f) {
  try {
    f();
    return 0;
  } catch (const std::exception& e) {
    gErr = e.what();
    return 1;
  }
}
}  // namespace

extern "C" {

const char* ref_last_error() { return gErr.c_str(); }

// debugging aid: print a native backtrace on SIGSEGV (the reference's own printStacktrace, SystemUtil.cpp:29-38)
static void refCrash(int) {
  void* stack[64];
  const int n = backtrace(stack, 64);
  backtrace_symbols_fd(stack, n, 2);
  _exit(139);
}
void ref_install_crash_handler() {
  static char altStack[1 << 16];
  stack_t ss;
  ss.ss_sp = altStack; ss.ss_size = sizeof altStack; ss.ss_flags = 0;
  sigaltstack(&ss, nullptr);
  struct sigaction sa;
  std::memset(&sa, 0, sizeof sa);
  sa.sa_handler = refCrash;
  sa.sa_flags = SA_ONSTACK;
  sigaction(SIGSEGV, &sa, nullptr);
}

// Camera::unitTest() (render/Camera.cpp:291-412); aborts through CHECK on failure like the reference
int ref_camera_unit_test() { return guarded([] { Camera::unitTest(); }); }

// Camera::pixel through a rig file: out[2]
int ref_camera_pixel(const char* rigJson, int cam, const double* xyz, double* out) {
  return guarded([&] {
    Camera::Rig rig = Camera::loadRig(rigJson);
    Camera::Vector2 p = rig.at(cam).pixel(Camera::Vector3(xyz[0], xyz[1], xyz[2]));
    out[0] = p.x();
    out[1] = p.y();
  });
}

// RigDescription + the sizing formulas' inputs: ids of the special cameras, ring radius, fov (TRSP:75-97)
int ref_rig_summary(const char* rigJson, int* nSide, float* ringRadius, float* hFov, float* vFov, double* poleFov, char* topId, char* bottomId, char* bottom2Id) {
  return guarded([&] {
    RigDescription rig(rigJson);
    *nSide = rig.getSideCameraCount();
    *ringRadius = rig.getRingRadius();
    *hFov = approximateFov(rig.rigSideOnly, false);
    *vFov = approximateFov(rig.rigSideOnly, true);
    *poleFov = rig.findCameraByDirection(-kGlobalUp).getFov();
    std::strcpy(topId, rig.getTopCameraId().c_str());
    std::strcpy(bottomId, rig.getBottomCameraId().c_str());
    std::strcpy(bottom2Id, rig.getBottomCamera2Id().c_str());
  });
}

// OpticalFlowInterface::computeOpticalFlow via makeOpticalFlowByName (OpticalFlowFactory.h:23-64). prev* may be null.
int ref_pixflow(const uint8_t* i0, const uint8_t* i1, int w, int h, const float* prevFlow, const uint8_t* prevI0, const uint8_t* prevI1,
                const char* alg, int hint, float* out) {
  return guarded([&] {
    Mat m0 = wrap8(i0, w, h, 4), m1 = wrap8(i1, w, h, 4), pf, p0, p1, flow;
    if (prevFlow) { pf = wrapF(prevFlow, w, h, 2); p0 = wrap8(prevI0, w, h, 4); p1 = wrap8(prevI1, w, h, 4); }
    OpticalFlowInterface* alg_ = makeOpticalFlowByName(alg);
    alg_->computeOpticalFlow(m0, m1, pf, p0, p1, flow, OpticalFlowInterface::DirectionHint(hint));
    delete alg_;
    outF(flow, out);
  });
}

// NovelViewGeneratorAsymmetricFlow::prepare (both flows, NovelView.cpp:270-299), timed like TestOpticalFlow.cpp:70-81
int ref_prepare_pair(const uint8_t* l, const uint8_t* r, int w, int h, const char* alg, float* flowLtoR, float* flowRtoL, double* seconds) {
  return guarded([&] {
    NovelViewGeneratorAsymmetricFlow gen(alg);
    Mat ml = wrap8(l, w, h, 4), mr = wrap8(r, w, h, 4);
    const double t0 = getCurrTimeSec();
    static_cast<NovelViewGenerator&>(gen).prepare(ml, mr);
    *seconds = getCurrTimeSec() - t0;
    outF(gen.getFlowLtoR(), flowLtoR);
    outF(gen.getFlowRtoL(), flowRtoL);
  });
}

// renderStereoPanoramaChunksThread (TRSP:259-292) on given overlap images and flows
int ref_novel_view_chunk(const uint8_t* ovL, const uint8_t* ovR, const float* fLR, const float* fRL, int w, int h, int eqrWidth, int numCams,
                         int camW, int numNovel, float fovHRad, float slabDisp, uint8_t* outL, uint8_t* outR) {
  return guarded([&] {
    NovelViewGeneratorLazyFlow* gen = new NovelViewGeneratorAsymmetricFlow("pixflow_low");
    gen->imageL = wrap8(ovL, w, h, 4);
    gen->imageR = wrap8(ovR, w, h, 4);
    gen->flowLtoR = wrapF(fLR, w, h, 2);
    gen->flowRtoL = wrapF(fRL, w, h, 2);
    const int saved = FLAGS_eqr_width;
    FLAGS_eqr_width = eqrWidth;
    Mat cl, cr;
    renderStereoPanoramaChunksThread(0, numCams, camW, h, numNovel, fovHRad, slabDisp, gen, &cl, &cr);
    FLAGS_eqr_width = saved;
    delete gen;
    out8(cl, outL);
    out8(cr, outR);
  });
}

int ref_flatten_deghost(const uint8_t* base, const uint8_t* top, int w, int h, uint8_t* out) {
  return guarded([&] { out8(flattenLayersDeghostPreferBase(wrap8(base, w, h, 4), wrap8(top, w, h, 4)), out); });
}
int ref_feather_alpha(const uint8_t* src, int w, int h, int size, uint8_t* out) {
  return guarded([&] { out8(featherAlphaChannel(wrap8(src, w, h, 4), size), out); });
}
int ref_offset_wrap(const uint8_t* src, int w, int h, int c, float offset, uint8_t* out) {
  return guarded([&] { out8(offsetHorizontalWrap(wrap8(src, w, h, c), offset), out); });
}
// sharpenThread (TRSP:688-696) on a 3-channel image, in place
int ref_sharpen(uint8_t* img, int w, int h, double sharpening) {
  return guarded([&] {
    Mat m = wrap8(img, w, h, 3);
    const double saved = FLAGS_sharpening;
    FLAGS_sharpening = sharpening;
    sharpenThread(&m);
    FLAGS_sharpening = saved;
    out8(m, img);
  });
}
// one eye's cubemap (TRSP:917-935 per eye): out is 3fw x 2fh ("video") or fw x 6fh ("photo")
int ref_cubemap(const uint8_t* eye, int w, int h, int fw, int fh, int photo, uint8_t* out) {
  return guarded([&] {
    out8(stackOutputCubemapFaces(photo ? "photo" : "video", convertSphericalToCubemapBicubicRemap(wrap8(eye, w, h, 3), M_PI, fw, fh)), out);
  });
}
// projectSideToSpherical (TRSP:99-135): src has 3 or 4 channels, dst is dw x dh BGRA
int ref_project_side(const char* rigJson, int sideIdx, const uint8_t* src, int sw, int sh, int sc, int dw, int dh, float l, float r, float t, float b,
                     int feather, uint8_t* out) {
  return guarded([&] {
    RigDescription rig(rigJson);
    Mat dst(dh, dw, CV_8UC4), s = wrap8(src, sw, sh, sc);
    const int saved = FLAGS_side_alpha_feather_size;
    FLAGS_side_alpha_feather_size = feather;
    projectSideToSpherical(dst, s, rig.rigSideOnly.at(sideIdx), l, r, t, b);
    FLAGS_side_alpha_feather_size = saved;
    out8(dst, out);
  });
}

// The whole frame through the reference's renderStereoPanorama() (TRSP:716-972). Camera images travel through the
// in-memory hook of the imread/imwrite stand-in (no PNG encode/decode); flow .bin files are written to workDir by the
// reference's own saveFlowToFile. images[i] is the i-th camera of the rig file (BGR, imgW x imgH).
// equirect: final_eqr_width x final_eqr_height x 3 (or eqr_width x 2*eqr_height when the resize guard of TRSP:938-941
// skips); cubemap may be null. seconds[0] = wall time of renderStereoPanorama().
struct RefFrameParams {
  double interpupilary_dist, sharpening, zero_parallax_dist;
  int side_alpha_feather_size, std_alpha_feather_size, enable_top, enable_bottom, eqr_width, eqr_height, final_eqr_width, final_eqr_height,
      cubemap_width, cubemap_height, want_cubemap, cubemap_photo, enable_pole_removal, side_search20, polar_search20;
};
int ref_render_frame(const char* rigJson, const uint8_t* const* images, int nImages, int imgW, int imgH, const RefFrameParams* P, const char* workDir,
                     const char* prevFrame, const char* frameNumber, const uint8_t* poleMask1, const uint8_t* poleMask2, uint8_t* equirect,
                     int* eqW, int* eqH, uint8_t* cubemap, int* cubeW, int* cubeH, double* seconds) {
  std::lock_guard<std::mutex> lock(gFrameMutex);
  return guarded([&] {
    const std::string work = workDir, frame = frameNumber;
    const std::string imgs = work + "/imgs", outDir = work + "/out", masks = work + "/masks";
    auto mk = [](const std::string& d) { mkdir(d.c_str(), 0777); };
    mk(work); mk(imgs); mk(outDir); mk(masks); mk(outDir + "/flow"); mk(outDir + "/flow/" + frame); mk(outDir + "/debug");
    mk(outDir + "/debug/" + frame); mk(outDir + "/debug/" + frame + "/flow_images");
    if (!std::getenv("REF_VERBOSE")) google::refShimMinLogLevel() = 1;   // LOG(INFO) off; warnings and errors stay
    cv::refshim::memClear();
    Camera::Rig rig = Camera::loadRig(rigJson);
    if ((int)rig.size() != nImages) throw VrCamException("ref_render_frame: image count does not match the rig");
    for (int i = 0; i < nImages; ++i) {
      const std::string dir = imgs + "/" + rig[i].id, path = dir + "/" + frame + ".png";
      mk(dir);
      FILE* f = std::fopen(path.c_str(), "ab");   // placeholder: the extension is sniffed from the directory (SystemUtil.h:96-105)
      if (f) std::fclose(f);
      cv::refshim::memPut(path, wrap8(images[i], imgW, imgH, 3));
    }
    RigDescription rd(rigJson);
    if (poleMask1 && poleMask2) {
      cv::refshim::memPut(masks + "/" + rd.getBottomCameraId() + ".png", wrap8(poleMask1, imgW, imgH, 3));
      cv::refshim::memPut(masks + "/" + rd.getBottomCamera2Id() + ".png", wrap8(poleMask2, imgW, imgH, 3));
    }
    cv::refshim::memSinkPrefix(outDir);
    FLAGS_rig_json_file = rigJson;
    FLAGS_imgs_dir = imgs;
    FLAGS_frame_number = frame;
    FLAGS_output_data_dir = outDir;
    FLAGS_prev_frame_data_dir = prevFrame && *prevFrame ? prevFrame : "NONE";
    FLAGS_output_equirect_path = outDir + "/eqr_" + frame + ".png";
    FLAGS_output_cubemap_path = P->want_cubemap ? outDir + "/cube_" + frame + ".png" : "";
    FLAGS_interpupilary_dist = P->interpupilary_dist;
    FLAGS_side_alpha_feather_size = P->side_alpha_feather_size;
    FLAGS_std_alpha_feather_size = P->std_alpha_feather_size;
    FLAGS_save_debug_images = false;
    FLAGS_sharpening = P->sharpening;
    FLAGS_enable_top = P->enable_top != 0;
    FLAGS_enable_bottom = P->enable_bottom != 0;
    FLAGS_enable_pole_removal = P->enable_pole_removal != 0;
    FLAGS_bottom_pole_masks_dir = masks;
    FLAGS_side_flow_alg = P->side_search20 ? "pixflow_search_20" : "pixflow_low";
    FLAGS_polar_flow_alg = P->polar_search20 ? "pixflow_search_20" : "pixflow_low";
    FLAGS_poleremoval_flow_alg = "pixflow_low";
    FLAGS_zero_parallax_dist = P->zero_parallax_dist;
    FLAGS_eqr_width = P->eqr_width;
    FLAGS_eqr_height = P->eqr_height;
    FLAGS_final_eqr_width = P->final_eqr_width;
    FLAGS_final_eqr_height = P->final_eqr_height;
    FLAGS_cubemap_width = P->cubemap_width;
    FLAGS_cubemap_height = P->cubemap_height;
    FLAGS_cubemap_format = P->cubemap_photo ? "photo" : "video";
    const double t0 = getCurrTimeSec();
    renderStereoPanorama();
    if (seconds) *seconds = getCurrTimeSec() - t0;
    Mat eq;
    if (!cv::refshim::memGet(FLAGS_output_equirect_path, eq)) throw VrCamException("ref_render_frame: no equirect was written");
    *eqW = eq.cols;
    *eqH = eq.rows;
    if (equirect) out8(eq, equirect);
    if (P->want_cubemap) {
      Mat cube;
      if (!cv::refshim::memGet(FLAGS_output_cubemap_path, cube)) throw VrCamException("ref_render_frame: no cubemap was written");
      *cubeW = cube.cols;
      *cubeH = cube.rows;
      if (cubemap) out8(cube, cubemap);
    }
  });
}

// an image the last ref_render_frame wrote below its output directory (e.g. "debug/000000/flow_images/overlap_3_L.png")
int ref_frame_image(const char* workDir, const char* relPath, uint8_t* out, int* w, int* h, int* c) {
  return guarded([&] {
    Mat m;
    if (!cv::refshim::memGet(std::string(workDir) + "/out/" + relPath, m)) throw VrCamException(std::string("no such image: ") + relPath);
    *w = m.cols; *h = m.rows; *c = m.channels();
    if (out) out8(m, out);
  });
}
// a flow file the last frame wrote (readFlowFromFile, CvUtil.cpp:184-199), e.g. "flow/000000/flowLtoR_3.bin"
int ref_frame_flow(const char* workDir, const char* relPath, float* out, int* w, int* h) {
  return guarded([&] {
    Mat f = readFlowFromFile(std::string(workDir) + "/out/" + relPath);
    *w = f.cols; *h = f.rows;
    if (out) outF(f, out);
  });
}

}  // extern "C"
