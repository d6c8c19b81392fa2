// TestRenderStereoPanorama — command-line drop-in for the reference binary of the same name
// (surround360_render/source/test/TestRenderStereoPanorama.cpp): same flags and defaults
// (TRSP:44-70), same input layout <imgs_dir>/<camId>/<frame_number>.<ext>, same outputs
// (stereo equirect PNG, optional cubemap PNG, flow .bin files and the flow input PNGs the next frame's temporal
// regularisation reads; --prev_frame_data_dir <previous frame number> as in TRSP:210-235, 418-436). The render runs on the GPU through
// the C ABI (include/s360.h); this file only does flag parsing and file I/O.
#include <dirent.h>
#include <sys/stat.h>
#include <chrono>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <map>
#include <string>
#include <vector>
#include "../../../include/s360.h"
#include "png_io.h"

namespace {

struct VrCamException : std::runtime_error { using std::runtime_error::runtime_error; };  // util/VrCamException.h:18

double nowSec() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

// gflags-style parsing: --name=value, --name value, -name..., booleans as --name / --noname / --name=false
struct Flags {
  std::map<std::string, std::string> v = {
      {"rig_json_file", ""}, {"imgs_dir", ""}, {"frame_number", ""}, {"output_data_dir", ""}, {"prev_frame_data_dir", "NONE"},
      {"output_cubemap_path", ""}, {"output_equirect_path", ""}, {"interpupilary_dist", "6.4"}, {"side_alpha_feather_size", "100"},
      {"std_alpha_feather_size", "31"}, {"save_debug_images", "false"}, {"sharpening", "0.0"}, {"enable_top", "false"},
      {"enable_bottom", "false"}, {"enable_pole_removal", "false"}, {"bottom_pole_masks_dir", ""}, {"side_flow_alg", "pixflow_low"},
      {"polar_flow_alg", "pixflow_low"}, {"poleremoval_flow_alg", "pixflow_low"}, {"zero_parallax_dist", "10000.0"},
      {"eqr_width", "256"}, {"eqr_height", "128"}, {"final_eqr_width", "3480"}, {"final_eqr_height", "960"},
      {"cubemap_width", "1536"}, {"cubemap_height", "1536"}, {"cubemap_format", "video"},
      // additions of this build
      {"save_flow_data", "true"}, {"device", "0"}, {"latency_mode", "true"},
      // glog flags the driver script passes (scripts/batch_process_video.py:31-34): accepted
      {"logbuflevel", "0"}, {"log_dir", ""}, {"stderrthreshold", "2"}, {"v", "0"}, {"logtostderr", "false"}, {"alsologtostderr", "false"}};
  const std::vector<std::string> bools = {"save_debug_images", "enable_top", "enable_bottom", "enable_pole_removal", "save_flow_data", "latency_mode", "logtostderr", "alsologtostderr"};
  bool isBool(const std::string& n) const { for (auto& b : bools) if (b == n) return true; return false; }
  void parse(int argc, char** argv) {
    for (int i = 1; i < argc; ++i) {
      std::string a = argv[i];
      if (a.size() < 2 || a[0] != '-') throw VrCamException("unexpected argument: " + a);
      a = a.substr(a[1] == '-' ? 2 : 1);
      std::string name = a, val;
      bool hasVal = false;
      const size_t eq = a.find('=');
      if (eq != std::string::npos) { name = a.substr(0, eq); val = a.substr(eq + 1); hasVal = true; }
      if (!v.count(name) && name.rfind("no", 0) == 0 && isBool(name.substr(2))) { v[name.substr(2)] = "false"; continue; }
      if (!v.count(name)) throw VrCamException("unknown command line flag '" + name + "'");
      if (isBool(name)) { v[name] = hasVal ? val : "true"; continue; }
      if (!hasVal) { if (i + 1 >= argc) throw VrCamException("flag '" + name + "' is missing its argument"); val = argv[++i]; }
      v[name] = val;
    }
  }
  std::string s(const std::string& n) const { return v.at(n); }
  double d(const std::string& n) const { return std::atof(v.at(n).c_str()); }
  int i(const std::string& n) const { return std::atoi(v.at(n).c_str()); }
  bool b(const std::string& n) const { const std::string& x = v.at(n); return x == "true" || x == "1" || x == "t" || x == "yes" || x == "y"; }
};

void requireArg(const Flags& f, const std::string& name) {  // util/SystemUtil.h:45-49
  if (f.s(name).empty()) throw VrCamException("missing required command line argument: " + name);
}

// util/SystemUtil.h:96-105: extension of the first file in the camera's directory
std::string getImageFileExtension(const std::string& dir) {
  DIR* d = opendir(dir.c_str());
  if (!d) throw VrCamException("can't read directory: " + dir);
  std::string ext;
  while (dirent* e = readdir(d)) {
    const std::string n = e->d_name;
    if (n == "." || n == "..") continue;
    const size_t dot = n.rfind('.');
    if (dot != std::string::npos) { ext = n.substr(dot + 1); break; }
  }
  closedir(d);
  if (ext.empty()) throw VrCamException("no image files in: " + dir);
  return ext;
}

void check(int rc) { if (rc) throw VrCamException(s360_last_error()); }

void saveIntermediateFlow(s360_ctx* ctx, const char* name, int idx, const std::string& path) {
  int w, h, c, f;
  check(s360_frame_intermediate(ctx, name, idx, nullptr, &w, &h, &c, &f));
  std::vector<float> buf((size_t)w * h * 2);
  check(s360_frame_intermediate(ctx, name, idx, buf.data(), &w, &h, &c, &f));
  check(s360_flow_save(path.c_str(), buf.data(), w, h));
}

void saveIntermediateImage(s360_ctx* ctx, const char* name, int idx, const std::string& path) {
  int w, h, c, f;
  check(s360_frame_intermediate(ctx, name, idx, nullptr, &w, &h, &c, &f));
  std::vector<uint8_t> buf((size_t)w * h * 4);
  check(s360_frame_intermediate(ctx, name, idx, buf.data(), &w, &h, &c, &f));
  pngio::write(path, buf.data(), w, h, 4);
}

struct LoadedFlow {
  float* p = nullptr;
  ~LoadedFlow() { s360_free(p); }
};
// readFlowFromFile + size check against what this frame's flow of the same slot will be
void loadPrevFlow(const std::string& path, int w, int h, LoadedFlow& out) {
  int fw = 0, fh = 0;
  check(s360_flow_load(path.c_str(), &out.p, &fw, &fh));
  if (fw != w || fh != h) throw VrCamException("previous frame's flow has a different size: " + path);
}
pngio::Image loadPrevImage(const std::string& path, int w, int h) {  // imreadExceptionOnFail(path, -1)
  pngio::Image im;
  try { im = pngio::readColor(path, true); }
  catch (const std::exception& e) { throw VrCamException(std::string("failed to load image: ") + path + " (" + e.what() + ")"); }
  if (im.w != w || im.h != h) throw VrCamException("previous frame's flow image has a different size: " + path);
  return im;
}
const char* kPoleSlotNames[4] = {"top_left", "top_right", "bottom_left", "bottom_right"};  // eyeName, TRSP:389-391

int run(int argc, char** argv) {
  Flags F;
  F.parse(argc, argv);
  for (const char* n : {"rig_json_file", "imgs_dir", "frame_number", "output_data_dir", "output_equirect_path"}) requireArg(F, n);  // TRSP:717-721
  const bool verbose = F.i("v") >= 1;
  const double startTime = nowSec();
  s360_ctx* ctx = nullptr;
  check(s360_ctx_create(F.i("device"), &ctx));
  check(s360_ctx_set_latency_mode(ctx, F.b("latency_mode") ? 1 : 0));   // one frame per process: the latency variant of the sweeps (same bytes)
  check(s360_rig_load_file(ctx, F.s("rig_json_file").c_str()));
  s360_params P;
  s360_params_default(&P);
  P.interpupilary_dist = F.d("interpupilary_dist"); P.sharpening = F.d("sharpening"); P.zero_parallax_dist = F.d("zero_parallax_dist");
  P.side_alpha_feather_size = F.i("side_alpha_feather_size"); P.std_alpha_feather_size = F.i("std_alpha_feather_size");
  P.enable_top = F.b("enable_top"); P.enable_bottom = F.b("enable_bottom");
  P.enable_pole_removal = F.b("enable_pole_removal");
  std::strncpy(P.poleremoval_flow_alg, F.s("poleremoval_flow_alg").c_str(), sizeof(P.poleremoval_flow_alg) - 1);
  P.eqr_width = F.i("eqr_width"); P.eqr_height = F.i("eqr_height");
  P.final_eqr_width = F.i("final_eqr_width"); P.final_eqr_height = F.i("final_eqr_height");
  P.cubemap_width = F.i("cubemap_width"); P.cubemap_height = F.i("cubemap_height");
  P.want_cubemap = !F.s("output_cubemap_path").empty();
  const std::string fmt = F.s("cubemap_format");
  if (fmt != "video" && fmt != "photo") throw VrCamException("unexpected cubemap format: " + fmt + ". valid formats are: video,photo");
  P.cubemap_photo = fmt == "photo";
  std::strncpy(P.side_flow_alg, F.s("side_flow_alg").c_str(), sizeof(P.side_flow_alg) - 1);
  std::strncpy(P.polar_flow_alg, F.s("polar_flow_alg").c_str(), sizeof(P.polar_flow_alg) - 1);

  // ---- load images (RigDescription.cpp:80-108; TRSP:600-603, 651-654) ----
  const double t0 = nowSec();
  const int nCam = s360_rig_num_cameras(ctx), nSide = s360_rig_num_side(ctx);
  std::vector<pngio::Image> imgs(nCam);
  std::vector<const uint8_t*> ptrs(nCam, nullptr);
  const std::string dir = F.s("imgs_dir"), frame = F.s("frame_number");
  std::string ext;
  int iw = 0, ih = 0;
  auto load = [&](int cam, const std::string& e) {
    const std::string path = dir + "/" + s360_rig_camera_id(ctx, cam) + "/" + frame + "." + e;
    imgs[cam] = pngio::readColor(path);
    if (iw == 0) { iw = imgs[cam].w; ih = imgs[cam].h; }
    if (imgs[cam].w != iw || imgs[cam].h != ih) throw VrCamException("all camera images must have the same size: " + path);
    ptrs[cam] = imgs[cam].data.data();
  };
  for (int i = 0; i < nSide; ++i) {
    const int cam = s360_rig_side_index(ctx, i);
    if (i == 0) ext = getImageFileExtension(dir + "/" + s360_rig_camera_id(ctx, cam));
    load(cam, ext);
  }
  if (P.enable_top) load(s360_rig_special_camera(ctx, 0), "png");
  if (P.enable_bottom) load(s360_rig_special_camera(ctx, 1), "png");
  if (P.enable_bottom && P.enable_pole_removal) {  // TRSP:569-597, PoleRemoval.cpp:52-67
    requireArg(F, "bottom_pole_masks_dir");
    const int cam2 = s360_rig_special_camera(ctx, 2);
    load(cam2, "png");
    const std::string md = F.s("bottom_pole_masks_dir");
    const std::string p1 = md + "/" + s360_rig_camera_id(ctx, s360_rig_special_camera(ctx, 1)) + ".png", p2 = md + "/" + s360_rig_camera_id(ctx, cam2) + ".png";
    pngio::Image m1, m2;
    try { m1 = pngio::readColor(p1); m2 = pngio::readColor(p2); }
    catch (const std::exception&) { throw VrCamException("missing or bad pole mask:" + p1 + "," + p2); }
    if (m1.w != iw || m1.h != ih || m2.w != iw || m2.h != ih) throw VrCamException("missing or bad pole mask:" + p1 + "," + p2);
    check(s360_set_pole_masks(ctx, m1.data.data(), m2.data.data(), iw, ih));
  }
  if (verbose) std::cerr << "Time to load images from file: " << nowSec() - t0 << " sec\n";

  // ---- render ----
  int ew, eh, cw, ch;
  check(s360_frame_output_size(ctx, &P, &ew, &eh, &cw, &ch));
  std::vector<uint8_t> eq((size_t)ew * eh * 3), cube((size_t)std::max(cw, 1) * std::max(ch, 1) * 3);
  const std::string outDir = F.s("output_data_dir"), prev = F.s("prev_frame_data_dir");
  if (prev != "NONE") {  // TRSP:210-235, 418-436: flow fields and flow input images of the previous frame
    if (verbose) std::cerr << "Reading previous frame flow and images from: " << prev << "\n";
    check(s360_ctx_set_temporal(ctx, 1));
    check(s360_frame_prepare(ctx, &P, iw, ih));
    int ow, oh, xw, xh;
    check(s360_temporal_sizes(ctx, &ow, &oh, &xw, &xh));
    const std::string flowPrevDir = outDir + "/flow/" + prev, imgPrevDir = outDir + "/debug/" + prev + "/flow_images";
    for (int i = 0; i < nSide; ++i) {
      LoadedFlow lr, rl;
      loadPrevFlow(flowPrevDir + "/flowLtoR_" + std::to_string(i) + ".bin", ow, oh, lr);
      loadPrevFlow(flowPrevDir + "/flowRtoL_" + std::to_string(i) + ".bin", ow, oh, rl);
      const pngio::Image oL = loadPrevImage(imgPrevDir + "/overlap_" + std::to_string(i) + "_L.png", ow, oh);
      const pngio::Image oR = loadPrevImage(imgPrevDir + "/overlap_" + std::to_string(i) + "_R.png", ow, oh);
      check(s360_temporal_import_side(ctx, i, lr.p, rl.p, oL.data.data(), oR.data.data()));
    }
    for (int k = 0; k < 4; ++k) {
      if (!((k < 2 && P.enable_top) || (k >= 2 && P.enable_bottom))) continue;
      LoadedFlow fl;
      loadPrevFlow(flowPrevDir + "/flow_" + kPoleSlotNames[k] + ".bin", xw, xh, fl);
      const pngio::Image xf = loadPrevImage(imgPrevDir + "/extendedFisheyeSpherical_" + kPoleSlotNames[k] + ".png", xw, xh);
      check(s360_temporal_import_pole(ctx, k, fl.p, xf.data.data()));
    }
    check(s360_temporal_commit(ctx));
    if (verbose) std::cerr << "Loaded previous frame's flow OK\n";
  }
  check(s360_ctx_set_graphs(ctx, 0));  // single frame: plain launches, so the per-stage timers are filled
  check(s360_render_frame(ctx, &P, ptrs.data(), iw, ih, eq.data(), cw ? cube.data() : nullptr));
  float ms[7];
  check(s360_frame_stage_ms(ctx, ms));

  // ---- outputs ----
  pngio::write(F.s("output_equirect_path"), eq.data(), ew, eh, 3);
  if (cw) pngio::write(F.s("output_cubemap_path"), cube.data(), cw, ch, 3);
  if (F.b("save_flow_data")) {  // TRSP:246-255, 450-452: what the next frame's temporal regularisation reads
    const std::string flowDir = outDir + "/flow/" + frame, flowImagesDir = outDir + "/debug/" + frame + "/flow_images";
    struct stat st;
    if (stat(flowImagesDir.c_str(), &st) == 0) {  // TRSP:203-208, 413-415: "we will need them in the next frame"
      for (int i = 0; i < nSide; ++i) {
        saveIntermediateImage(ctx, "overlapL", i, flowImagesDir + "/overlap_" + std::to_string(i) + "_L.png");
        saveIntermediateImage(ctx, "overlapR", i, flowImagesDir + "/overlap_" + std::to_string(i) + "_R.png");
      }
      for (int k = 0; k < 4; ++k)
        if ((k < 2 && P.enable_top) || (k >= 2 && P.enable_bottom)) {
          saveIntermediateImage(ctx, "extSide", k, flowImagesDir + "/extendedSideSpherical_" + kPoleSlotNames[k] + ".png");
          saveIntermediateImage(ctx, "extFish", k, flowImagesDir + "/extendedFisheyeSpherical_" + kPoleSlotNames[k] + ".png");
        }
    } else {   // the reference fails here: imwriteExceptionOnFail (TRSP:203-208, util/CvUtil.cpp:44-52)
      throw VrCamException("failed to write image: " + flowImagesDir + "/overlap_0_L.png (the directory does not exist; the driver script creates it, "
                           "scripts/batch_process_video.py:116-118, or pass --save_flow_data=false)");
    }
    if (stat(flowDir.c_str(), &st) == 0) {  // directories are created by the driver script (batch_process_video.py:116-118)
      for (int i = 0; i < nSide; ++i) {
        saveIntermediateFlow(ctx, "flowLtoR", i, flowDir + "/flowLtoR_" + std::to_string(i) + ".bin");
        saveIntermediateFlow(ctx, "flowRtoL", i, flowDir + "/flowRtoL_" + std::to_string(i) + ".bin");
      }
      for (int k = 0; k < 4; ++k)
        if ((k < 2 && P.enable_top) || (k >= 2 && P.enable_bottom))
          saveIntermediateFlow(ctx, "poleFlow", k, flowDir + "/flow_" + kPoleSlotNames[k] + ".bin");
    } else {   // the reference fails here: saveFlowToFile (util/CvUtil.cpp:159-166)
      throw VrCamException("file not found: " + flowDir + "/flowLtoR_0.bin (the directory does not exist; pass --save_flow_data=false to skip the flow files)");
    }
  }
  if (F.b("save_debug_images")) {
    // The debug images the frame's kept intermediates provide (names as the reference writes them); the rest of the reference's
    // debug dumps (croppedSideSpherical_*, eqr_side*.png before sharpening, bottomImage*.png of the pole removal) are not produced.
    const std::string debugDir = outDir + "/debug/" + frame, projDir = debugDir + "/projections";
    struct stat st;
    if (stat(debugDir.c_str(), &st) != 0) throw VrCamException("failed to write image: " + debugDir + "/sphericalImgL.png");
    if (stat(projDir.c_str(), &st) == 0) {   // TRSP:177-184
      for (int i = 0; i < nSide; ++i) {
        const char* id = s360_rig_camera_id(ctx, s360_rig_side_index(ctx, i));
        saveIntermediateImage(ctx, "projection", i, projDir + "/crop_" + std::string(id ? id : "") + ".png");
      }
    } else {
      throw VrCamException("failed to write image: " + projDir + "/crop_*.png");
    }
    saveIntermediateImage(ctx, "panoL", 0, debugDir + "/sphericalImgL.png");   // TRSP:792-793
    saveIntermediateImage(ctx, "panoR", 0, debugDir + "/sphericalImgR.png");
    if (P.enable_top) saveIntermediateImage(ctx, "topSpherical", 0, debugDir + "/_topSpherical.png");          // TRSP:680
    if (P.enable_bottom) saveIntermediateImage(ctx, "bottomSpherical", 0, debugDir + "/_bottomSpherical.png");  // TRSP:639
    for (int k = 0; k < 4; ++k)
      if ((k < 2 && P.enable_top) || (k >= 2 && P.enable_bottom)) {
        saveIntermediateImage(ctx, "poleWarped", k, debugDir + "/warpedSpherical_" + kPoleSlotNames[k] + ".png");        // TRSP:551-553
        saveIntermediateImage(ctx, "extSide", k, debugDir + "/extendedSideSpherical_" + kPoleSlotNames[k] + ".png");     // TRSP:554-556
      }
    std::cerr << "note: --save_debug_images wrote projections/crop_*, sphericalImg{L,R}, _{top,bottom}Spherical, warpedSpherical_*, "
                 "extendedSideSpherical_*; the reference's other debug dumps are not produced by this build\n";
  }
  if (verbose) {  // TRSP:964-971
    std::cerr << "--- Runtime breakdown (sec) ---\n";
    std::cerr << "Total:\t\t\t" << nowSec() - startTime << "\n";
    std::cerr << "Spherical projection:\t" << ms[1] / 1e3 << "\n";
    std::cerr << "Side optical flow:\t\t" << ms[2] / 1e3 << "\n";
    std::cerr << "Novel view panorama:\t" << ms[3] / 1e3 << "\n";
    std::cerr << "Flow top+bottom with sides:\t" << ms[4] / 1e3 << "\n";
    std::cerr << "Sharpen:\t\t" << ms[5] / 1e3 << "\n";
    std::cerr << "Equirect -> Cubemap:\t" << ms[6] / 1e3 << "\n";
  }
  s360_ctx_destroy(ctx);
  return EXIT_SUCCESS;
}

}  // namespace

int main(int argc, char** argv) {
  try {
    return run(argc, argv);
  } catch (const std::exception& e) {
    // the reference's terminate handler logs what() and aborts (util/SystemUtil.cpp:42-61)
    std::cerr << "Terminated with VrCamException: " << e.what() << std::endl;
    return EXIT_FAILURE;
  }
}
