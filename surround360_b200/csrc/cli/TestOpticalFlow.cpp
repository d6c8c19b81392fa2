// TestOpticalFlow --mode test: the reference's harness for BASELINE config 1 (source/test/TestOpticalFlow.cpp:50-143), on the
// B200 path. Reads --test_dir/--left_img and --right_img (PNG; BGRA kept, an alpha channel of 255 is added when missing),
// runs NovelViewGeneratorAsymmetricFlow::prepare = flow L->R with hint LEFT and flow R->L with hint RIGHT
// (NovelView.cpp:270-299) --repetitions times and prints the reference's timing line per repetition:
//     "RUNTIME (sec) = <seconds>"                                     (TestOpticalFlow.cpp:77-81)
// The flow fields are written as <test_dir>/disparity/flowLtoR_<alg>.bin and flowRtoL_<alg>.bin in the reference's flow format
// (saveFlowToFile, util/CvUtil.cpp:159-182) when that directory exists. Not reproduced: the flow visualisation PNGs and the
// morphing novel views (OpticalFlowVisualization.cpp, NovelViewUtil::generateNovelViewSimpleCvRemap: debugging aids off the
// render path), and --mode middlebury_interpolation_experiment.
#include <sys/stat.h>

#include <chrono>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <map>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../../include/s360.h"
#include "png_io.h"

namespace {
struct VrCamException : std::runtime_error { using std::runtime_error::runtime_error; };
double nowSec() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
void check(int rc) { if (rc) throw VrCamException(s360_last_error()); }

int run(int argc, char** argv) {
  std::map<std::string, std::string> v = {{"mode", ""}, {"test_dir", ""}, {"left_img", ""}, {"right_img", ""}, {"num_intermediate_views", "11"},
                                          {"flow_alg", ""}, {"repetitions", "1"}, {"save_asymmetric_novel_views", "false"},
                                          {"show_interpolated_view", "false"}, {"device", "0"},
                                          {"logbuflevel", "0"}, {"log_dir", ""}, {"stderrthreshold", "2"}, {"v", "0"}, {"logtostderr", "false"}};
  for (int i = 1; i < argc; ++i) {
    std::string a = argv[i];
    if (a.size() < 2 || a[0] != '-') throw VrCamException("unexpected argument: " + a);
    a = a.substr(a[1] == '-' ? 2 : 1);
    std::string name = a, val;
    const size_t eq = a.find('=');
    const bool isBool = name == "save_asymmetric_novel_views" || name == "show_interpolated_view" || name == "logtostderr";
    if (eq != std::string::npos) { name = a.substr(0, eq); val = a.substr(eq + 1); }
    else if (isBool) val = "true";
    else { if (i + 1 >= argc) throw VrCamException("flag '" + name + "' is missing its argument"); val = argv[++i]; }
    if (!v.count(name)) throw VrCamException("unknown command line flag '" + name + "'");
    v[name] = val;
  }
  auto req = [&](const char* n) { if (v[n].empty()) throw VrCamException(std::string("missing required command line argument: ") + n); };
  req("mode");
  if (v["mode"] != "test") throw VrCamException("unrecongized mode: " + v["mode"] + " (this build implements --mode test)");
  req("test_dir"); req("left_img"); req("right_img"); req("flow_alg");   // TestOpticalFlow.cpp:51-54
  const std::string dir = v["test_dir"];
  auto load = [&](const std::string& p) {
    try { return pngio::readColor(p, true); }      // imreadExceptionOnFail(path, -1) + cvtColor(BGR2BGRA) when there is no alpha
    catch (const std::exception& e) { throw VrCamException("failed to load image: " + p); }
  };
  std::cerr << "reading opencv image matrices\n";
  const pngio::Image L = load(dir + "/" + v["left_img"]), R = load(dir + "/" + v["right_img"]);
  if (L.w != R.w || L.h != R.h) throw VrCamException("left and right image differ in size");
  s360_ctx* ctx = nullptr;
  check(s360_ctx_create(std::atoi(v["device"].c_str()), &ctx));
  check(s360_ctx_set_latency_mode(ctx, 1));   // one pair at a time: the latency variant of the sweeps (same bytes)
  std::vector<float> fLR((size_t)L.w * L.h * 2), fRL(fLR.size());
  const int reps = std::max(1, std::atoi(v["repetitions"].c_str()));
  for (int rep = 0; rep < reps; ++rep) {
    std::cerr << "---- repetition " << rep << "\n";
    std::cerr << "calling prepare\n";
    const double t0 = nowSec();
    check(s360_flow_compute(ctx, v["flow_alg"].c_str(), L.data.data(), R.data.data(), L.w, L.h, nullptr, nullptr, nullptr, fLR.data(), S360_HINT_LEFT));
    check(s360_flow_compute(ctx, v["flow_alg"].c_str(), R.data.data(), L.data.data(), L.w, L.h, nullptr, nullptr, nullptr, fRL.data(), S360_HINT_RIGHT));
    const double t1 = nowSec();
    std::cerr << "RUNTIME (sec) = " << (t1 - t0) << "\n";
  }
  struct stat st;
  const std::string out = dir + "/disparity";
  if (stat(out.c_str(), &st) == 0) {
    check(s360_flow_save((out + "/flowLtoR_" + v["flow_alg"] + ".bin").c_str(), fLR.data(), L.w, L.h));
    check(s360_flow_save((out + "/flowRtoL_" + v["flow_alg"] + ".bin").c_str(), fRL.data(), L.w, L.h));
  }
  s360_ctx_destroy(ctx);
  return EXIT_SUCCESS;
}
}  // namespace

int main(int argc, char** argv) {
  try {
    return run(argc, argv);
  } catch (const std::exception& e) {
    std::cerr << "Terminated with VrCamException: " << e.what() << std::endl;   // util/SystemUtil.cpp:42-61
    return EXIT_FAILURE;
  }
}
