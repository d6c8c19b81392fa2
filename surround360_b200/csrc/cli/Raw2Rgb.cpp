// Raw2Rgb: the reference's raw -> RGB tool (source/camera_isp/Raw2Rgb.cpp:33-45, 377-462) on the B200 path. Same flags:
//   --input_image_path (a grey 8/16-bit PNG, or a headerless 16-bit ".raw" whose size comes from the configuration's "width" /
//   "height", Raw2Rgb.cpp:393-404) --output_image_path (PNG) --isp_config_path --black_level_offset --demosaic_filter (0 bilinear,
//   2 edge aware = default; 1, the frequency-domain filter, is rejected) --resize (1, 2, 4, 8) --output_bpp (8 | 16)
//   --disable_tone_curve.   --accelerate / --fast select the reference's Halide pipeline and are accepted but ignored (one GPU path).
// 8-bit input is widened with the reference's convert8bitTo16bit (value << 8), all input is fed as 16-bit (kIspInputBitsPerPixel).
// Prints the reference's "Runtime = <ms>ms" line (Raw2Rgb.cpp:369-372). Not reproduced: --output_dng_path (a DNG container writer,
// host I/O) and TIFF input (no TIFF codec here: convert to PNG).
#include <chrono>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <iterator>
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

// the two integers Raw2Rgb reads from the configuration for ".raw" input (getInteger(config, "CameraIsp", "width"), ...)
int configInt(const std::string& json, const std::string& key) {
  const size_t k = json.find("\"" + key + "\"");
  if (k == std::string::npos) throw VrCamException("JSON error. missing key:" + key);
  const size_t c = json.find(':', k);
  if (c == std::string::npos) throw VrCamException("JSON error. missing key:" + key);
  return std::atoi(json.c_str() + c + 1);
}

int run(int argc, char** argv) {
  std::map<std::string, std::string> v = {{"input_image_path", ""}, {"output_image_path", ""}, {"output_dng_path", ""}, {"isp_config_path", ""},
                                          {"black_level_offset", "0"}, {"demosaic_filter", "2"}, {"resize", "1"}, {"output_bpp", "8"},
                                          {"accelerate", "false"}, {"fast", "false"}, {"disable_tone_curve", "false"}, {"device", "0"},
                                          {"logbuflevel", "0"}, {"log_dir", ""}, {"stderrthreshold", "2"}, {"v", "0"}, {"logtostderr", "false"}};
  for (int i = 1; i < argc; ++i) {
    std::string a = argv[i];
    if (a.size() < 2 || a[0] != '-') throw VrCamException("unexpected argument: " + a);
    a = a.substr(a[1] == '-' ? 2 : 1);
    std::string name = a, val;
    const size_t eq = a.find('=');
    const bool isBool = name == "accelerate" || name == "fast" || name == "disable_tone_curve" || name == "logtostderr";
    if (eq != std::string::npos) { name = a.substr(0, eq); val = a.substr(eq + 1); }
    else if (isBool) val = "true";
    else { if (i + 1 >= argc) throw VrCamException("flag '" + name + "' is missing its argument"); val = argv[++i]; }
    if (!v.count(name)) throw VrCamException("unknown command line flag '" + name + "'");
    v[name] = val;
  }
  auto req = [&](const char* n) { if (v[n].empty()) throw VrCamException(std::string("missing required command line argument: ") + n); };
  req("input_image_path"); req("output_image_path"); req("isp_config_path");   // Raw2Rgb.cpp:379-381
  if (!v["output_dng_path"].empty()) throw VrCamException("--output_dng_path is not implemented by this tool");
  std::ifstream ifs(v["isp_config_path"], std::ios::in);
  if (!ifs) throw VrCamException("file read failed: " + v["isp_config_path"]);
  const std::string json((std::istreambuf_iterator<char>(ifs)), std::istreambuf_iterator<char>());
  const int outputBpp = std::atoi(v["output_bpp"].c_str()), resize = std::atoi(v["resize"].c_str());

  int w = 0, h = 0;
  std::vector<uint16_t> raw16;
  const std::string& in = v["input_image_path"];
  if (in.find(".raw") != std::string::npos) {   // readRaw, Raw2Rgb.cpp:332-361
    w = configInt(json, "width"); h = configInt(json, "height");
    if (w <= 0 || h <= 0) throw VrCamException("bad raw image size in the ISP configuration");
    raw16.assign((size_t)w * h, 0);
    std::ifstream f(in, std::ios::in | std::ios::binary);
    if (!f) throw VrCamException("file read failed: " + in);
    f.read(reinterpret_cast<char*>(raw16.data()), (std::streamsize)raw16.size() * 2);
    if ((size_t)f.gcount() != raw16.size() * 2) std::cerr << "Warning: expected " << raw16.size() * 2 << " but only read " << f.gcount() << "\n";
  } else {
    pngio::Grey g;
    try { g = pngio::readGrey(in); } catch (const std::exception& e) { throw VrCamException(std::string("failed to load image: ") + e.what()); }
    w = g.w; h = g.h;
    raw16.resize((size_t)w * h);
    if (g.depthBits == 8) { std::cerr << "8 bit raw\n"; for (size_t i = 0; i < raw16.size(); ++i) raw16[i] = (uint16_t)(g.data[i] << 8); }   // convert8bitTo16bit
    else { std::cerr << "16 bit raw\n"; std::memcpy(raw16.data(), g.data.data(), raw16.size() * 2); }
  }
  if (!(w > 2 && h > 2)) return 0;   // Raw2Rgb.cpp:406: nothing to do
  s360_ctx* ctx = nullptr;
  check(s360_ctx_create(std::atoi(v["device"].c_str()), &ctx));
  s360_isp* isp = nullptr;
  check(s360_isp_create(json.c_str(), outputBpp, &isp));
  check(s360_isp_set_bits_per_pixel(isp, 16));                                       // kIspInputBitsPerPixel
  check(s360_isp_set_demosaic_filter(isp, std::atoi(v["demosaic_filter"].c_str())));
  check(s360_isp_set_resize(isp, resize));
  check(s360_isp_set_tone_map(isp, v["disable_tone_curve"] == "true" ? 0 : 1));
  check(s360_isp_add_black_level_offset(isp, std::atoi(v["black_level_offset"].c_str())));
  const int ow = w / resize, oh = h / resize;
  std::vector<uint8_t> out((size_t)ow * oh * 3 * (outputBpp / 8));
  const double t0 = nowSec();
  check(s360_isp_process(ctx, isp, raw16.data(), w, h, 16, 0, out.data(), 1));
  std::cerr << "Runtime = " << (nowSec() - t0) * 1000.0 << "ms\n";
  if (outputBpp == 8) pngio::write(v["output_image_path"], out.data(), ow, oh, 3);
  else pngio::write16(v["output_image_path"], reinterpret_cast<const uint16_t*>(out.data()), ow, oh);
  s360_isp_destroy(isp);
  s360_ctx_destroy(ctx);
  return 0;
}
}  // namespace

int main(int argc, char** argv) {
  try {
    return run(argc, argv);
  } catch (const std::exception& e) {
    std::cerr << "Raw2Rgb: " << e.what() << "\n";
    return 1;
  }
}
