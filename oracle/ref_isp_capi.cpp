// ORACLE — TEST INFRASTRUCTURE ONLY. Not part of the product path.
//
// C entry points over the reference's OWN soft ISP (camera_isp/CameraIsp.h, compiled unchanged from /root/reference):
// the call sequence of Raw2Rgb's CPU branch (camera_isp/Raw2Rgb.cpp:445-456) -- CameraIsp(json, outputBpp), setBitsPerPixel(16),
// setDemosaicFilter, setResize, enableToneMap / disableToneMap, addBlackLevelOffset, loadImage, getImage. The Halide branch
// (CameraIspPipe, --accelerate, what Unpacker.cpp:174-181 uses) needs the Halide compiler to generate CameraIspGen* and cannot be
// built in this image; DESIGN.md says what differs.
#include <cstring>
#include <mutex>
#include <string>

#include "camera_isp/CameraIsp.h"
#ifdef REF_WITH_B200_ISP
#include "B200CameraIsp.h"
#endif

namespace {
std::string gErr;
std::mutex gMutex;
using namespace surround360;

// read access to the tables setup() derives (protected members) -- for pinning the host-side restatement of the product
struct IspProbe : CameraIsp {
  IspProbe(const std::string& json, int bpp) : CameraIsp(json, bpp) {}
  const cv::Mat& ccmComposite() const { return compositeCCM; }
  const std::vector<cv::Vec3f>& lut() const { return toneCurveLut; }
  cv::Vec3f vigH(int x) { return curveHAtPixel(x); }
  cv::Vec3f vigV(int y) { return curveVAtPixel(y); }
};

void configure(IspProbe& isp, int demosaicFilter, int resize, int toneCurve, int blackLevelOffset) {
  static const int kIspInputBitsPerPixel = 16;   // Raw2Rgb.cpp:47
  isp.setBitsPerPixel(kIspInputBitsPerPixel);
  isp.setDemosaicFilter(demosaicFilter);
  isp.setResize(resize);
  if (toneCurve) isp.enableToneMap(); else isp.disableToneMap();
  isp.addBlackLevelOffset(blackLevelOffset);
}
}  // namespace

extern "C" {
const char* refisp_last_error() { return gErr.c_str(); }

// raw: w x h Bayer samples, 8 bit (depthBits 8) or 16 bit (depthBits 16), tightly packed; out: (h / resize) x (w / resize) x 3,
// uint8 (outputBpp 8) or uint16 (outputBpp 16), BGR when swizzle != 0 (getImage's default) else RGB
int refisp_process(const char* json, const void* raw, int w, int h, int depthBits, int demosaicFilter, int resize, int outputBpp, int toneCurve,
                   int blackLevelOffset, int swizzle, void* out) {
  std::lock_guard<std::mutex> lock(gMutex);
  try {
    google::refShimMinLogLevel() = 1;
    IspProbe isp(json, outputBpp);
    configure(isp, demosaicFilter, resize, toneCurve, blackLevelOffset);
    cv::Mat input(h, w, depthBits == 8 ? CV_8U : CV_16U);
    std::memcpy(input.data, raw, (size_t)w * h * (depthBits == 8 ? 1 : 2));
    isp.loadImage(input);
    const int ow = w / resize, oh = h / resize;
    cv::Mat output(oh, ow, outputBpp == 8 ? CV_8UC3 : CV_16UC3);
    isp.getImage(output, swizzle != 0);
    std::memcpy(out, output.data, (size_t)ow * oh * 3 * (outputBpp == 8 ? 1 : 2));
    return 0;
  } catch (const std::exception& e) {
    gErr = e.what();
    return 1;
  }
}

#ifdef REF_WITH_B200_ISP
// The adapter of integration/reference_tree/camera_isp/B200CameraIsp.h inside the reference build, driven through the reference's
// base-class interface exactly like Raw2Rgb.cpp:445-456 drives CameraIsp (links the product library, like a maintainer's build).
int refisp_process_b200(const char* json, const void* raw, int w, int h, int depthBits, int demosaicFilter, int resize, int outputBpp, int toneCurve,
                        int blackLevelOffset, int swizzle, void* out) {
  std::lock_guard<std::mutex> lock(gMutex);
  try {
    google::refShimMinLogLevel() = 1;
    B200CameraIsp isp(json, outputBpp);
    isp.setBitsPerPixel(16);
    isp.setDemosaicFilter(demosaicFilter);
    isp.setResize(resize);
    if (toneCurve) isp.enableToneMap(); else isp.disableToneMap();
    isp.addBlackLevelOffset(blackLevelOffset);
    cv::Mat input(h, w, depthBits == 8 ? CV_8U : CV_16U);
    std::memcpy(input.data, raw, (size_t)w * h * (depthBits == 8 ? 1 : 2));
    CameraIsp* base = &isp;   // the pixel path goes through the virtual members
    base->loadImage(input);
    const int ow = w / resize, oh = h / resize;
    cv::Mat output(oh, ow, outputBpp == 8 ? CV_8UC3 : CV_16UC3);
    base->getImage(output, swizzle != 0);
    std::memcpy(out, output.data, (size_t)ow * oh * 3 * (outputBpp == 8 ? 1 : 2));
    return 0;
  } catch (const std::exception& e) {
    gErr = e.what();
    return 1;
  }
}
#endif

// the tables of setup(): compositeCCM (9 floats, row major), toneCurveLut (4096 x 3 floats), vignette gains of every column / row
// (w x 3, h x 3) for an image of w x h after the resize
int refisp_tables(const char* json, int outputBpp, int toneCurve, int w, int h, float* ccm9, float* lut, float* vigH, float* vigV) {
  std::lock_guard<std::mutex> lock(gMutex);
  try {
    google::refShimMinLogLevel() = 1;
    IspProbe isp(json, outputBpp);
    configure(isp, 0, 1, toneCurve, 0);
    cv::Mat input = cv::Mat::zeros(h, w, CV_16U);
    isp.loadImage(input);
    for (int i = 0; i < 9; ++i) ccm9[i] = isp.ccmComposite().at<float>(i / 3, i % 3);
    for (int i = 0; i < kToneCurveLutSize; ++i) for (int c = 0; c < 3; ++c) lut[i * 3 + c] = isp.lut()[i][c];
    for (int x = 0; x < w; ++x) { const cv::Vec3f v = isp.vigH(x); for (int c = 0; c < 3; ++c) vigH[x * 3 + c] = v[c]; }
    for (int y = 0; y < h; ++y) { const cv::Vec3f v = isp.vigV(y); for (int c = 0; c < 3; ++c) vigV[y * 3 + c] = v[c]; }
    return 0;
  } catch (const std::exception& e) {
    gErr = e.what();
    return 1;
  }
}
}  // extern "C"
