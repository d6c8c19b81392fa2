// ORACLE — TEST INFRASTRUCTURE ONLY. Not part of the product path.
//
// C entry point over the reference's OWN TestHyperPreview (test/TestHyperPreview.cpp:77-185, compiled unchanged from /root/reference
// with `main` renamed): PreviewRenderer = simpleDemosaic of the three raw Bayer frames (top, bottom, secondary bottom),
// alpha fades, projection to an equirect through a precomputed warp, flattenLayersAlphaSoftmax (util/CvUtil.cpp:336-360). The JPEG
// the reference writes is intercepted by the in-memory hook of the imwrite stand-in, so the float image `eqrImage * kNorm16to8` is
// returned as is.
#include <cstring>
#include <mutex>
#include <string>

#define main ref_hyper_preview_main
#include "test/TestHyperPreview.cpp"
#undef main

namespace {
std::string gErr;
std::mutex gMutex;
}

extern "C" {
const char* refp_last_error() { return gErr.c_str(); }

// raw*: uint16 Bayer frames imgW x imgH; out: eqrH x eqrW x 3 float32
int refp_render(const char* rigJson, const uint16_t* rawTop, const uint16_t* rawBottom, const uint16_t* rawBottom2, int imgW, int imgH, int eqrW,
                int eqrH, double gamma, double softmaxCoef, float* out) {
  std::lock_guard<std::mutex> lock(gMutex);
  try {
    google::refShimMinLogLevel() = 1;
    FLAGS_rig_json_file = rigJson;
    FLAGS_eqr_width = eqrW;
    FLAGS_eqr_height = eqrH;
    FLAGS_gamma = gamma;
    FLAGS_softmax_coef = softmaxCoef;
    FLAGS_preview_dest = "/refmem/preview";
    cv::refshim::memClear();
    cv::refshim::memSinkPrefix(FLAGS_preview_dest);
    PreviewRenderer r;
    auto wrap = [&](const uint16_t* p) { Mat m(imgH, imgW, CV_16U); std::memcpy(m.data, p, (size_t)imgW * imgH * 2); return m; };
    r.initRaws(Size(imgW, imgH), CV_16U);
    r.addTopImage(wrap(rawTop));
    r.addBottomImage(wrap(rawBottom));
    r.addBottomImage2(wrap(rawBottom2));
    r.render(7);
    Mat res;
    if (!cv::refshim::memGet(FLAGS_preview_dest + "/000007.jpg", res)) throw VrCamException("no preview image was written");
    if (res.type() != CV_32FC3 || res.cols != eqrW || res.rows != eqrH) throw VrCamException("unexpected preview image type");
    for (int y = 0; y < res.rows; ++y) std::memcpy(out + (size_t)y * eqrW * 3, res.ptr(y), (size_t)eqrW * 12);
    return 0;
  } catch (const std::exception& e) {
    gErr = e.what();
    return 1;
  }
}
}  // extern "C"
