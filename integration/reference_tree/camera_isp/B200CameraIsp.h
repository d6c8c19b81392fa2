// B200CameraIsp.h -- drop this file into surround360_render/source/camera_isp/ and link libs360b200.so. Wherever the reference
// constructs its soft ISP through the base-class interface, e.g. Raw2Rgb.cpp:445
//
//   CameraIsp cameraIsp(json, FLAGS_output_bpp);          ->   B200CameraIsp cameraIsp(json, FLAGS_output_bpp);
//
// the rest of the call sequence (setBitsPerPixel, setDemosaicFilter, setResize, enableToneMap / disableToneMap,
// addBlackLevelOffset, loadImage, runPipeline -> getImage, Raw2Rgb.cpp:446-456) stays as it is: the adapter forwards the setters to
// the GPU object of include/s360.h and overrides the two virtual members that touch pixels (loadImage, getImage; CameraIsp.h:831,
// 1275). The output is bit-identical to CameraIsp's (tests/test_gpu_isp.py runs this adapter inside the reference build). The
// frequency-domain demosaic filter is not available on the GPU path and raises VrCamException.
#pragma once
#include <string>

#include "CameraIsp.h"
#include "VrCamException.h"
#include "s360.h"

namespace surround360 {

class B200CameraIsp : public CameraIsp {
  s360_ctx* ctx_ = nullptr;
  s360_isp* isp_ = nullptr;
  cv::Mat input_;
  static void check(int rc) { if (rc) throw VrCamException(s360_last_error()); }

 public:
  B200CameraIsp(const std::string& jsonInput, const int outputBpp) : CameraIsp(jsonInput, outputBpp) {
    check(s360_ctx_create(0, &ctx_));   // no CUDA device: fails loudly, there is no CPU path behind this adapter
    check(s360_isp_create(jsonInput.c_str(), outputBpp, &isp_));
  }
  ~B200CameraIsp() { s360_isp_destroy(isp_); s360_ctx_destroy(ctx_); }
  B200CameraIsp(const B200CameraIsp&) = delete;
  B200CameraIsp& operator=(const B200CameraIsp&) = delete;

  // the base class's setters are not virtual: these shadow them, keep the base object's state in step and forward
  void setBitsPerPixel(const int bitsPerPixel) { CameraIsp::setBitsPerPixel(bitsPerPixel); check(s360_isp_set_bits_per_pixel(isp_, bitsPerPixel)); }
  void setDemosaicFilter(const int demosaicFilter) { CameraIsp::setDemosaicFilter(demosaicFilter); check(s360_isp_set_demosaic_filter(isp_, demosaicFilter)); }
  void setResize(const int resize) { CameraIsp::setResize(resize); check(s360_isp_set_resize(isp_, resize)); }
  void enableToneMap() { CameraIsp::enableToneMap(); check(s360_isp_set_tone_map(isp_, 1)); }
  void disableToneMap() { CameraIsp::disableToneMap(); check(s360_isp_set_tone_map(isp_, 0)); }
  void addBlackLevelOffset(const int offset) { CameraIsp::addBlackLevelOffset(offset); check(s360_isp_add_black_level_offset(isp_, offset)); }

  void loadImage(const cv::Mat& inputImage) override {
    const int depth = inputImage.type() & CV_MAT_DEPTH_MASK;
    if (depth != CV_8U && depth != CV_16U) throw VrCamException("input is larger that 16 bits per pixel");   // CameraIsp.h:855
    input_ = inputImage;   // cv::Mat shares the pixels: nothing is copied or converted on the host
  }
  void getImage(cv::Mat& outputImage, const bool swizzle = true) override {
    const int depthBits = (input_.type() & CV_MAT_DEPTH_MASK) == CV_8U ? 8 : 16;
    check(s360_isp_process(ctx_, isp_, input_.data, input_.cols, input_.rows, depthBits, (size_t)input_.step, outputImage.data, swizzle ? 1 : 0));
  }
};

}  // namespace surround360
