// B200Flow.h -- drop this file into surround360_render/source/optical_flow/ and add the two lines below to
// makeOpticalFlowByName (OpticalFlowFactory.h:23-64); link libs360b200.so. Nothing else in the reference changes:
//
//   if (flowAlgName == "b200_pixflow_low")       return new B200Flow("pixflow_low");
//   if (flowAlgName == "b200_pixflow_search_20") return new B200Flow("pixflow_search_20");
//
// The adapter implements OpticalFlowInterface (OpticalFlowInterface.h:19-42) on top of the C ABI of include/s360.h. cv::Mat
// views (the renderer crops its overlap images as ROIs, TestRenderStereoPanorama.cpp:196-198) are passed with their row
// stride, nothing is cloned. Results are bit-identical to PixFlow's (tests/test_gpu_reference_adapter.py runs the reference's
// own renderer, compiled from /root/reference, with --side_flow_alg b200_pixflow_low and compares the panoramas).
#pragma once
#include <string>

#include "OpticalFlowInterface.h"
#include "VrCamException.h"
#include "s360.h"

namespace surround360 {
namespace optical_flow {

struct B200Flow : public OpticalFlowInterface {
  s360_ctx* ctx = nullptr;
  std::string alg;   // "pixflow_low" or "pixflow_search_20"
  explicit B200Flow(const std::string& a) : alg(a) {
    if (s360_ctx_create(0, &ctx)) throw VrCamException(s360_last_error());   // no CUDA device: fails loudly, there is no CPU path
  }
  ~B200Flow() { s360_ctx_destroy(ctx); }

  void computeOpticalFlow(
      const cv::Mat& I0BGRA,
      const cv::Mat& I1BGRA,
      const cv::Mat& prevFlow,
      const cv::Mat& prevI0BGRA,
      const cv::Mat& prevI1BGRA,
      cv::Mat& flow,
      DirectionHint hint) override {
    (void)prevI0BGRA;   // PixFlow resizes it and never reads it (PixFlow.h:106)
    flow.create(I0BGRA.size(), CV_32FC2);
    const bool temporal = prevFlow.dims > 0;
    const int rc = s360_flow_compute_strided(
      ctx, alg.c_str(),
      I0BGRA.data, (size_t)I0BGRA.step, I1BGRA.data, (size_t)I1BGRA.step, I0BGRA.cols, I0BGRA.rows,
      temporal ? (const float*)prevFlow.data : nullptr, temporal ? (size_t)prevFlow.step : 0,
      temporal ? prevI1BGRA.data : nullptr, temporal ? (size_t)prevI1BGRA.step : 0,
      (float*)flow.data, (size_t)flow.step, (s360_hint)int(hint));   // DirectionHint and s360_hint share their numeric order
    if (rc) throw VrCamException(s360_last_error());                 // the exception type the factory itself throws
  }
};

} // namespace optical_flow
} // namespace surround360
