// Soft ISP on sm_100a: raw Bayer frame -> RGB, the step in front of the render path. Mirrors the reference's
// `CameraIsp` (camera_isp/CameraIsp.h:44-1301; the CPU branch of Raw2Rgb, camera_isp/Raw2Rgb.cpp:445-456): same configuration JSON
// (res/config/isp/*.json), same setters, same pipeline order (executePipeline, CameraIsp.h:1262-1272):
//   loadImage/resizeInput -> blackLevelAdjust -> antiVignette -> whiteBalance -> clampAndStretch -> removeStuckPixels -> demosaic
//   (bilinear or edge aware) -> colorCorrect (composite CCM + tone-curve LUT) -> sharpen (two-tap IIR low pass + cored unsharp mask)
//   -> 8/16-bit BGR or RGB.
// Every table the reference derives with the C library on the host (tone curve: powf/tanf; vignette Bezier curves; composite colour
// matrix through cv::gemm) is derived on the host here too, with the same float operations; the kernels repeat the per-pixel float
// arithmetic in the reference's order without contraction, so the output is bit-identical to the reference build (tests/).
// Not restated: the frequency-domain demosaic (cv::dct, CameraIsp.h:1176-1196) and a non-zero stuckPixelRadius (a sequential in-place
// scan, CameraIsp.h:1024-1103) -- both raise an error; every shipped configuration has stuckPixelRadius 0.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <string>
#include <vector>

namespace s360 {

constexpr int kToneCurveLutSize = 4096;   // CameraIsp.h:42

struct IspP3 { float x = 0, y = 0, z = 0; };

class IspPipeline {
 public:
  IspPipeline(const std::string& jsonInput, int outputBpp);   // CameraIsp::CameraIsp, CameraIsp.h:428-611
  ~IspPipeline();
  IspPipeline(const IspPipeline&) = delete;
  IspPipeline& operator=(const IspPipeline&) = delete;

  void setBitsPerPixel(int b) { bitsPerPixel_ = b; }            // :860 (maxPixelValue is NOT recomputed, as in the reference)
  void setDemosaicFilter(int f);                               // :968
  void setResize(int r);                                       // :978
  void enableToneMap() { disableToneCurve_ = false; buildToneCurveLut(); }    // :958
  void disableToneMap() { disableToneCurve_ = true; buildToneCurveLut(); }    // :963
  void addBlackLevelOffset(int o) { blackLevel_.x += o; blackLevel_.y += o; blackLevel_.z += o; }   // :884
  int outputBpp() const { return outputBpp_; }
  int resize() const { return resize_; }

  // the tables of setup() for an image of w x h pixels after the resize (host only; no device needed)
  void tables(int w, int h, float* ccm9, float* lut, float* vigH, float* vigV);

  // loadImage + getImage (CameraIsp.h:831-857, 1275-1299) on device buffers, asynchronous on `st`:
  // raw: inW x inH Bayer samples of depthBits (8 or 16) with strideBytes between rows; out: (inH / resize) x (inW / resize) x 3 of
  // outputBpp bits, tightly packed, BGR when swizzle else RGB.
  // highByte (16-bit ISPs only): store the high byte of every 16-bit sample as uint8 -- what cv::imread(CV_LOAD_IMAGE_COLOR) makes of
  // the 16-bit PNG the reference's Unpacker writes, i.e. the renderer's input without the PNG round trip.
  void process(cudaStream_t st, const void* rawDev, int inW, int inH, int depthBits, size_t strideBytes, void* outDev, bool swizzle, bool highByte = false);
  // n images of one size through the same launches (blockIdx.z / .y = image): the rig's 17 cameras per frame. The IIR passes have
  // only one thread per image line, so batching is what fills the GPU there.
  // footageBits 8 / 12 (with depthBits 16): rawDev holds the footage container's packed samples (row pitch strideBytes = inW *
  // footageBits / 8 when dense); they are widened on the fly as the Unpacker's RawConverter widens them before loadImage
  // (Unpacker.cpp:141-180, RawConverter.cpp:14-59).
  void processBatch(cudaStream_t st, int n, const void* const* rawDev, int inW, int inH, int depthBits, size_t strideBytes, void* const* outDev, bool swizzle,
                    bool highByte = false, int footageBits = 0);

 private:
  void setup();               // :612-715
  void buildToneCurveLut();   // :390-425
  void ensureDevice(int w, int h, int n);

  // configuration (CameraIsp's members)
  std::string bayerPattern_;
  IspP3 blackLevel_, clampMin_, clampMax_, whiteBalanceGain_, gamma_, lowKeyBoost_, highKeyBoost_, sharpening_;
  std::vector<IspP3> vignetteRollOffH_, vignetteRollOffV_;
  int stuckPixelThreshold_ = 0, stuckPixelRadius_ = 0, bitsPerPixel_ = 8, maxPixelValue_ = 255;
  float stuckPixelDarknessThreshold_ = 0, saturation_ = 1, contrast_ = 1, sharpeningSupport_ = 0, noiseCore_ = 1000;
  float ccm_[9], compositeCCM_[9];
  bool redBayerPixel_[2][2], greenBayerPixel_[2][2];
  int demosaicFilter_ = 2, resize_ = 1;
  bool disableToneCurve_ = false;
  std::vector<float> toneCurveLut_;   // 4096 x 3
  const int outputBpp_;

  // device side
  float* dLut_ = nullptr;        // 4096 x 3
  bool lutDirty_ = true;
  float* dVig_ = nullptr;        // vigH (w x 3) then vigV (h x 3)
  int vigW_ = 0, vigH_ = 0;
  void* dWork_ = nullptr;
  size_t workBytes_ = 0;
};

// RawConverter::convert8Frame / convert12Frame (camera_isp/RawConverter.cpp:14-59) for n frames of one size: packed footage samples
// (bits 8 or 12, dense) -> w x h uint16 planes. Device pointers in host arrays, asynchronous on `st`.
void rawConvertBatch(cudaStream_t st, int n, const void* const* srcDev, int bits, int w, int h, void* const* dstDev);

// host restatement of the vignette curve evaluation (BezierCurve, MathUtil.h:186-216; curveHAtPixel / curveVAtPixel, CameraIsp.h:707-713)
void ispBezier(const std::vector<IspP3>& pts, float t, float out[3]);

}  // namespace s360
