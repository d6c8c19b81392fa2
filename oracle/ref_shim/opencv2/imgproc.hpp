// ORACLE — TEST INFRASTRUCTURE ONLY (see core.hpp). OpenCV imgproc entry points used by the reference's render
// path; each forwards to the cv2-pinned primitive in oracle/ocv.h (implementation: ../ref_shim.cpp).
#pragma once
#include "core.hpp"

namespace cv {
enum InterpolationFlags { INTER_NEAREST = 0, INTER_LINEAR = 1, INTER_CUBIC = 2, INTER_AREA = 3, INTER_LANCZOS4 = 4 };
enum MorphShapes { MORPH_RECT = 0, MORPH_CROSS = 1, MORPH_ELLIPSE = 2 };
enum ColorConversionCodes { COLOR_BGR2BGRA = 0, COLOR_BGRA2BGR = 1, COLOR_BGR2GRAY = 6, COLOR_GRAY2BGR = 8, COLOR_BGRA2GRAY = 10 };

void resize(const Mat& src, OutputArray dst, Size dsize, double fx = 0, double fy = 0, int interpolation = INTER_LINEAR);
void remap(const Mat& src, OutputArray dst, const Mat& map1, const Mat& map2, int interpolation, int borderMode = BORDER_CONSTANT,
           const Scalar& borderValue = Scalar());
void GaussianBlur(const Mat& src, OutputArray dst, Size ksize, double sigmaX, double sigmaY = 0, int borderType = BORDER_DEFAULT);
void Sobel(const Mat& src, OutputArray dst, int ddepth, int dx, int dy, int ksize = 3, double scale = 1, double delta = 0,
           int borderType = BORDER_DEFAULT);
void medianBlur(const Mat& src, OutputArray dst, int ksize);
Mat getStructuringElement(int shape, Size ksize, Point anchor = Point(-1, -1));
void erode(const Mat& src, OutputArray dst, const Mat& kernel);
void cvtColor(const Mat& src, OutputArray dst, int code, int dstCn = 0);
}  // namespace cv

// OpenCV 2.x C constants the reference still spells (imgproc/types_c.h, imgcodecs/imgcodecs_c.h)
#define CV_INTER_NN cv::INTER_NEAREST
#define CV_INTER_LINEAR cv::INTER_LINEAR
#define CV_INTER_CUBIC cv::INTER_CUBIC
#define CV_BGR2BGRA cv::COLOR_BGR2BGRA
#define CV_BGRA2BGR cv::COLOR_BGRA2BGR
#define CV_BGR2GRAY cv::COLOR_BGR2GRAY
#define CV_BGRA2GRAY cv::COLOR_BGRA2GRAY
#define CV_GRAY2BGR cv::COLOR_GRAY2BGR
