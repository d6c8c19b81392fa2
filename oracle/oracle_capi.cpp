// ORACLE — TEST INFRASTRUCTURE ONLY (see ocv.h header). C ABI used by tests/ via ctypes.
#include "ocv.h"
using namespace orc;
extern "C" {

static ImgF wrapF(const float* p, int w, int h, int c) { ImgF m(w, h, c); memcpy(m.d.data(), p, m.d.size() * 4); return m; }
static Img8 wrap8(const uint8_t* p, int w, int h, int c) { Img8 m(w, h, c); memcpy(m.d.data(), p, m.d.size()); return m; }

void o_gauss_kernel(int n, double sigma, float* out) { auto k = gaussKernel(n, sigma); memcpy(out, k.data(), n * 4); }
void o_gaussian_blur_f32(const float* s, int w, int h, int c, int k, double sigma, float* d) {
  ImgF a = wrapF(s, w, h, c), b; gaussianBlurF(a, b, k, sigma); memcpy(d, b.d.data(), b.d.size() * 4);
}
void o_sobel1(const float* s, int w, int h, int dx, float* d) { ImgF a = wrapF(s, w, h, 1), b; sobel1(a, b, dx != 0); memcpy(d, b.d.data(), b.d.size() * 4); }
void o_median5(const float* s, int w, int h, int c, float* d) { ImgF a = wrapF(s, w, h, c), b; median5(a, b); memcpy(d, b.d.data(), b.d.size() * 4); }
void o_resize_linear_f32(const float* s, int w, int h, int c, int dw, int dh, float* d) {
  ImgF a = wrapF(s, w, h, c), b; resizeLinearF(a, b, dw, dh); memcpy(d, b.d.data(), b.d.size() * 4);
}
void o_resize_cubic_f32(const float* s, int w, int h, int c, int dw, int dh, float* d) {
  ImgF a = wrapF(s, w, h, c), b; resizeCubicF(a, b, dw, dh); memcpy(d, b.d.data(), b.d.size() * 4);
}
void o_resize_cubic_u8(const uint8_t* s, int w, int h, int c, int dw, int dh, uint8_t* d) {
  Img8 a = wrap8(s, w, h, c), b; resizeCubic8(a, b, dw, dh); memcpy(d, b.d.data(), b.d.size());
}
void o_remap_cubic_u8(const uint8_t* s, int w, int h, int c, const float* map, int dw, int dh, int border, uint8_t* d) {
  Img8 a = wrap8(s, w, h, c), b; ImgF m = wrapF(map, dw, dh, 2); remapCubic8(a, m, b, border); memcpy(d, b.d.data(), b.d.size());
}
void o_remap_cubic_f32(const float* s, int w, int h, int c, const float* map, int dw, int dh, float* d) {
  ImgF a = wrapF(s, w, h, c), b; ImgF m = wrapF(map, dw, dh, 2); remapCubicF(a, m, b); memcpy(d, b.d.data(), b.d.size() * 4);
}
void o_bgra2gray(const uint8_t* s, int n, int c, int legacy, uint8_t* d) { for (int i = 0; i < n; ++i) d[i] = gray8(s[i * c], s[i * c + 1], s[i * c + 2], legacy != 0); }
void o_erode_cross(const uint8_t* s, int w, int h, int r, uint8_t* d) { Img8 a = wrap8(s, w, h, 1), b; erodeCross8(a, b, r); memcpy(d, b.d.data(), b.d.size()); }
void o_gauss_kernel_fixed8(int n, double sigma, int* out) { auto k = gaussKernelFixed8(n, sigma); memcpy(out, k.data(), n * 4); }
void o_gaussian_blur_u8(const uint8_t* s, int w, int h, int k, double sigma, uint8_t* d) {
  Img8 a = wrap8(s, w, h, 1), b; gaussianBlur8Fixed(a, b, gaussKernelFixed8(k, sigma)); memcpy(d, b.d.data(), b.d.size());
}
int o_offset_wrap_index(int x, int cols, float offset) { return offsetWrapIndex(x, cols, offset); }
}
