// expf with the bits of the host C library, as device code (shared by the soft ISP's noise coring and the preview's softmax blend).
#pragma once
#include <cuda_runtime.h>

namespace s360 {

// ------------------------------------------------------------------------------------------------------------------------------
// expf with the bits of the C library (glibc >= 2.27, sysdeps/ieee754/flt-32/e_expf.c: Szabolcs Nagy's exp2f-table algorithm --
// double arithmetic, 32-entry table, one final rounding to float). The unsharp mask's noise coring evaluates expf on arbitrary
// floats (CameraIsp.h:1244-1259 -> Filter.h:112-118); CUDA's expf may differ from the host's in the last place, which the 8/16-bit
// truncation after it turns into rare 1-LSB differences. The same IEEE double operations give the same bits on the device.
// Table: T[i] = bits(2^(i/32)) - (i << 47); pinned against the host's expf in tests/test_isp_tables.py.
// ------------------------------------------------------------------------------------------------------------------------------
static __device__ const unsigned long long kExp2fTab[32] = {
    0x3ff0000000000000ull, 0x3fefd9b0d3158574ull, 0x3fefb5586cf9890full, 0x3fef9301d0125b51ull, 0x3fef72b83c7d517bull, 0x3fef54873168b9aaull,
    0x3fef387a6e756238ull, 0x3fef1e9df51fdee1ull, 0x3fef06fe0a31b715ull, 0x3feef1a7373aa9cbull, 0x3feedea64c123422ull, 0x3feece086061892dull,
    0x3feebfdad5362a27ull, 0x3feeb42b569d4f82ull, 0x3feeab07dd485429ull, 0x3feea47eb03a5585ull, 0x3feea09e667f3bcdull, 0x3fee9f75e8ec5f74ull,
    0x3feea11473eb0187ull, 0x3feea589994cce13ull, 0x3feeace5422aa0dbull, 0x3feeb737b0cdc5e5ull, 0x3feec49182a3f090ull, 0x3feed503b23e255dull,
    0x3feee89f995ad3adull, 0x3feeff76f2fb5e47ull, 0x3fef199bdd85529cull, 0x3fef3720dcef9069ull, 0x3fef5818dcfba487ull, 0x3fef7c97337b9b5full,
    0x3fefa4afa2a490daull, 0x3fefd0765b6e4540ull};
__device__ __forceinline__ float expfHostBits(float x) {
  const unsigned ux = __float_as_uint(x);
  const unsigned abstop = (ux >> 20) & 0x7ffu;
  if (abstop >= 0x42bu) {                                            // |x| >= 88 or NaN
    if (ux == 0xff800000u) return 0.0f;
    if (abstop >= 0x7f8u) return x + x;
    if (x > __uint_as_float(0x42b17217u)) return __uint_as_float(0x7f800000u);   // x > log(2^128): overflow
    if (x < __uint_as_float(0xc2cff1b4u)) return 0.0f;                           // x < log(2^-150): underflow
  }
  const double invLn2N = __longlong_as_double(0x40471547652b82feLL);             // 0x1.71547652b82fep+0 * 32
  const double shift = __longlong_as_double(0x4338000000000000LL);               // 0x1.8p+52
  const double c0 = __longlong_as_double(0x3ebc6af84b912394LL);                  // 0x1.c6af84b912394p-5 / 32^3
  const double c1 = __longlong_as_double(0x3f2ebfce50fac4f3LL);                  // 0x1.ebfce50fac4f3p-3 / 32^2
  const double c2 = __longlong_as_double(0x3f962e42ff0c52d6LL);                  // 0x1.62e42ff0c52d6p-1 / 32
  const double z = invLn2N * (double)x;
  double kd = z + shift;
  const unsigned long long ki = (unsigned long long)__double_as_longlong(kd);
  kd = kd - shift;
  const double r = z - kd;
  unsigned long long t = kExp2fTab[ki & 31u];
  t += ki << 47;
  const double s = __longlong_as_double((long long)t);
  const double zz = c0 * r + c1;
  const double r2 = r * r;
  double y = c2 * r + 1.0;
  y = zz * r2 + y;
  y = y * s;
  return (float)y;
}

}  // namespace s360
