"""CPU tier of the soft ISP: everything the product derives on the HOST -- composite colour matrix, tone-curve LUT, vignette
gains (surround360_b200/csrc/isp_kernels.cu, IspPipeline::setup / buildToneCurveLut / tables; needs no GPU) -- against the REFERENCE'S
OWN CameraIsp::setup() compiled unchanged (oracle/_ref/libref_isp.so), the stand-in matrix product of the reference build against
cv2.gemm, and the exact expf port (expfHostBits) against the C library. No GPU compute here."""
import ctypes as C
import glob
import json
import os
import subprocess
import sys
import tempfile

import numpy as np
import pytest

import ref_bind as rb
from surround360_b200 import api

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CONFIGS = json.load(open(os.path.join(HERE, "golden", "isp_configs.json")))
REF_ISP_DIR = "/root/reference/surround360_render/res/config/isp"


def all_configs():
    out = dict(CONFIGS)
    for p in sorted(glob.glob(os.path.join(REF_ISP_DIR, "*.json"))):   # the reference's own files where the reference tree exists
        out["reference:" + os.path.basename(p)] = open(p).read()
    return out


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


@pytest.mark.skipif(not rb.available(), reason="oracle/_ref not built (no /root/reference and no prebuilt copy)")
@pytest.mark.parametrize("bpp", [8, 16])
@pytest.mark.parametrize("tone", [True, False])
def test_host_tables_match_reference_setup(bpp, tone):
    for name, text in all_configs().items():
        isp = api.CameraIsp(text, bpp)
        (isp.enableToneMap if tone else isp.disableToneMap)()
        for (w, h) in ((200, 136), (97, 131), (2048, 2048)):
            got = isp.tables(w, h)
            ref = rb.isp_tables(text, w, h, bpp, tone)
            for g, r, what in zip(got, ref, ("compositeCCM", "toneCurveLut", "vignetteH", "vignetteV")):
                assert np.array_equal(bits(g), bits(r)), f"{name} {what} {w}x{h} bpp {bpp} tone {tone}"
        isp.close()


def test_composite_ccm_against_cv2_gemm():
    """The stand-in `Mat * Mat` of the reference build (oracle/ref_shim/opencv2/core.hpp) and the product's mat3Mul restate cv::gemm's
    fixed-size branch: pin both against the real cv2.gemm on the matrices setup() multiplies (CameraIsp.h:675-692)."""
    cv2 = pytest.importorskip("cv2")
    rgb2yuv = np.array([0.299, 0.587, 0.114, -0.14713, -0.28886, 0.436, 0.615, -0.51499, -0.10001], np.float32).reshape(3, 3)
    yuv2rgb = np.array([1.0, 0.0, 1.13983, 1.0, -0.39465, -0.58060, 1.0, 2.03211, 0.0], np.float32).reshape(3, 3)
    for name, text in all_configs().items():
        cfg = json.loads(text).get("CameraIsp", {})
        ccm = np.array(cfg.get("ccm", np.eye(3)), np.float64).astype(np.float32)
        sat = np.float32(cfg.get("saturation", 1.0))
        sm = np.zeros((3, 3), np.float32); sm[0, 0] = 1; sm[1, 1] = sat; sm[2, 2] = sat
        sm = cv2.gemm(cv2.gemm(yuv2rgb, sm, 1, None, 0), rgb2yuv, 1, None, 0)
        comp = cv2.gemm(cv2.transpose(ccm), sm, 1, None, 0) * np.float32(4095.0)
        isp = api.CameraIsp(text, 8)
        assert np.array_equal(bits(isp.tables(16, 16)[0]), bits(comp)), name
        if rb.available():
            assert np.array_equal(bits(rb.isp_tables(text, 16, 16)[0]), bits(comp)), name


EXPF_TEST = r'''
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#define __device__
#define __forceinline__ inline
static inline unsigned __float_as_uint(float f) { unsigned u; std::memcpy(&u, &f, 4); return u; }
static inline float __uint_as_float(unsigned u) { float f; std::memcpy(&f, &u, 4); return f; }
static inline double __longlong_as_double(long long v) { double d; std::memcpy(&d, &v, 8); return d; }
static inline long long __double_as_longlong(double d) { long long v; std::memcpy(&v, &d, 8); return v; }
%s
int main() {
  // every 61st float bit pattern of both signs below |x| = 128 (the band outside is constant 0 / inf), plus NaN / inf
  unsigned long long n = 0, bad = 0;
  for (uint32_t sign = 0; sign < 2; ++sign)
    for (uint32_t u = 0; u <= 0x43000000u; u += 61) {
      const float x = __uint_as_float(u | (sign << 31));
      const float a = expfHostBits(x), b = std::exp(x);
      ++n;
      if (__float_as_uint(a) != __float_as_uint(b)) { if (bad < 5) std::printf("x=%%a port=%%a libm=%%a\n", x, a, b); ++bad; }
    }
  const uint32_t special[4] = {0x7f800000u, 0xff800000u, 0x7fc00000u, 0xc2d00000u};
  for (uint32_t u : special) {
    const float x = __uint_as_float(u); const float a = expfHostBits(x), b = std::exp(x);
    if (!(a == b || (a != a && b != b))) { std::printf("special %%08x\n", u); ++bad; }
  }
  std::printf("%%llu %%llu\n", n, bad);
  return bad ? 1 : 0;
}
'''


def test_expf_port_matches_libm():
    """expfHostBits (the device function the unsharp mask's noise coring uses) is extracted from the CUDA source, compiled for the host
    and compared with the C library's expf on ~35 million float bit patterns: identical bits."""
    src = open(os.path.join(ROOT, "surround360_b200", "csrc", "expf_bits.cuh")).read()
    a = src.index("static __device__ const unsigned long long kExp2fTab[32]")
    b = src.index("}  // namespace s360")
    with tempfile.TemporaryDirectory() as d:
        cpp = os.path.join(d, "t.cpp"); exe = os.path.join(d, "t")
        open(cpp, "w").write(EXPF_TEST % src[a:b].replace("static __device__ const", "static const"))
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-o", exe, cpp])
        r = subprocess.run([exe], capture_output=True, text=True)
        assert r.returncode == 0, r.stdout[-600:]
        n, bad = map(int, r.stdout.strip().splitlines()[-1].split())
        assert n > 30_000_000 and bad == 0


def test_isp_object_semantics_without_gpu():
    """Constructor / setter errors mirror the reference (CameraIsp.h:968-989); no device is touched."""
    with pytest.raises(api.VrCamException, match="Demosaic filter"):
        api.CameraIsp(CONFIGS["sensor_a"], 8).setDemosaicFilter(-1)
    with pytest.raises(api.VrCamException, match="resize value"):
        api.CameraIsp(CONFIGS["sensor_a"], 8).setResize(16)
    with pytest.raises(api.VrCamException, match="unknown bayerPattern"):
        api.CameraIsp(json.dumps({"CameraIsp": {"bayerPattern": "XYZW"}}), 8)
    with pytest.raises(api.VrCamException, match="3 elements"):
        api.CameraIsp(json.dumps({"CameraIsp": {"blackLevel": [1.0, 2.0]}}), 8)
    isp = api.CameraIsp("{}", 16)   # no "CameraIsp" object: all defaults (CameraIsp.h:441-463, :604-606)
    ccm, lut, vh, vv = isp.tables(8, 8)
    assert np.array_equal(vh, np.ones((8, 3), np.float32)) and lut[-1, 0] == 65535.0
