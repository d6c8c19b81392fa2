"""Pins the CPU oracle's OpenCV restatements (oracle/ocv.h) against golden vectors produced by
python cv2 4.13 (IPP off) in the build container (tests/golden/make_golden.py). Bit-exact."""
import ctypes as C
import os

import numpy as np
import pytest

import oracle_bind as ob
from conftest import GOLDEN

G = np.load(os.path.join(GOLDEN, "cv_vectors.npz"))
L = ob.lib()
P = ob.P


def eq(a, b):
    assert a.shape == b.shape and np.array_equal(a, b), f"{np.count_nonzero(a != b)} of {a.size} differ"


@pytest.mark.parametrize("k,s", [(5, 0.25), (3, 0.5), (15, 8.0), (3, 1.0)])
def test_gaussian_kernel_and_blur(k, s):
    tag = f"{k}_{int(s * 100)}"
    kk = np.zeros(k, np.float32)
    L.o_gauss_kernel(k, C.c_double(s), P(kk))
    eq(kk, G[f"gk_{tag}"])
    for name, c in (("f1", 1), ("f2", 2)):
        a = np.ascontiguousarray(G[name])
        o = np.zeros_like(a)
        L.o_gaussian_blur_f32(P(a), a.shape[1], a.shape[0], c, k, C.c_double(s), P(o))
        eq(o, G[f"blur{c}_{tag}"])


def test_sobel_median():
    a = np.ascontiguousarray(G["f1"])
    o = np.zeros_like(a)
    L.o_sobel1(P(a), a.shape[1], a.shape[0], 1, P(o)); eq(o, G["sobel_x"])
    L.o_sobel1(P(a), a.shape[1], a.shape[0], 0, P(o)); eq(o, G["sobel_y"])
    b = np.ascontiguousarray(G["f2"])
    o2 = np.zeros_like(b)
    L.o_median5(P(b), b.shape[1], b.shape[0], 2, P(o2)); eq(o2, G["median5"])


def test_resize_float():
    a = np.ascontiguousarray(G["f1"]); b = np.ascontiguousarray(G["f2"])
    o = np.zeros((48, 69), np.float32)
    L.o_resize_linear_f32(P(a), 77, 53, 1, 69, 48, P(o)); eq(o, G["lin1"])
    o = np.zeros((106, 154, 2), np.float32)
    L.o_resize_linear_f32(P(b), 77, 53, 2, 154, 106, P(o)); eq(o, G["lin2_up"])
    o = np.zeros((59, 86, 2), np.float32)
    L.o_resize_cubic_f32(P(b), 77, 53, 2, 86, 59, P(o)); eq(o, G["cub2_up"])


def test_resize_cubic_u8():
    eq(ob.resize_cubic_u8(G["u4"], 41, 30), G["cub8_half"])
    eq(ob.resize_cubic_u8(G["u3"], 108, 40), G["cub8_c3"])


def test_remap():
    u4, u3, f2 = (np.ascontiguousarray(G[k]) for k in ("u4", "u3", "f2"))
    m, m3, mf = (np.ascontiguousarray(G[k]) for k in ("map", "map3", "mapf"))
    o = np.zeros((45, 70, 4), np.uint8)
    L.o_remap_cubic_u8(P(u4), 83, 61, 4, P(m), 70, 45, 0, P(o)); eq(o, G["remap8_const"])
    o = np.zeros((45, 70, 3), np.uint8)
    L.o_remap_cubic_u8(P(u3), 111, 40, 3, P(m3), 70, 45, 3, P(o)); eq(o, G["remap8_wrap"])
    o = np.zeros((45, 70, 2), np.float32)
    L.o_remap_cubic_f32(P(f2), 77, 53, 2, P(mf), 70, 45, P(o)); eq(o, G["remapf"])


def test_gray_erode_blur8():
    u4 = np.ascontiguousarray(G["u4"])
    o = np.zeros(u4.shape[:2], np.uint8)
    L.o_bgra2gray(P(u4), o.size, 4, 0, P(o)); eq(o, G["gray"])
    a = np.ascontiguousarray(G["alpha"])
    e = np.zeros_like(a)
    L.o_erode_cross(P(a), a.shape[1], a.shape[0], 31, P(e)); eq(e, G["erode"])
    b = np.zeros_like(a)
    L.o_gaussian_blur_u8(P(e), a.shape[1], a.shape[0], 31, C.c_double(15.5), P(b)); eq(b, G["blur8"])
    r = np.ascontiguousarray(G["rnd8"])
    L.o_gaussian_blur_u8(P(r), r.shape[1], r.shape[0], 31, C.c_double(15.5), P(b)); eq(b, G["blur8_rnd"])


def test_offset_wrap_nearest_index_map():
    img = np.ascontiguousarray(G["wrap_img"])
    for i, off in enumerate(G["wrap_offsets"]):
        eq(ob.offset_wrap(img, float(off)), G[f"wrap_{i}"])


def test_live_cv2_agrees_when_available():
    cv2 = pytest.importorskip("cv2")
    cv2.ipp.setUseIPP(False); cv2.setUseOptimized(False); cv2.setNumThreads(1)
    rng = np.random.default_rng(7)
    a = rng.integers(0, 256, (97, 131, 4), dtype=np.uint8)
    eq(ob.resize_cubic_u8(a, 65, 48), cv2.resize(a, (65, 48), interpolation=cv2.INTER_CUBIC))
    f = rng.standard_normal((40, 60, 2)).astype(np.float32)
    o = np.zeros_like(f)
    L.o_gaussian_blur_f32(P(f), 60, 40, 2, 15, C.c_double(8.0), P(o))
    eq(o, cv2.GaussianBlur(f, (15, 15), 8.0))
