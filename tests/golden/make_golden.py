"""Generates tests/golden/*.npz in THIS container (python cv2 4.13, IPP off, 1 thread).

cv_vectors.npz   : seeded inputs + cv2 outputs for every OpenCV primitive the path calls
                   (the reference delegates them to OpenCV, absent from /root/reference).
pixflow_o1.npz   : PixFlow restated in Python with cv2 calls for exactly the calls the reference
                   makes (PixFlow.h:81-534) and numba for its hand-written loops, run on a small
                   pair -- an independent second restatement that pins the C++ oracle end to end.
Run:  python tests/golden/make_golden.py
"""
import os
import sys

import cv2
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
cv2.ipp.setUseIPP(False)
cv2.setUseOptimized(False)
cv2.setNumThreads(1)


def cv_vectors():
    rng = np.random.default_rng(20260922)
    out = {}
    f1 = rng.standard_normal((53, 77)).astype(np.float32)
    f2 = rng.standard_normal((53, 77, 2)).astype(np.float32)
    out["f1"], out["f2"] = f1, f2
    for k, s in [(5, 0.25), (3, 0.5), (15, 8.0), (3, 1.0)]:
        tag = f"{k}_{int(s * 100)}"
        out[f"gk_{tag}"] = cv2.getGaussianKernel(k, s, cv2.CV_32F).ravel()
        out[f"blur1_{tag}"] = cv2.GaussianBlur(f1, (k, k), s)
        out[f"blur2_{tag}"] = cv2.GaussianBlur(f2, (k, k), s)
    out["sobel_x"] = cv2.Sobel(f1, -1, 1, 0, ksize=1, scale=1, delta=0, borderType=cv2.BORDER_REPLICATE)
    out["sobel_y"] = cv2.Sobel(f1, -1, 0, 1, ksize=1, scale=1, delta=0, borderType=cv2.BORDER_REPLICATE)
    out["median5"] = cv2.medianBlur(f2, 5)
    out["lin1"] = cv2.resize(f1, (69, 48), interpolation=cv2.INTER_LINEAR)
    out["lin2_up"] = cv2.resize(f2, (154, 106), interpolation=cv2.INTER_LINEAR)
    out["cub2_up"] = cv2.resize(f2, (86, 59), interpolation=cv2.INTER_CUBIC)
    u4 = rng.integers(0, 256, (61, 83, 4), dtype=np.uint8)
    u3 = rng.integers(0, 256, (40, 111, 3), dtype=np.uint8)
    out["u4"], out["u3"] = u4, u3
    out["cub8_half"] = cv2.resize(u4, (41, 30), interpolation=cv2.INTER_CUBIC)
    out["cub8_c3"] = cv2.resize(u3, (108, 40), interpolation=cv2.INTER_CUBIC)
    m = np.stack([rng.uniform(-6, 89, (45, 70)), rng.uniform(-6, 67, (45, 70))], -1).astype(np.float32)
    out["map"] = m
    out["remap8_const"] = cv2.remap(u4, m, None, cv2.INTER_CUBIC, borderMode=cv2.BORDER_CONSTANT)
    m3 = np.stack([rng.uniform(-9, 120, (45, 70)), rng.uniform(-9, 49, (45, 70))], -1).astype(np.float32)
    out["map3"] = m3
    out["remap8_wrap"] = cv2.remap(u3, m3, None, cv2.INTER_CUBIC, borderMode=cv2.BORDER_WRAP)
    mf = np.stack([rng.uniform(-6, 83, (45, 70)), rng.uniform(-6, 59, (45, 70))], -1).astype(np.float32)
    out["mapf"] = mf
    out["remapf"] = cv2.remap(f2, mf, None, cv2.INTER_CUBIC)
    out["gray"] = cv2.cvtColor(u4, cv2.COLOR_BGRA2GRAY)
    a = ((rng.uniform(0, 1, (90, 140)) > 0.02) * 255).astype(np.uint8)
    out["alpha"] = a
    ker = cv2.getStructuringElement(cv2.MORPH_CROSS, (63, 63), (31, 31))
    out["erode"] = cv2.erode(a, ker)
    out["blur8"] = cv2.GaussianBlur(out["erode"], (31, 31), 15.5)
    rnd = rng.integers(0, 256, (90, 140), dtype=np.uint8)
    out["rnd8"] = rnd
    out["blur8_rnd"] = cv2.GaussianBlur(rnd, (31, 31), 15.5)
    # nearest + wrap semantics of offsetHorizontalWrap's map (CvUtil.cpp:93-115)
    img = rng.integers(0, 256, (3, 37, 4), dtype=np.uint8)
    out["wrap_img"] = img
    for i, off in enumerate([4.25, -4.25, 0.5, 1.5, 36.6]):
        xs = np.arange(37, dtype=np.float32) - np.float32(off)
        xs = np.where(xs < 0, xs + 37, xs)
        xs = np.where(xs >= 37, xs - 37, xs).astype(np.float32)
        mm = np.stack([np.tile(xs, (3, 1)), np.tile(np.arange(3, dtype=np.float32)[:, None], (1, 37))], -1)
        out[f"wrap_{i}"] = cv2.remap(img, mm, None, cv2.INTER_NEAREST, borderMode=cv2.BORDER_WRAP)
    out["wrap_offsets"] = np.array([4.25, -4.25, 0.5, 1.5, 36.6], np.float32)
    np.savez_compressed(os.path.join(HERE, "cv_vectors.npz"), **out)
    print("cv_vectors.npz written")


def pixflow_o1():
    import numba

    F = np.float32

    @numba.njit(cache=False, fastmath=False)
    def bil(img, x, y):
        h, w = img.shape
        x = min(F(w) - F(2.0), max(F(0.0), x))
        y = min(F(h) - F(2.0), max(F(0.0), y))
        x0 = int(x); y0 = int(y)
        xR = F(x - F(x0)); yR = F(y - F(y0))
        f00 = img[y0, x0]; f01 = img[y0 + 1, x0]; f10 = img[y0, x0 + 1]; f11 = img[y0 + 1, x0 + 1]
        a1 = f00; a2 = F(f10 - f00); a3 = F(f01 - f00); a4 = F(F(F(f00 + f11) - f10) - f01)
        return F(F(F(a1 + F(a2 * xR)) + F(a3 * yR)) + F(F(a4 * xR) * yR))

    @numba.njit(cache=False, fastmath=False)
    def err(I0x, I0y, I1x, I1y, bl, x, y, fx, fy):
        h, w = I0x.shape
        i1x = bil(I1x, F(F(x) + fx), F(F(y) + fy)); i1y = bil(I1y, F(F(x) + fx), F(F(y) + fy))
        dx = F(bl[y, x, 0] - fx); dy = F(bl[y, x, 1] - fy)
        sm = F(np.sqrt(F(F(dx * dx) + F(dy * dy))))
        ex = F(I0x[y, x] - i1x); ey = F(I0y[y, x] - i1y)
        e = F(np.sqrt(F(F(ex * ex) + F(ey * ey))))
        e = F(e + F(sm * F(0.001)))
        e = F(e + F(F(F(0.01) * abs(fy)) / F(w)))
        e = F(e + F(F(F(0.01) * abs(fx)) / F(h)))
        return e

    @numba.njit(cache=False, fastmath=False)
    def sweep(a0, a1, I0x, I0y, I1x, I1y, flow, bl, d):
        h, w = a0.shape
        ys = range(h) if d > 0 else range(h - 1, -1, -1)
        for y in ys:
            xs = range(w) if d > 0 else range(w - 1, -1, -1)
            for x in xs:
                if a0[y, x] > F(0.9) and a1[y, x] > F(0.9):
                    cur = err(I0x, I0y, I1x, I1y, bl, x, y, flow[y, x, 0], flow[y, x, 1])
                    xn = x - d
                    if 0 <= xn < w:
                        px = flow[y, xn, 0]; py = flow[y, xn, 1]
                        e = err(I0x, I0y, I1x, I1y, bl, x, y, px, py)
                        if e < cur:
                            flow[y, x, 0] = px; flow[y, x, 1] = py; cur = e
                    yn = y - d
                    if 0 <= yn < h:
                        px = flow[yn, x, 0]; py = flow[yn, x, 1]
                        e = err(I0x, I0y, I1x, I1y, bl, x, y, px, py)
                        if e < cur:
                            flow[y, x, 0] = px; flow[y, x, 1] = py; cur = e
                    fx = flow[y, x, 0]; fy = flow[y, x, 1]
                    gx = F(F(err(I0x, I0y, I1x, I1y, bl, x, y, F(fx + F(0.001)), fy) - cur) / F(0.001))
                    gy = F(F(err(I0x, I0y, I1x, I1y, bl, x, y, fx, F(fy + F(0.001))) - cur) / F(0.001))
                    flow[y, x, 0] = F(fx - F(F(0.5) * gx)); flow[y, x, 1] = F(fy - F(F(0.5) * gy))

    def pyramid(img):
        pyr = [img]
        while True:
            nw = int(np.float32(pyr[-1].shape[1]) * np.float32(0.9) + np.float32(0.5))
            nh = int(np.float32(pyr[-1].shape[0]) * np.float32(0.9) + np.float32(0.5))
            if nh <= 24 or nw <= 24:
                break
            pyr.append(cv2.resize(pyr[-1], (nw, nh), interpolation=cv2.INTER_LINEAR))
        return pyr

    def prep(rgba, dw, dh):
        small = cv2.resize(rgba, (dw, dh), interpolation=cv2.INTER_CUBIC)
        inv = np.float32(1.0 / 255.0)
        I = cv2.cvtColor(small, cv2.COLOR_BGRA2GRAY).astype(np.float32) * inv
        A = small[..., 3].astype(np.float32) * inv
        return cv2.GaussianBlur(I, (5, 5), 0.25), A

    def level(I0, I1, a0, a1, flow):
        sob = lambda im, dx, dy: cv2.Sobel(im, -1, dx, dy, ksize=1, scale=1, delta=0, borderType=cv2.BORDER_REPLICATE)
        g = [cv2.GaussianBlur(sob(im, dx, dy), (3, 3), 0.5) for im, dx, dy in ((I0, 1, 0), (I0, 0, 1), (I1, 1, 0), (I1, 0, 1))]
        if flow is None:
            flow = np.zeros(I0.shape + (2,), np.float32)
        bl = cv2.GaussianBlur(flow, (15, 15), 8.0)
        sweep(a0, a1, g[0], g[1], g[2], g[3], flow, bl, 1)
        flow = cv2.medianBlur(flow, 5)
        sweep(a0, a1, g[0], g[1], g[2], g[3], flow, bl, -1)
        flow = cv2.medianBlur(flow, 5)
        b2 = cv2.GaussianBlur(flow, (15, 15), 8.0)
        d = (np.float32(1.0) - a0 * a1)[..., None]
        return (d * b2 + (np.float32(1.0) - d) * flow).astype(np.float32)

    def pixflow(i0, i1):
        h, w = i0.shape[:2]
        dw, dh = int(np.float32(w) * np.float32(0.5)), int(np.float32(h) * np.float32(0.5))
        I0, A0 = prep(i0, dw, dh)
        I1, A1 = prep(i1, dw, dh)
        p = [pyramid(x) for x in (I0, I1, A0, A1)]
        flow = None
        for l in range(len(p[0]) - 1, -1, -1):
            flow = level(p[0][l], p[1][l], p[2][l], p[3][l], flow)
            if l > 0:
                flow = cv2.resize(flow, (p[0][l - 1].shape[1], p[0][l - 1].shape[0]), interpolation=cv2.INTER_CUBIC)
                flow = flow * (np.float32(1.0) / np.float32(0.9))
        flow = cv2.resize(flow, (w, h), interpolation=cv2.INTER_LINEAR) * np.float32(2.0)
        return cv2.GaussianBlur(flow, (3, 3), 1.0)

    from surround360_b200 import synth
    i0, i1 = synth.pair_512()
    a = np.ascontiguousarray(i0[40:200, 30:230])
    b = np.ascontiguousarray(i1[40:200, 30:230])
    a[:, 100:110, 3] = 0   # exercise the alpha gate and the diffusion
    flow = pixflow(a, b)
    np.savez_compressed(os.path.join(HERE, "pixflow_o1.npz"), i0=a, i1=b, flow=flow)
    print("pixflow_o1.npz written", flow.shape, float(np.abs(flow).max()))


if __name__ == "__main__":
    cv_vectors()
    if "--cv-only" not in sys.argv:
        pixflow_o1()
