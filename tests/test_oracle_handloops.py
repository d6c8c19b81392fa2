"""Pins the oracle's restatements of the reference's HAND-WRITTEN loops (not OpenCV calls) against
independent numpy / cv2-composed restatements written from the reference source:
  combineLazyViews + renderLazyNovelView   NovelView.cpp:101-224
  flattenLayersDeghostPreferBase           CvUtil.cpp:224-260
  iirLowPass + sharpenWithIirLowPass       Filter.h:39-137
  cubemap map + stacking                   ImageWarper.cpp:26-141, CvUtil.cpp:117-138
  featherAlphaChannel                      CvUtil.cpp:140-157
Float-to-uchar conversions truncate; libm calls (tanhf / exp / expf / acosf) are the container's."""
import ctypes
import math

import numpy as np
import pytest

import oracle_bind as ob

cv2 = pytest.importorskip("cv2")
cv2.ipp.setUseIPP(False)
cv2.setUseOptimized(False)
cv2.setNumThreads(1)
F = np.float32
libm = ctypes.CDLL("libm.so.6")
for _n in ("tanhf", "expf", "acosf", "powf"):
    getattr(libm, _n).restype = ctypes.c_float
libm.tanhf.argtypes = [ctypes.c_float]
libm.expf.argtypes = [ctypes.c_float]
libm.acosf.argtypes = [ctypes.c_float]
libm.powf.argtypes = [ctypes.c_float, ctypes.c_float]
rng = np.random.default_rng(99)


def lerp(a, b, t):
    return F(F(a * F(F(1.0) - t)) + F(b * t))


def test_flatten_deghost_against_scalar_restatement():
    base = rng.integers(0, 256, (24, 40, 4), dtype=np.uint8)
    top = rng.integers(0, 256, (24, 40, 4), dtype=np.uint8)
    top[:, :10, :3] = base[:, :10, :3]
    top[:8, :, 3] = 0
    out = np.zeros_like(base)
    for y in range(24):
        for x in range(40):
            b, t = base[y, x].astype(int), top[y, x].astype(int)
            cd = F(F(abs(b[0] - t[0]) + abs(b[1] - t[1]) + abs(b[2] - t[2])) / F(255.0))
            d = F(libm.tanhf(F(cd * F(5.0))))
            aR = F(F(t[3]) / F(255.0))
            aL = F(F(1.0) - aR)
            eL = float(libm.expf(F(F(F(5.0) * aL) * F(2.0))))  # float overload of exp (C++ <math.h> + using namespace std)
            eR = float(libm.expf(F(F(5.0) * aR)))
            sL = F(eL / (eL + eR + 0.00001))
            sR = F(F(1.0) - sL)
            wL, wR = lerp(aL, sL, d), lerp(aR, sR, d)
            for c in range(3):
                out[y, x, c] = int(F(F(F(b[c]) * wL) + F(F(t[c]) * wR)))
            out[y, x, 3] = max(t[3], b[3])
    assert np.array_equal(ob.flatten_deghost(base, top), out)


def test_sharpen_against_scalar_restatement():
    img = rng.integers(0, 256, (13, 17, 3), dtype=np.uint8)
    h, w = img.shape[:2]
    alpha = F(libm.powf(F(0.25), F(0.25)))
    lp = np.zeros_like(img)

    def wrap(x, r):
        return r + x if x < 0 else (x - r if x >= r else x)

    def refl(x, r):
        return -x if x < 0 else (2 * r - x - 2 if x >= r else x)

    def clampf(v):
        return F(0.0) if v < 0 else (F(255.0) if v > 255 else v)

    for c in range(3):
        for i in range(h):
            buf = np.zeros(w, np.float32)
            v = F(img[i, 0, c])
            for j in range(1, w + 1):
                v = lerp(F(img[i, wrap(j, w), c]), v, alpha)
                buf[wrap(j - 1, w)] = v
            for j in range(w - 2, -2, -1):
                v = lerp(buf[wrap(j, w)], v, alpha)
                lp[i, j + 1, c] = int(clampf(v))
        for j in range(w):
            buf = np.zeros(h, np.float32)
            v = F(lp[0, j, c])
            for i in range(1, h + 1):
                v = lerp(F(lp[refl(i, h), j, c]), v, alpha)
                buf[refl(i - 1, h)] = v
            for i in range(h - 2, -2, -1):
                v = lerp(buf[refl(i, h)], v, alpha)
                lp[i + 1, j, c] = int(clampf(v))
    out = img.copy()
    amount = F(1.0 + 0.25)
    for idx in np.ndindex(h, w, 3):
        l = F(lp[idx])
        hp = F(F(img[idx]) - l)
        ng = F(F(1.0) - F(libm.expf(F(-F(F(hp * hp) * F(100.0))))))
        out[idx] = int(clampf(F(l + F(F(hp * ng) * amount))))
    assert np.array_equal(ob.sharpen(img, 0.25), out)


def test_cubemap_against_cv2_remap_of_scalar_map():
    eye = rng.integers(0, 256, (32, 64, 3), dtype=np.uint8)
    fw = fh = 8
    order = ["RIGHT", "LEFT", "TOP", "BOTTOM", "BACK", "FRONT"]
    imgs = []
    for name in order:
        m = np.zeros((fh, fw, 2), np.float32)
        for j in range(fh):
            for i in range(fw):
                x = F(F(F(i) * F(F(1.0) / F(fw))) - F(0.5))
                y = F(F(F(j) * F(F(1.0) / F(fh))) - F(0.5))
                z = F(0.5)
                d = {"BACK": (x, z, -y), "LEFT": (-z, x, -y), "TOP": (x, y, z), "BOTTOM": (x, -y, -z),
                     "FRONT": (-x, -z, -y), "RIGHT": (z, -x, -y)}[name]
                d = [F(v) for v in d]
                r = F(np.sqrt(F(F(d[0] * d[0]) + F(d[1] * d[1]))))
                nrm = math.sqrt(float(d[0]) ** 2 + float(d[1]) ** 2 + float(d[2]) ** 2)
                phi = F(libm.acosf(F(float(d[2]) / nrm)))
                theta = F(libm.acosf(F(abs(F(d[0] / r))))) if r > 0 else F(0.0)
                if d[0] > 0 and d[1] > 0:
                    pass
                elif d[0] <= 0 and d[1] > 0:
                    theta = F(math.pi - float(theta))
                elif d[0] <= 0 and d[1] <= 0:
                    theta = F(math.pi + float(theta))
                else:
                    theta = F(2 * math.pi - float(theta))
                fov = F(math.pi)
                phiP = min(max(phi, F(0.0)), fov)
                thP = min(max(theta, F(0.0)), F(2.0 * math.pi))
                m[j, i, 0] = F(float(F(F(64) * thP)) / (2.0 * math.pi))
                m[j, i, 1] = F(F(F(32) * phiP) / fov)
        imgs.append(cv2.remap(eye, m, None, cv2.INTER_CUBIC, borderMode=cv2.BORDER_WRAP))
    video = np.vstack([np.hstack([imgs[1][:, ::-1], imgs[0][:, ::-1], imgs[2][:, ::-1]]),
                       np.hstack([imgs[3][:, ::-1], imgs[4][:, ::-1], imgs[5][:, ::-1]])])
    assert np.array_equal(ob.cubemap(eye, fw, fh, False), video)
    assert np.array_equal(ob.cubemap(eye, fw, fh, True), np.vstack(imgs))


def test_feather_alpha_against_cv2_composition():
    f = rng.integers(0, 256, (90, 140, 4), dtype=np.uint8)
    f[..., 3] = (rng.uniform(0, 1, (90, 140)) > 0.03) * 255
    a = cv2.erode(f[..., 3].copy(), cv2.getStructuringElement(cv2.MORPH_CROSS, (63, 63), (31, 31)))
    a = cv2.GaussianBlur(a, (31, 31), 31 / 2.0)
    ref = f.copy()
    ref[..., 3] = a
    assert np.array_equal(ob.feather_alpha(f, 31), ref)


def test_novel_view_against_cv2_composition():
    from surround360_b200 import synth
    i0, i1 = synth.pair_512()
    L, R = np.ascontiguousarray(i0[:60, :90]), np.ascontiguousarray(i1[:60, :90])
    fLR = (rng.standard_normal((60, 90, 2)) * 2).astype(np.float32)
    fRL = (rng.standard_normal((60, 90, 2)) * 2).astype(np.float32)
    buf_w, cam_w, n_novel, disp = 20, 120, 20, F(7.25)
    h, w = L.shape[:2]

    def lazy(src, flow, sign, invert):
        wx = np.array([F(F(F(F(cam_w) * F(0.5)) - F(n_novel - u)) + F(F(sign) * disp)) for u in range(buf_w)], np.float32)
        t = np.array([F(F(u) / F(n_novel)) for u in range(buf_w)], np.float32)
        if invert:
            t = (F(1.0) - t).astype(np.float32)
        m = np.zeros((h, buf_w, 2), np.float32)
        m[..., 0] = wx[None, :]
        m[..., 1] = np.arange(h, dtype=np.float32)[:, None]
        rf = cv2.remap(flow, m, None, cv2.INTER_CUBIC)
        comp = np.stack([m[..., 0] + rf[..., 0] * t[None, :], m[..., 1] + rf[..., 1] * t[None, :]], -1).astype(np.float32)
        nv = cv2.remap(src, comp, None, cv2.INTER_CUBIC)
        nv[..., 3] = ((F(1.0) - t)[None, :] * nv[..., 3].astype(np.float32)).astype(np.int32).astype(np.uint8)
        return nv, np.sqrt(rf[..., 0] * rf[..., 0] + rf[..., 1] * rf[..., 1]).astype(np.float32)

    def combine(a, b, ma, mb):
        out = np.zeros_like(a)
        for y in range(h):
            for x in range(buf_w):
                cl, cr = a[y, x].astype(int), b[y, x].astype(int)
                oa = 255 if float(F(F(max(cl[3], cr[3])) / F(255.0))) > 0.1 else 0
                if cl[3] == 0 and cr[3] == 0:
                    out[y, x] = (0, 0, 0, oa)
                elif cl[3] == 0:
                    out[y, x] = (cr[0], cr[1], cr[2], oa)
                elif cr[3] == 0:
                    out[y, x] = (cl[0], cl[1], cl[2], oa)
                else:
                    mgl, mgr = F(ma[y, x] / F(buf_w)), F(mb[y, x] / F(buf_w))
                    bl, br = F(cl[3]), F(cr[3])
                    nrm = F(bl + br)
                    bl = F(bl / nrm)
                    br = F(br / nrm)
                    cd = F(F(abs(cl[0] - cr[0]) + abs(cl[1] - cr[1]) + abs(cl[2] - cr[2])) / F(255.0))
                    d = F(libm.tanhf(F(cd * F(10.0))))
                    eL = math.exp(float(F(F(10.0) * bl)) * (1.0 + float(F(F(20.0) * mgl))))
                    eR = math.exp(float(F(F(10.0) * br)) * (1.0 + float(F(F(20.0) * mgr))))
                    s = eL + eR + 0.00001
                    wl, wr = lerp(bl, F(eL / s), d), lerp(br, F(eR / s), d)
                    out[y, x] = [int(F(F(F(cl[c]) * wl) + F(F(cr[c]) * wr))) for c in range(3)] + [255]
        return out

    ref = []
    for sign in (+1, -1):
        a, ma = lazy(L, fRL, sign, False)
        b, mb = lazy(R, fLR, sign, True)
        ref.append(combine(a, b, ma, mb))
    gl, gr = ob.novel_view_chunk(L, R, fLR, fRL, buf_w, cam_w, n_novel, float(disp))
    assert np.array_equal(gl, ref[0]) and np.array_equal(gr, ref[1])
