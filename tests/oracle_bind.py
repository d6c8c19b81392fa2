"""ctypes binding of the CPU oracle (oracle/liboracle.so). TEST INFRASTRUCTURE ONLY:
imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference."""
import ctypes as C
import os
import subprocess
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_DIR = os.path.join(os.path.dirname(_HERE), "oracle")
_LIB = None


class ORenderParams(C.Structure):
    _fields_ = [("interpupilary_dist", C.c_double), ("sharpening", C.c_double), ("zero_parallax_dist", C.c_double),
                ("side_alpha_feather_size", C.c_int), ("std_alpha_feather_size", C.c_int),
                ("enable_top", C.c_int), ("enable_bottom", C.c_int),
                ("eqr_width", C.c_int), ("eqr_height", C.c_int), ("final_eqr_width", C.c_int),
                ("final_eqr_height", C.c_int), ("cubemap_width", C.c_int), ("cubemap_height", C.c_int),
                ("want_cubemap", C.c_int), ("cubemap_photo", C.c_int), ("threads", C.c_int),
                ("keep_intermediates", C.c_int), ("side_search20", C.c_int), ("polar_search20", C.c_int),
                ("enable_pole_removal", C.c_int)]


def make_params(**kw):
    p = ORenderParams(6.4, 0.0, 10000.0, 100, 31, 0, 0, 256, 128, 3480, 960, 1536, 1536, 0, 0, 1, 0, 0, 0, 0)
    for k, v in kw.items():
        if not hasattr(p, k):
            raise KeyError(k)
        setattr(p, k, v)
    return p


def lib():
    global _LIB
    if _LIB is None:
        so = os.path.join(ORACLE_DIR, "liboracle.so")
        if not os.path.exists(so):
            subprocess.check_call(["make", "-C", ORACLE_DIR, "-s"])
        L = C.CDLL(so)
        L.o_rig_from_json.restype = C.c_void_p
        L.o_render_frame.restype = C.c_void_p
        L.o_state_new.restype = C.c_void_p
        L.o_last_error.restype = C.c_char_p
        L.o_camera_id.restype = C.c_char_p
        L.o_camera_get_fov.restype = C.c_double
        L.o_camera_distort.restype = C.c_double
        L.o_camera_undistort.restype = C.c_double
        L.o_approximate_fov.restype = C.c_float
        L.o_ring_radius.restype = C.c_float
        L.o_usable_radius.restype = C.c_float
        _LIB = L
    return _LIB


def P(a):
    return a.ctypes.data_as(C.c_void_p)


def _c(a, dt):
    return np.ascontiguousarray(a, dtype=dt)


class Rig:
    def __init__(self, json_text):
        self.h = C.c_void_p(lib().o_rig_from_json(json_text.encode()))
        if not self.h:
            raise RuntimeError(lib().o_last_error().decode())

    @classmethod
    def load(cls, path):
        return cls(open(path).read())

    def __del__(self):
        if getattr(self, "h", None):
            lib().o_rig_free(self.h)

    n_cameras = property(lambda s: lib().o_rig_num_cameras(s.h))
    n_side = property(lambda s: lib().o_rig_num_side(s.h))

    def special(self, which):
        return lib().o_rig_special_camera(self.h, which)

    def side_index(self, i):
        return lib().o_rig_side_index(self.h, i)

    def camera_id(self, i):
        return lib().o_camera_id(self.h, i).decode()

    def pixel(self, cam, xyz):
        o = np.zeros(2); lib().o_camera_pixel(self.h, cam, P(_c(xyz, np.float64)), P(o)); return o

    def rig_ray(self, cam, px):
        o = np.zeros(6); lib().o_camera_rig_ray(self.h, cam, P(_c(px, np.float64)), P(o)); return o[:3], o[3:]

    def params(self, cam):
        o = np.zeros(22); lib().o_camera_params(self.h, cam, P(o)); return o

    def geometry(self, p):
        oi = np.zeros(4, np.int32); of = np.zeros(5, np.float32)
        lib().o_side_geometry(self.h, C.byref(p), P(oi), P(of))
        return dict(camW=int(oi[0]), camH=int(oi[1]), overlapW=int(oi[2]), numNovel=int(oi[3]),
                    hRad=of[0], vRad=of[1], fovHRad=of[2], slabDisp=of[3], shiftPx=of[4])

    def warp_map(self, cam, w, h, l, r, t, b):
        o = np.zeros((h, w, 2), np.float32)
        lib().o_spherical_warp_map(self.h, cam, w, h, C.c_float(l), C.c_float(r), C.c_float(t), C.c_float(b), P(o))
        return o

    def project_side(self, cam, src, dw, dh, l, r, t, b, feather=100):
        src = _c(src, np.uint8); o = np.zeros((dh, dw, 4), np.uint8)
        lib().o_project_side(self.h, cam, P(src), src.shape[1], src.shape[0], src.shape[2], dw, dh,
                             C.c_float(l), C.c_float(r), C.c_float(t), C.c_float(b), feather, P(o))
        return o

    def prepare_pole(self, p, src_bgr, top):
        src = _c(src_bgr, np.uint8); w = C.c_int(); h = C.c_int()
        lib().o_prepare_pole(self.h, C.byref(p), P(src), src.shape[1], src.shape[0], int(top), None, C.byref(w), C.byref(h))
        o = np.zeros((h.value, w.value, 4), np.uint8)
        lib().o_prepare_pole(self.h, C.byref(p), P(src), src.shape[1], src.shape[0], int(top), P(o), C.byref(w), C.byref(h))
        return o

    def pole_to_side(self, p, side, fish):
        side = _c(side, np.uint8); fish = _c(fish, np.uint8)
        ext_w = int(np.float32(fish.shape[1]) * np.float32(1.2))
        warped = np.zeros_like(side); flow = np.zeros((fish.shape[0], ext_w, 2), np.float32)
        lib().o_pole_to_side(self.h, C.byref(p), P(side), side.shape[1], side.shape[0], P(fish), fish.shape[1], fish.shape[0],
                             P(warped), P(flow))
        return warped, flow

    def render_frame(self, images, p, pole_masks=None, state=None):
        ims = [_c(i, np.uint8) for i in images]
        arr = (C.c_void_p * len(ims))(*[i.ctypes.data for i in ims])
        m1 = m2 = None
        if pole_masks is not None:
            m1, m2 = _c(pole_masks[0], np.uint8), _c(pole_masks[1], np.uint8)
        h = lib().o_render_frame(self.h, arr, ims[0].shape[1], ims[0].shape[0], C.byref(p), P(m1) if m1 is not None else None,
                                 P(m2) if m2 is not None else None, state.h if state is not None else None)
        if not h:
            raise RuntimeError(lib().o_last_error().decode())
        return FrameResult(C.c_void_p(h))


class FrameState:
    """Temporal state carried from one rendered frame to the next (video mode)."""

    def __init__(self):
        self.h = C.c_void_p(lib().o_state_new())

    def __del__(self):
        lib().o_state_free(self.h)


class FrameResult:
    def __init__(self, h):
        self.h = h

    def __del__(self):
        lib().o_result_free(self.h)

    def get(self, name, idx=0):
        ptr = C.c_void_p(); w = C.c_int(); h = C.c_int(); c = C.c_int(); f = C.c_int()
        if lib().o_result_get(self.h, name.encode(), idx, C.byref(ptr), C.byref(w), C.byref(h), C.byref(c), C.byref(f)):
            raise KeyError(name)
        n = w.value * h.value * c.value
        if n == 0:
            return None
        ct = C.c_float if f.value else C.c_uint8
        a = np.ctypeslib.as_array(C.cast(ptr, C.POINTER(ct)), shape=(n,)).copy()
        return a.reshape(h.value, w.value, c.value) if c.value > 1 else a.reshape(h.value, w.value)

    def times(self):
        t = np.zeros(6); lib().o_result_times(self.h, P(t))
        return dict(total=t[0], project=t[1], flow=t[2], novel=t[3], pole=t[4], sharpen=t[5])


def pixflow(i0, i1, alg="pixflow_low", hint=3, prev=None, dump_levels=False):
    """prev = (prevFlow, prevI0, prevI1) or None. hint: 0 unknown 1 right 2 down 3 left 4 up."""
    i0 = _c(i0, np.uint8); i1 = _c(i1, np.uint8); h, w = i0.shape[:2]
    out = np.zeros((h, w, 2), np.float32)
    if prev is not None:
        pf, p0, p1 = _c(prev[0], np.float32), _c(prev[1], np.uint8), _c(prev[2], np.uint8)
        a = (P(pf), P(p0), P(p1))
    else:
        a = (None, None, None)
    if lib().o_pixflow(P(i0), P(i1), w, h, a[0], a[1], a[2], int(alg == "pixflow_search_20"), hint, int(dump_levels), P(out)):
        raise RuntimeError(lib().o_last_error().decode())
    if not dump_levels:
        return out
    levels = []
    for i in range(lib().o_pixflow_num_levels()):
        lw = C.c_int(); lh = C.c_int(); lib().o_pixflow_level(i, C.byref(lw), C.byref(lh), None)
        a2 = np.zeros((lh.value, lw.value, 2), np.float32); lib().o_pixflow_level(i, C.byref(lw), C.byref(lh), P(a2)); levels.append(a2)
    return out, levels


def level_solve(I0, I1, a0, a1, flow=None, alg="pixflow_low", hint=3):
    I0, I1, a0, a1 = (_c(x, np.float32) for x in (I0, I1, a0, a1)); h, w = I0.shape
    f = np.zeros((h, w, 2), np.float32) if flow is None else _c(flow, np.float32).copy()
    lib().o_pixflow_level_solve(P(I0), P(I1), P(a0), P(a1), w, h, int(flow is not None), P(f), int(alg == "pixflow_search_20"), hint)
    return f


def novel_view_chunk(ovL, ovR, fLR, fRL, bufW, camW, numNovel, slabDisp):
    ovL, ovR = _c(ovL, np.uint8), _c(ovR, np.uint8); fLR, fRL = _c(fLR, np.float32), _c(fRL, np.float32)
    h, w = ovL.shape[:2]; oL = np.zeros((h, bufW, 4), np.uint8); oR = np.zeros_like(oL)
    lib().o_novel_view_chunk(P(ovL), P(ovR), P(fLR), P(fRL), w, h, bufW, camW, numNovel, C.c_float(slabDisp), P(oL), P(oR))
    return oL, oR


def flatten_deghost(base, top):
    base, top = _c(base, np.uint8), _c(top, np.uint8); o = np.zeros_like(base)
    lib().o_flatten_deghost(P(base), P(top), base.shape[1], base.shape[0], P(o)); return o


def feather_alpha(src, size):
    src = _c(src, np.uint8); o = np.zeros_like(src); lib().o_feather_alpha(P(src), src.shape[1], src.shape[0], size, P(o)); return o


def offset_wrap(src, offset):
    src = _c(src, np.uint8); o = np.zeros_like(src)
    lib().o_offset_wrap(P(src), src.shape[1], src.shape[0], src.shape[2], C.c_float(offset), P(o)); return o


def sharpen(img, sharpening):
    o = _c(img, np.uint8).copy(); lib().o_sharpen(P(o), o.shape[1], o.shape[0], C.c_double(sharpening)); return o


def cubemap(eye, fw, fh, photo=False):
    eye = _c(eye, np.uint8)
    o = np.zeros((6 * fh, fw, 3) if photo else (2 * fh, 3 * fw, 3), np.uint8)
    lib().o_cubemap(P(eye), eye.shape[1], eye.shape[0], fw, fh, int(photo), P(o)); return o


def resize_cubic_u8(src, dw, dh):
    src = _c(src, np.uint8); o = np.zeros((dh, dw, src.shape[2]), np.uint8)
    lib().o_resize_cubic_u8(P(src), src.shape[1], src.shape[0], src.shape[2], dw, dh, P(o)); return o
