"""ctypes binding of oracle/_ref/libref.so: the REFERENCE'S OWN sources compiled unchanged against the
third-party stand-ins of oracle/ref_shim/ (recipe: oracle/ref_build.sh, C API: oracle/ref_capi.cpp).
TEST INFRASTRUCTURE ONLY: imported by tests/ and by bench.py's reference arm / cpu_baseline leg."""
import ctypes as C
import os
import shutil
import subprocess
import tempfile

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_DIR = os.path.join(os.path.dirname(_HERE), "oracle")
REF_SO = os.path.join(ORACLE_DIR, "_ref", "libref.so")
REF_BIN = os.path.join(ORACLE_DIR, "_ref", "TestRenderStereoPanorama")
_LIB = None


def available():
    """True when oracle/_ref exists (built here from /root/reference; travels to the GPU box prebuilt)."""
    if not os.path.exists(REF_SO) and os.path.isdir("/root/reference"):
        subprocess.call(["bash", os.path.join(ORACLE_DIR, "ref_build.sh")])
    return os.path.exists(REF_SO)


class RefFrameParams(C.Structure):
    _fields_ = [("interpupilary_dist", C.c_double), ("sharpening", C.c_double), ("zero_parallax_dist", C.c_double),
                ("side_alpha_feather_size", C.c_int), ("std_alpha_feather_size", C.c_int), ("enable_top", C.c_int),
                ("enable_bottom", C.c_int), ("eqr_width", C.c_int), ("eqr_height", C.c_int), ("final_eqr_width", C.c_int),
                ("final_eqr_height", C.c_int), ("cubemap_width", C.c_int), ("cubemap_height", C.c_int), ("want_cubemap", C.c_int),
                ("cubemap_photo", C.c_int), ("enable_pole_removal", C.c_int), ("side_search20", C.c_int), ("polar_search20", C.c_int)]


def make_params(**kw):
    """Defaults of the reference's flags (TestRenderStereoPanorama.cpp:44-70)."""
    p = RefFrameParams(6.4, 0.0, 10000.0, 100, 31, 0, 0, 256, 128, 3480, 960, 1536, 1536, 0, 0, 0, 0, 0)
    for k, v in kw.items():
        if not hasattr(p, k):
            raise KeyError(k)
        setattr(p, k, v)
    return p


def lib():
    global _LIB
    if _LIB is None:
        if not available():
            raise RuntimeError("oracle/_ref/libref.so is missing (built by oracle/ref_build.sh where /root/reference exists)")
        L = C.CDLL(REF_SO)
        L.ref_last_error.restype = C.c_char_p
        _LIB = L
    return _LIB


def P(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def _c(a, dt):
    return np.ascontiguousarray(a, dtype=dt)


def _ck(rc):
    if rc:
        raise RuntimeError(lib().ref_last_error().decode())


def camera_unit_test():
    _ck(lib().ref_camera_unit_test())


def camera_pixel(rig_json, cam, xyz):
    o = np.zeros(2)
    _ck(lib().ref_camera_pixel(rig_json.encode(), cam, P(_c(xyz, np.float64)), P(o)))
    return o


def rig_summary(rig_json):
    n = C.c_int(); rr = C.c_float(); hf = C.c_float(); vf = C.c_float(); pf = C.c_double()
    t = C.create_string_buffer(256); b = C.create_string_buffer(256); b2 = C.create_string_buffer(256)
    _ck(lib().ref_rig_summary(rig_json.encode(), C.byref(n), C.byref(rr), C.byref(hf), C.byref(vf), C.byref(pf), t, b, b2))
    return dict(n_side=n.value, ring_radius=rr.value, h_fov=hf.value, v_fov=vf.value, pole_fov=pf.value,
                top=t.value.decode(), bottom=b.value.decode(), bottom2=b2.value.decode())


def pixflow(i0, i1, alg="pixflow_low", hint=3, prev=None):
    i0 = _c(i0, np.uint8); i1 = _c(i1, np.uint8); h, w = i0.shape[:2]
    out = np.zeros((h, w, 2), np.float32)
    if prev is not None:
        pf, p0, p1 = _c(prev[0], np.float32), _c(prev[1], np.uint8), _c(prev[2], np.uint8)
        a = (P(pf), P(p0), P(p1))
    else:
        a = (None, None, None)
    _ck(lib().ref_pixflow(P(i0), P(i1), w, h, a[0], a[1], a[2], alg.encode(), hint, P(out)))
    return out


def prepare_pair(l, r, alg="pixflow_low"):
    """NovelViewGeneratorAsymmetricFlow::prepare; returns (flowLtoR, flowRtoL, seconds) -- TestOpticalFlow's RUNTIME."""
    l = _c(l, np.uint8); r = _c(r, np.uint8); h, w = l.shape[:2]
    a = np.zeros((h, w, 2), np.float32); b = np.zeros_like(a); s = C.c_double()
    _ck(lib().ref_prepare_pair(P(l), P(r), w, h, alg.encode(), P(a), P(b), C.byref(s)))
    return a, b, s.value


def novel_view_chunk(ovL, ovR, fLR, fRL, eqr_width, num_cams, camW, numNovel, fovHRad, slabDisp):
    ovL, ovR = _c(ovL, np.uint8), _c(ovR, np.uint8); fLR, fRL = _c(fLR, np.float32), _c(fRL, np.float32)
    h, w = ovL.shape[:2]; bw = eqr_width // num_cams
    oL = np.zeros((h, bw, 4), np.uint8); oR = np.zeros_like(oL)
    _ck(lib().ref_novel_view_chunk(P(ovL), P(ovR), P(fLR), P(fRL), w, h, eqr_width, num_cams, camW, numNovel, C.c_float(fovHRad),
                                   C.c_float(slabDisp), P(oL), P(oR)))
    return oL, oR


def flatten_deghost(base, top):
    base, top = _c(base, np.uint8), _c(top, np.uint8); o = np.zeros_like(base)
    _ck(lib().ref_flatten_deghost(P(base), P(top), base.shape[1], base.shape[0], P(o))); return o


def feather_alpha(src, size):
    src = _c(src, np.uint8); o = np.zeros_like(src)
    _ck(lib().ref_feather_alpha(P(src), src.shape[1], src.shape[0], size, P(o))); return o


def offset_wrap(src, offset):
    src = _c(src, np.uint8); o = np.zeros_like(src)
    _ck(lib().ref_offset_wrap(P(src), src.shape[1], src.shape[0], src.shape[2], C.c_float(offset), P(o))); return o


def sharpen(img, sharpening):
    o = _c(img, np.uint8).copy()
    _ck(lib().ref_sharpen(P(o), o.shape[1], o.shape[0], C.c_double(sharpening))); return o


def cubemap(eye, fw, fh, photo=False):
    eye = _c(eye, np.uint8)
    o = np.zeros((6 * fh, fw, 3) if photo else (2 * fh, 3 * fw, 3), np.uint8)
    _ck(lib().ref_cubemap(P(eye), eye.shape[1], eye.shape[0], fw, fh, int(photo), P(o))); return o


def project_side(rig_json, side_idx, src, dw, dh, l, r, t, b, feather=100):
    src = _c(src, np.uint8); o = np.zeros((dh, dw, 4), np.uint8)
    _ck(lib().ref_project_side(rig_json.encode(), side_idx, P(src), src.shape[1], src.shape[0], src.shape[2], dw, dh,
                               C.c_float(l), C.c_float(r), C.c_float(t), C.c_float(b), feather, P(o)))
    return o


class Frame:
    """One frame pushed through the reference's renderStereoPanorama() (TestRenderStereoPanorama.cpp:716-972)."""

    def __init__(self, rig_json, images, params, pole_masks=None, work_dir=None, frame="000000", prev_frame=None):
        self.own = work_dir is None
        self.work = work_dir or tempfile.mkdtemp(prefix="s360_ref_", dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
        ims = [_c(i, np.uint8) for i in images]
        arr = (C.c_void_p * len(ims))(*[i.ctypes.data for i in ims])
        m1 = m2 = None
        if pole_masks is not None:
            m1, m2 = _c(pole_masks[0], np.uint8), _c(pole_masks[1], np.uint8)
        ew, eh = ((params.final_eqr_width, params.final_eqr_height) if self._resizes(params) else (params.eqr_width, 2 * params.eqr_height))
        self.equirect = np.zeros((eh, ew, 3), np.uint8)
        self.cubemap = None
        cw = C.c_int(); ch = C.c_int(); w = C.c_int(); h = C.c_int(); secs = C.c_double()
        cube = None
        if params.want_cubemap:
            shape = (2 * 6 * params.cubemap_height, params.cubemap_width, 3) if params.cubemap_photo else (2 * 2 * params.cubemap_height, 3 * params.cubemap_width, 3)
            cube = np.zeros(shape, np.uint8)
        _ck(lib().ref_render_frame(rig_json.encode(), arr, len(ims), ims[0].shape[1], ims[0].shape[0], C.byref(params), self.work.encode(),
                                   (prev_frame or "").encode(), frame.encode(), P(m1), P(m2), P(self.equirect), C.byref(w), C.byref(h),
                                   P(cube), C.byref(cw), C.byref(ch), C.byref(secs)))
        assert (w.value, h.value) == (ew, eh), (w.value, h.value, ew, eh)
        self.cubemap = cube
        self.seconds = secs.value
        self.frame = frame

    @staticmethod
    def _resizes(p):   # the guard of TestRenderStereoPanorama.cpp:938-941
        return (p.final_eqr_width != 0 and p.final_eqr_height != 0 and p.final_eqr_width != p.eqr_width
                and p.final_eqr_height != p.eqr_height // 2)

    def image(self, rel):
        w = C.c_int(); h = C.c_int(); c = C.c_int()
        _ck(lib().ref_frame_image(self.work.encode(), rel.encode(), None, C.byref(w), C.byref(h), C.byref(c)))
        o = np.zeros((h.value, w.value, c.value), np.uint8)
        _ck(lib().ref_frame_image(self.work.encode(), rel.encode(), P(o), C.byref(w), C.byref(h), C.byref(c)))
        return o

    def flow(self, name):
        rel = f"flow/{self.frame}/{name}.bin"
        w = C.c_int(); h = C.c_int()
        _ck(lib().ref_frame_flow(self.work.encode(), rel.encode(), None, C.byref(w), C.byref(h)))
        o = np.zeros((h.value, w.value, 2), np.float32)
        _ck(lib().ref_frame_flow(self.work.encode(), rel.encode(), P(o), C.byref(w), C.byref(h)))
        return o

    def close(self):
        if self.own and self.work and os.path.isdir(self.work):
            shutil.rmtree(self.work, ignore_errors=True)
            self.work = None

    def __del__(self):
        self.close()


# ---- TestHyperPreview of the reference (oracle/_ref/libref_preview.so, oracle/ref_preview_capi.cpp) ----
REF_PREVIEW_SO = os.path.join(ORACLE_DIR, "_ref", "libref_preview.so")
_PLIB = None


def hyper_preview(rig_json, raw_top, raw_bottom, raw_bottom2, eqr_w, eqr_h, gamma=1.0, softmax_coef=30.0):
    """PreviewRenderer::render of the compiled reference: three uint16 Bayer frames -> float32 (eqr_h, eqr_w, 3)."""
    global _PLIB
    if _PLIB is None:
        available()
        _PLIB = C.CDLL(REF_PREVIEW_SO)
        _PLIB.refp_last_error.restype = C.c_char_p
    r = [_c(a, np.uint16) for a in (raw_top, raw_bottom, raw_bottom2)]
    h, w = r[0].shape
    out = np.zeros((eqr_h, eqr_w, 3), np.float32)
    if _PLIB.refp_render(rig_json.encode(), P(r[0]), P(r[1]), P(r[2]), w, h, eqr_w, eqr_h, C.c_double(gamma), C.c_double(softmax_coef), P(out)):
        raise RuntimeError(_PLIB.refp_last_error().decode())
    return out


# ---- the reference's soft ISP (camera_isp/CameraIsp.h; oracle/_ref/libref_isp.so, oracle/ref_isp_capi.cpp) ----
REF_ISP_SO = os.path.join(ORACLE_DIR, "_ref", "libref_isp.so")
_ILIB = None


def _isp_lib():
    global _ILIB
    if _ILIB is None:
        available()
        if not os.path.exists(REF_ISP_SO):
            raise RuntimeError("oracle/_ref/libref_isp.so is missing (oracle/ref_build.sh)")
        _ILIB = C.CDLL(REF_ISP_SO)
        _ILIB.refisp_last_error.restype = C.c_char_p
    return _ILIB


def isp_process(json_text, raw, demosaic_filter=2, resize=1, output_bpp=8, tone_curve=True, black_level_offset=0, swizzle=True):
    """Raw2Rgb's CPU branch (Raw2Rgb.cpp:445-456): raw Bayer uint8/uint16 (h, w) -> (h/resize, w/resize, 3) uint8/uint16."""
    L = _isp_lib()
    raw = np.ascontiguousarray(raw)
    assert raw.dtype in (np.uint8, np.uint16) and raw.ndim == 2
    h, w = raw.shape
    out = np.zeros((h // resize, w // resize, 3), np.uint8 if output_bpp == 8 else np.uint16)
    if L.refisp_process(json_text.encode(), P(raw), w, h, 8 if raw.dtype == np.uint8 else 16, demosaic_filter, resize, output_bpp, int(tone_curve),
                        black_level_offset, int(swizzle), P(out)):
        raise RuntimeError(L.refisp_last_error().decode())
    return out


def isp_tables(json_text, w, h, output_bpp=8, tone_curve=True):
    """Tables CameraIsp::setup() derives: compositeCCM (3, 3), toneCurveLut (4096, 3), vignette gains per column (w, 3) and row (h, 3)."""
    L = _isp_lib()
    ccm = np.zeros((3, 3), np.float32); lut = np.zeros((4096, 3), np.float32)
    vh = np.zeros((w, 3), np.float32); vv = np.zeros((h, 3), np.float32)
    if L.refisp_tables(json_text.encode(), output_bpp, int(tone_curve), w, h, P(ccm), P(lut), P(vh), P(vv)):
        raise RuntimeError(L.refisp_last_error().decode())
    return ccm, lut, vh, vv


REF_ISP_B200_SO = os.path.join(ORACLE_DIR, "_ref", "libref_isp_b200.so")
_IBLIB = None


def isp_process_b200(json_text, raw, demosaic_filter=2, resize=1, output_bpp=8, tone_curve=True, black_level_offset=0, swizzle=True):
    """The same call sequence inside the reference build, but on the adapter of integration/reference_tree/camera_isp/B200CameraIsp.h
    (the reference's CameraIsp interface, pixels on the GPU through include/s360.h)."""
    global _IBLIB
    if _IBLIB is None:
        _IBLIB = C.CDLL(REF_ISP_B200_SO)
        _IBLIB.refisp_last_error.restype = C.c_char_p
    raw = np.ascontiguousarray(raw)
    h, w = raw.shape
    out = np.zeros((h // resize, w // resize, 3), np.uint8 if output_bpp == 8 else np.uint16)
    if _IBLIB.refisp_process_b200(json_text.encode(), P(raw), w, h, 8 if raw.dtype == np.uint8 else 16, demosaic_filter, resize, output_bpp,
                                  int(tone_curve), black_level_offset, int(swizzle), P(out)):
        raise RuntimeError(_IBLIB.refisp_last_error().decode())
    return out


# ---- the reference's footage container reader and raw-sample converter (oracle/_ref/libref_unpack.so, oracle/ref_unpack_capi.cpp) ----
REF_UNPACK_SO = os.path.join(ORACLE_DIR, "_ref", "libref_unpack.so")
_ULIB = None


def unpack_available():
    return os.path.exists(REF_UNPACK_SO)


def _unpack_lib():
    global _ULIB
    if _ULIB is None:
        if not os.path.exists(REF_UNPACK_SO):
            raise RuntimeError("oracle/_ref/libref_unpack.so is missing (oracle/ref_build.sh)")
        _ULIB = C.CDLL(REF_UNPACK_SO)
        _ULIB.refunpack_last_error.restype = C.c_char_p
        _ULIB.refunpack_open.restype = C.c_void_p
    return _ULIB


def raw_convert(packed, bits, w, h):
    """RawConverter::convert8Frame / convert12Frame (RawConverter.cpp:14-59): packed bytes -> (h, w) uint16."""
    L = _unpack_lib()
    packed = np.ascontiguousarray(packed, np.uint8)
    out = np.zeros((h, w), np.uint16)
    if L.refunpack_convert(P(packed), bits, w, h, P(out)):
        raise RuntimeError(L.refunpack_last_error().decode())
    return out


class FootageFile:
    """BinaryFootageFile of the reference build."""

    def __init__(self, path):
        L = _unpack_lib()
        meta = (C.c_uint32 * 8)()
        frames, fsize = C.c_uint64(), C.c_uint64()
        self.h = L.refunpack_open(str(path).encode(), meta, C.byref(frames), C.byref(fsize))
        if not self.h:
            raise RuntimeError(L.refunpack_last_error().decode())
        self.meta = list(meta)
        self.frames, self.frame_size = frames.value, fsize.value

    def frame(self, f, c):
        L = _unpack_lib()
        out = np.zeros(self.frame_size, np.uint8)
        if L.refunpack_frame(C.c_void_p(self.h), f, c, P(out)):
            raise RuntimeError(L.refunpack_last_error().decode())
        return out

    def close(self):
        if self.h:
            _unpack_lib().refunpack_close(C.c_void_p(self.h))
            self.h = None
