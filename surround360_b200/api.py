"""Python host-side mirror of the reference's operator interface for the render path, over
the C ABI in include/s360.h (libs360b200.so, hand-written sm_100a kernels).

Names follow the reference: computeOpticalFlow / DirectionHint
(source/optical_flow/OpticalFlowInterface.h:20-42), makeOpticalFlowByName's algorithm names
(OpticalFlowFactory.h:23-64), bicubicRemapToSpherical (render/ImageWarper.h:59-66),
flattenLayersDeghostPreferBase / featherAlphaChannel (util/CvUtil.h), sharpen
(util/Filter.h), renderStereoPanorama (test/TestRenderStereoPanorama.cpp:716-972).

There is no CPU fallback: importing works without a GPU (so the CPU test tier can check the
exported symbols), but every call needs a B200 and fails loudly otherwise.
"""
import ctypes as C
import os
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libs360b200.so")


class VrCamException(RuntimeError):
    """Mirror of surround360::VrCamException (util/VrCamException.h:18)."""


class S360Params(C.Structure):
    _fields_ = [("interpupilary_dist", C.c_double), ("sharpening", C.c_double), ("zero_parallax_dist", C.c_double),
                ("side_alpha_feather_size", C.c_int), ("std_alpha_feather_size", C.c_int),
                ("enable_top", C.c_int), ("enable_bottom", C.c_int),
                ("eqr_width", C.c_int), ("eqr_height", C.c_int), ("final_eqr_width", C.c_int),
                ("final_eqr_height", C.c_int), ("cubemap_width", C.c_int), ("cubemap_height", C.c_int),
                ("want_cubemap", C.c_int), ("cubemap_photo", C.c_int),
                ("side_flow_alg", C.c_char * 32), ("polar_flow_alg", C.c_char * 32),
                ("enable_pole_removal", C.c_int), ("reserved0", C.c_int), ("poleremoval_flow_alg", C.c_char * 32)]


class S360Geometry(C.Structure):
    _fields_ = [("cam_w", C.c_int), ("cam_h", C.c_int), ("overlap_w", C.c_int), ("num_novel", C.c_int),
                ("h_radians", C.c_float), ("v_radians", C.c_float), ("fov_h_radians", C.c_float),
                ("slab_disp", C.c_float), ("shift_px", C.c_float), ("pole_h", C.c_int)]


class DirectionHint:  # OpticalFlowInterface.h:24
    UNKNOWN, RIGHT, DOWN, LEFT, UP = range(5)


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(f"{LIB_PATH} is missing: build it with `python __graft_entry__.py` "
                              "(nvcc, sm_100a). There is no CPU fallback.")
        L = C.CDLL(LIB_PATH)
        L.s360_last_error.restype = C.c_char_p
        L.s360_version.restype = C.c_char_p
        L.s360_rig_camera_id.restype = C.c_char_p
        L.s360_ctx_stream.restype = C.c_void_p
        L.s360_ctx_launch_count.restype = C.c_ulonglong
        _lib = L
    return _lib


def _check(rc):
    if rc != 0:
        raise VrCamException(lib().s360_last_error().decode())


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _c(a, dt):
    return np.ascontiguousarray(a, dtype=dt)


def default_params(**kw):
    p = S360Params()
    lib().s360_params_default(C.byref(p))
    for k, v in kw.items():
        if not hasattr(p, k):
            raise KeyError(k)
        setattr(p, k, v.encode() if isinstance(v, str) else v)
    return p


# quality presets of scripts/batch_process_video.py:176-199
PRESETS = {
    "3k": dict(eqr_width=3080, eqr_height=1540, final_eqr_width=3080, final_eqr_height=3080),
    "4k": dict(eqr_width=4200, eqr_height=1024, final_eqr_width=4096, final_eqr_height=2048),
    "6k": dict(eqr_width=6300, eqr_height=3072, final_eqr_width=6144, final_eqr_height=6144),
    "8k": dict(eqr_width=8400, eqr_height=4096, final_eqr_width=8192, final_eqr_height=8192),
}


class Context:
    """One render context = one CUDA device + stream (s360_ctx)."""

    def __init__(self, device=0):
        self.h = C.c_void_p()
        _check(lib().s360_ctx_create(int(device), C.byref(self.h)))

    def close(self):
        if getattr(self, "h", None):
            lib().s360_ctx_destroy(self.h)
            self.h = None

    __del__ = close

    # ---- rig ----
    def load_rig(self, path):
        _check(lib().s360_rig_load_file(self.h, path.encode()))

    def load_rig_json(self, text):
        _check(lib().s360_rig_load_json(self.h, text.encode()))

    def rig_blob(self):
        n = C.c_size_t()
        _check(lib().s360_rig_serialize(self.h, None, 0, C.byref(n)))
        buf = np.zeros(n.value, np.uint8)
        _check(lib().s360_rig_serialize(self.h, _p(buf), buf.size, C.byref(n)))
        return buf

    def load_rig_blob(self, buf):
        buf = _c(buf, np.uint8)
        _check(lib().s360_rig_deserialize(self.h, _p(buf), buf.size))

    n_cameras = property(lambda s: lib().s360_rig_num_cameras(s.h))
    n_side = property(lambda s: lib().s360_rig_num_side(s.h))

    def special_camera(self, which):
        return lib().s360_rig_special_camera(self.h, which)

    def side_index(self, i):
        return lib().s360_rig_side_index(self.h, i)

    def camera_id(self, cam):
        return lib().s360_rig_camera_id(self.h, cam).decode()

    def camera_pixel(self, cam, xyz):
        o = np.zeros(2)
        _check(lib().s360_camera_pixel(self.h, cam, _p(_c(xyz, np.float64)), _p(o)))
        return o

    def geometry(self, params):
        g = S360Geometry()
        _check(lib().s360_geometry_query(self.h, C.byref(params), C.byref(g)))
        return g

    # ---- operators ----
    def bicubicRemapToSpherical(self, cam, src, dst_w, dst_h, left, right, top, bottom, feather=0):
        src = _c(src, np.uint8)
        out = np.zeros((dst_h, dst_w, 4), np.uint8)
        _check(lib().s360_project_spherical(self.h, cam, _p(src), src.shape[1], src.shape[0], src.shape[2], C.c_float(left),
                                            C.c_float(right), C.c_float(top), C.c_float(bottom), feather, _p(out), dst_w, dst_h))
        return out

    def computeOpticalFlow(self, alg_name, I0BGRA, I1BGRA, prevFlow=None, prevI0BGRA=None, prevI1BGRA=None,
                           hint=DirectionHint.UNKNOWN):
        i0, i1 = _c(I0BGRA, np.uint8), _c(I1BGRA, np.uint8)
        h, w = i0.shape[:2]
        flow = np.zeros((h, w, 2), np.float32)
        pf = p0 = p1 = None
        if prevFlow is not None:
            pf, p0, p1 = _c(prevFlow, np.float32), _c(prevI0BGRA, np.uint8), _c(prevI1BGRA, np.uint8)
        _check(lib().s360_flow_compute(self.h, alg_name.encode(), _p(i0), _p(i1), w, h, _p(pf) if pf is not None else None,
                                       _p(p0) if p0 is not None else None, _p(p1) if p1 is not None else None, _p(flow), int(hint)))
        return flow

    def computeOpticalFlowStrided(self, alg_name, I0view, I1view, prevFlow=None, prevI1BGRA=None, hint=DirectionHint.UNKNOWN):
        """computeOpticalFlow on non-contiguous views (numpy slices of wider images, like the reference's ROI crops, TRSP:196-198):
        the row strides go through the ABI, nothing is copied on the host."""
        for a in (I0view, I1view):
            assert a.dtype == np.uint8 and a.ndim == 3 and a.shape[2] == 4 and a.strides[1] == 4 and a.strides[2] == 1
        h, w = I0view.shape[:2]
        flow = np.zeros((h, w, 2), np.float32)
        pf = p1 = None
        if prevFlow is not None:
            pf, p1 = prevFlow, prevI1BGRA
        _check(lib().s360_flow_compute_strided(self.h, alg_name.encode(), C.c_void_p(I0view.ctypes.data), C.c_size_t(I0view.strides[0]),
                                               C.c_void_p(I1view.ctypes.data), C.c_size_t(I1view.strides[0]), w, h,
                                               C.c_void_p(pf.ctypes.data) if pf is not None else None, C.c_size_t(pf.strides[0] if pf is not None else 0),
                                               C.c_void_p(p1.ctypes.data) if p1 is not None else None, C.c_size_t(p1.strides[0] if p1 is not None else 0),
                                               _p(flow), C.c_size_t(0), int(hint)))
        return flow

    def flow_compute_dev(self, alg_name, i0_ptr, i0_stride, i1_ptr, i1_stride, w, h, flow_ptr, hint=DirectionHint.UNKNOWN):
        """Device pointers (ints), strides in bytes; enqueues on the context stream, no synchronisation."""
        _check(lib().s360_flow_compute_dev(self.h, alg_name.encode(), C.c_void_p(int(i0_ptr)), C.c_size_t(i0_stride), C.c_void_p(int(i1_ptr)),
                                           C.c_size_t(i1_stride), w, h, C.c_void_p(int(flow_ptr)), int(hint)))

    def novel_view_chunk_dev(self, ovl_ptr, ovr_ptr, ov_stride, flr_ptr, frl_ptr, w, h, buf_w, cam_w, num_novel, slab_disp, out_l_ptr, out_r_ptr):
        _check(lib().s360_novel_view_chunk_dev(self.h, C.c_void_p(int(ovl_ptr)), C.c_void_p(int(ovr_ptr)), C.c_size_t(ov_stride), C.c_void_p(int(flr_ptr)),
                                               C.c_void_p(int(frl_ptr)), w, h, buf_w, cam_w, num_novel, C.c_float(slab_disp),
                                               C.c_void_p(int(out_l_ptr)), C.c_void_p(int(out_r_ptr))))

    def computeOpticalFlowBatch(self, alg_name, I0s, I1s, hints):
        i0 = [_c(a, np.uint8) for a in I0s]
        i1 = [_c(a, np.uint8) for a in I1s]
        n = len(i0)
        h, w = i0[0].shape[:2]
        flows = [np.zeros((h, w, 2), np.float32) for _ in range(n)]
        a0 = (C.c_void_p * n)(*[a.ctypes.data for a in i0])
        a1 = (C.c_void_p * n)(*[a.ctypes.data for a in i1])
        af = (C.c_void_p * n)(*[a.ctypes.data for a in flows])
        hs = (C.c_int * n)(*[int(x) for x in hints])
        _check(lib().s360_flow_compute_batch(self.h, alg_name.encode(), n, a0, a1, w, h, af, hs))
        return flows

    def combineLazyNovelViews(self, overlapL, overlapR, flowLtoR, flowRtoL, buf_w, cam_w, num_novel, slab_disp):
        l, r = _c(overlapL, np.uint8), _c(overlapR, np.uint8)
        flr, frl = _c(flowLtoR, np.float32), _c(flowRtoL, np.float32)
        h, w = l.shape[:2]
        oL = np.zeros((h, buf_w, 4), np.uint8)
        oR = np.zeros_like(oL)
        _check(lib().s360_novel_view_chunk(self.h, _p(l), _p(r), _p(flr), _p(frl), w, h, buf_w, cam_w, num_novel,
                                           C.c_float(slab_disp), _p(oL), _p(oR)))
        return oL, oR

    def flattenLayersDeghostPreferBase(self, bottomLayer, topLayer):
        b, t = _c(bottomLayer, np.uint8), _c(topLayer, np.uint8)
        o = np.zeros_like(b)
        _check(lib().s360_flatten_deghost(self.h, _p(b), _p(t), b.shape[1], b.shape[0], _p(o)))
        return o

    def featherAlphaChannel(self, src, erode_size):
        s = _c(src, np.uint8)
        o = np.zeros_like(s)
        _check(lib().s360_feather_alpha(self.h, _p(s), s.shape[1], s.shape[0], erode_size, _p(o)))
        return o

    def offsetHorizontalWrap(self, src, offset):
        s = _c(src, np.uint8)
        o = np.zeros_like(s)
        _check(lib().s360_offset_wrap(self.h, _p(s), s.shape[1], s.shape[0], s.shape[2], C.c_float(offset), _p(o)))
        return o

    def sharpen(self, bgr, sharpening):
        o = _c(bgr, np.uint8).copy()
        _check(lib().s360_sharpen(self.h, _p(o), o.shape[1], o.shape[0], C.c_double(sharpening)))
        return o

    def cubemap(self, eye_bgr, face_w, face_h, photo=False):
        e = _c(eye_bgr, np.uint8)
        o = np.zeros((6 * face_h, face_w, 3) if photo else (2 * face_h, 3 * face_w, 3), np.uint8)
        _check(lib().s360_cubemap(self.h, _p(e), e.shape[1], e.shape[0], face_w, face_h, int(photo), _p(o)))
        return o

    def resizeCubic(self, src, dst_w, dst_h):
        s = _c(src, np.uint8)
        o = np.zeros((dst_h, dst_w, s.shape[2]), np.uint8)
        _check(lib().s360_resize_cubic(self.h, _p(s), s.shape[1], s.shape[0], s.shape[2], dst_w, dst_h, _p(o)))
        return o

    def poleToSideFlow(self, params, side_bgra, fisheye_bgra):
        s, f = _c(side_bgra, np.uint8), _c(fisheye_bgra, np.uint8)
        ext_w = int(np.float32(f.shape[1]) * np.float32(1.2))
        warped = np.zeros_like(s)
        flow = np.zeros((f.shape[0], ext_w, 2), np.float32)
        _check(lib().s360_pole_to_side(self.h, C.byref(params), _p(s), s.shape[1], s.shape[0], _p(f), f.shape[1], f.shape[0],
                                       _p(warped), _p(flow)))
        return warped, flow

    def hyperPreview(self, raw_top, raw_bottom, raw_bottom2, eqr_w, eqr_h, gamma=1.0, softmax_coef=30.0):
        """PreviewRenderer::render (TestHyperPreview.cpp:141-159): three uint16 Bayer frames -> float32 equirect (eqr_h, eqr_w, 3)."""
        r = [_c(a, np.uint16) for a in (raw_top, raw_bottom, raw_bottom2)]
        h, w = r[0].shape
        out = np.zeros((eqr_h, eqr_w, 3), np.float32)
        _check(lib().s360_hyper_preview(self.h, _p(r[0]), _p(r[1]), _p(r[2]), w, h, eqr_w, eqr_h, C.c_double(gamma), C.c_double(softmax_coef), _p(out)))
        return out

    # ---- whole frame ----
    def frame_to_nv12(self, bgr, matrix=0):
        """Encode hand-off: (h, w, 3) BGR8 -> (y (h, w), uv (ceil(h / 2), 2 * ceil(w / 2))) NV12 planes."""
        b = _c(bgr, np.uint8)
        h, w = b.shape[:2]
        y = np.zeros((h, w), np.uint8); uv = np.zeros(((h + 1) // 2, 2 * ((w + 1) // 2)), np.uint8)
        _check(lib().s360_frame_to_nv12(self.h, _p(b), w, h, matrix, _p(y), _p(uv)))
        return y, uv

    def frame_to_nv12_dev(self, bgr_ptr, w, h, y_ptr, uv_ptr, matrix=0, bgr_stride=0, y_pitch=0, uv_pitch=0):
        _check(lib().s360_frame_to_nv12_dev(self.h, C.c_void_p(bgr_ptr), C.c_size_t(bgr_stride), w, h, matrix, C.c_void_p(y_ptr), C.c_size_t(y_pitch),
                                            C.c_void_p(uv_ptr), C.c_size_t(uv_pitch)))

    def output_size(self, params):
        a, b, c, d = C.c_int(), C.c_int(), C.c_int(), C.c_int()
        _check(lib().s360_frame_output_size(self.h, C.byref(params), C.byref(a), C.byref(b), C.byref(c), C.byref(d)))
        return a.value, b.value, c.value, d.value

    def set_pole_masks(self, mask_bottom_bgr, mask_bottom2_bgr):
        m1, m2 = _c(mask_bottom_bgr, np.uint8), _c(mask_bottom2_bgr, np.uint8)
        _check(lib().s360_set_pole_masks(self.h, _p(m1), _p(m2), m1.shape[1], m1.shape[0]))

    def keep_intermediates(self, on=True):
        _check(lib().s360_ctx_set_keep_intermediates(self.h, int(on)))

    def renderStereoPanorama(self, params, images_bgr, equirect_out=None, cubemap_out=None):
        """images_bgr: list (rig order) of HxWx3 uint8 host arrays (pinned or pageable)."""
        ims = [_c(i, np.uint8) for i in images_bgr]
        eq_w, eq_h, cw, ch = self.output_size(params)
        eq = equirect_out if equirect_out is not None else np.zeros((eq_h, eq_w, 3), np.uint8)
        cube = None
        if cw > 0:
            cube = cubemap_out if cubemap_out is not None else np.zeros((ch, cw, 3), np.uint8)
        arr = (C.c_void_p * len(ims))(*[i.ctypes.data for i in ims])
        _check(lib().s360_render_frame(self.h, C.byref(params), arr, ims[0].shape[1], ims[0].shape[0], _p(eq),
                                       _p(cube) if cube is not None else None))
        return (eq, cube) if cube is not None else eq

    def render_frame_dev(self, params, image_ptrs, img_w, img_h, equirect_ptr, cubemap_ptr=None):
        """Device-pointer variant; enqueues on the context stream without synchronising."""
        arr = (C.c_void_p * len(image_ptrs))(*[int(p) for p in image_ptrs])
        _check(lib().s360_render_frame_dev(self.h, C.byref(params), arr, img_w, img_h, C.c_void_p(int(equirect_ptr)),
                                           C.c_void_p(int(cubemap_ptr)) if cubemap_ptr else None))

    def synchronize(self):
        _check(lib().s360_ctx_synchronize(self.h))

    def stream(self):
        return lib().s360_ctx_stream(self.h)

    def launch_count(self):
        return int(lib().s360_ctx_launch_count(self.h))

    def stage_ms(self):
        o = (C.c_float * 7)()
        _check(lib().s360_frame_stage_ms(self.h, o))
        names = ["total", "projection", "side_flow", "novel_view", "poles", "sharpen", "output"]
        return dict(zip(names, list(o)))

    def set_temporal(self, on=True):
        _check(lib().s360_ctx_set_temporal(self.h, int(on)))

    def reset_temporal(self):
        _check(lib().s360_ctx_reset_temporal(self.h))

    # video mode through files (TRSP:200-255, 413-452): previous-frame state from host buffers
    def frame_prepare(self, params, img_w, img_h):
        _check(lib().s360_frame_prepare(self.h, C.byref(params), int(img_w), int(img_h)))

    def temporal_sizes(self):
        v = [C.c_int() for _ in range(4)]
        _check(lib().s360_temporal_sizes(self.h, *[C.byref(x) for x in v]))
        return tuple(x.value for x in v)   # overlap_w, overlap_h, ext_w, ext_h

    def temporal_import_side(self, pair, flowLtoR, flowRtoL, overlapL, overlapR):
        a, b = _c(flowLtoR, np.float32), _c(flowRtoL, np.float32)
        l, r = _c(overlapL, np.uint8), _c(overlapR, np.uint8)
        _check(lib().s360_temporal_import_side(self.h, int(pair), _p(a), _p(b), _p(l), _p(r)))

    def temporal_import_pole(self, slot, flow, extFisheye):
        a, x = _c(flow, np.float32), _c(extFisheye, np.uint8)
        _check(lib().s360_temporal_import_pole(self.h, int(slot), _p(a), _p(x)))

    def temporal_commit(self):
        _check(lib().s360_temporal_commit(self.h))

    def set_graphs(self, on=True):
        _check(lib().s360_ctx_set_graphs(self.h, int(on)))

    def set_latency_mode(self, on=True):
        """s360_ctx_set_latency_mode: the sweep variant that shortens one frame rendered alone (bit-identical results)."""
        _check(lib().s360_ctx_set_latency_mode(self.h, int(on)))

    def profile_enable(self, on=True):
        _check(lib().s360_profile_enable(self.h, int(on)))

    def profile_sweep(self):
        ms, n, px, steps = C.c_double(), C.c_ulonglong(), C.c_double(), C.c_double()
        _check(lib().s360_profile_sweep(self.h, C.byref(ms), C.byref(n), C.byref(px), C.byref(steps)))
        fb = C.c_ulonglong()
        _check(lib().s360_profile_sweep_fallbacks(self.h, C.byref(fb)))
        return dict(ms=ms.value, launches=n.value, level_pixels=px.value, steps=steps.value, fallbacks=fb.value)

    def profile_kernels(self):
        """Per-kernel {kernel, ms, algorithmic_bytes, launches} since profile_enable(True)."""
        import json
        need = C.c_size_t()
        _check(lib().s360_profile_kernels(self.h, None, 0, C.byref(need)))
        buf = C.create_string_buffer(need.value)
        _check(lib().s360_profile_kernels(self.h, buf, need.value, C.byref(need)))
        return json.loads(buf.value.decode())

    def intermediate(self, name, idx=0):
        w, h, c, f = C.c_int(), C.c_int(), C.c_int(), C.c_int()
        _check(lib().s360_frame_intermediate(self.h, name.encode(), idx, None, C.byref(w), C.byref(h), C.byref(c), C.byref(f)))
        out = np.zeros((h.value, w.value, c.value), np.float32 if f.value else np.uint8)
        _check(lib().s360_frame_intermediate(self.h, name.encode(), idx, _p(out), C.byref(w), C.byref(h), C.byref(c), C.byref(f)))
        return out


class CameraIsp:
    """The reference's soft ISP object (camera_isp/CameraIsp.h:44-1301) on the GPU: same constructor arguments, same setters,
    loadImage + getImage in one call. Driven like Raw2Rgb's CPU branch (Raw2Rgb.cpp:445-456):
        isp = CameraIsp(json_text, output_bpp); isp.setBitsPerPixel(16); isp.setDemosaicFilter(2); isp.setResize(1)
        isp.enableToneMap(); isp.addBlackLevelOffset(0); bgr = isp.getImage(ctx, raw)"""
    BILINEAR_DM_FILTER, FREQUENCY_DM_FILTER, EDGE_AWARE_DM_FILTER = 0, 1, 2   # CameraIsp.h:35-40

    def __init__(self, json_input, output_bpp=8):
        self.h = C.c_void_p()
        self.output_bpp = output_bpp
        self._resize = 1
        _check(lib().s360_isp_create(json_input.encode(), output_bpp, C.byref(self.h)))

    def close(self):
        if self.h:
            lib().s360_isp_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def setBitsPerPixel(self, bits):
        _check(lib().s360_isp_set_bits_per_pixel(self.h, bits))

    def setDemosaicFilter(self, f):
        _check(lib().s360_isp_set_demosaic_filter(self.h, f))

    def setResize(self, r):
        _check(lib().s360_isp_set_resize(self.h, r))
        self._resize = r

    def enableToneMap(self):
        _check(lib().s360_isp_set_tone_map(self.h, 1))

    def disableToneMap(self):
        _check(lib().s360_isp_set_tone_map(self.h, 0))

    def addBlackLevelOffset(self, offset):
        _check(lib().s360_isp_add_black_level_offset(self.h, offset))

    def tables(self, w, h):
        """compositeCCM (3, 3), toneCurveLut (4096, 3), vignette gains per column (w, 3) and per row (h, 3); host only."""
        ccm = np.zeros((3, 3), np.float32); lut = np.zeros((4096, 3), np.float32)
        vh = np.zeros((w, 3), np.float32); vv = np.zeros((h, 3), np.float32)
        _check(lib().s360_isp_tables(self.h, w, h, _p(ccm), _p(lut), _p(vh), _p(vv)))
        return ccm, lut, vh, vv

    def getImage(self, ctx, raw, swizzle=True):
        """loadImage(raw) + getImage(): raw (H, W) uint8 / uint16 Bayer -> (H / resize, W / resize, 3) uint8 / uint16, BGR when swizzle."""
        raw = np.asarray(raw)
        if raw.ndim != 2 or raw.dtype not in (np.uint8, np.uint16):
            raise VrCamException("input is larger that 16 bits per pixel")
        if raw.strides[1] != raw.itemsize:
            raw = np.ascontiguousarray(raw)
        h, w = raw.shape
        out = np.zeros((h // self._resize, w // self._resize, 3), np.uint8 if self.output_bpp == 8 else np.uint16)
        _check(lib().s360_isp_process(ctx.h, self.h, C.c_void_p(raw.ctypes.data), w, h, 8 * raw.itemsize, C.c_size_t(raw.strides[0]), _p(out), int(swizzle)))
        return out

    def frameImageDev(self, ctx, raw_dev_ptr, w, h, depth_bits, out_bgr8_dev_ptr, stride_bytes=0):
        """16-bit ISP result narrowed to the renderer's BGR8 input (high byte, like imread of Unpacker's PNG); device pointers."""
        _check(lib().s360_isp_process_frame_dev(ctx.h, self.h, C.c_void_p(raw_dev_ptr), w, h, depth_bits, C.c_size_t(stride_bytes), C.c_void_p(out_bgr8_dev_ptr)))

    def frameImagesDev(self, ctx, raw_dev_ptrs, w, h, depth_bits, out_bgr8_dev_ptrs, stride_bytes=0):
        """The cameras of one frame in shared launches (s360_isp_process_frames_dev)."""
        n = len(raw_dev_ptrs)
        a = (C.c_void_p * n)(*[int(p) for p in raw_dev_ptrs]); b = (C.c_void_p * n)(*[int(p) for p in out_bgr8_dev_ptrs])
        _check(lib().s360_isp_process_frames_dev(ctx.h, self.h, n, a, w, h, depth_bits, C.c_size_t(stride_bytes), b))

    def getImageDev(self, ctx, raw_dev_ptr, w, h, depth_bits, out_dev_ptr, stride_bytes=0, swizzle=True):
        """Device-pointer variant, asynchronous on the context stream (s360_isp_process_dev)."""
        _check(lib().s360_isp_process_dev(ctx.h, self.h, C.c_void_p(raw_dev_ptr), w, h, depth_bits, C.c_size_t(stride_bytes), C.c_void_p(out_dev_ptr), int(swizzle)))


    def footageImage(self, ctx, packed, bits, w, h, want_raw=False):
        """Unpacker.cpp:141-190 for one camera frame: packed footage samples (bytes as they lie in the .bin) -> (h, w, 3) uint16 BGR, the
        image the Unpacker writes as PNG; with want_raw also the widened (h, w) uint16 samples. The ISP must have 16 output bits."""
        packed = np.ascontiguousarray(packed, np.uint8).reshape(-1)
        if packed.size != w * h * bits // 8:
            raise VrCamException("packed frame size does not match w * h * bits / 8")
        out = np.zeros((h, w, 3), np.uint16)
        raw = np.zeros((h, w), np.uint16) if want_raw else None
        _check(lib().s360_isp_process_footage(ctx.h, self.h, _p(packed), bits, w, h, _p(out), _p(raw) if want_raw else None))
        return (out, raw) if want_raw else out

    def footageImagesDev(self, ctx, packed_dev_ptrs, bits, w, h, out_dev_ptrs, high_byte=False):
        """The cameras of one frame in shared launches, device pointers (s360_isp_process_footage_dev)."""
        n = len(packed_dev_ptrs)
        a = (C.c_void_p * n)(*[int(p) for p in packed_dev_ptrs]); b = (C.c_void_p * n)(*[int(p) for p in out_dev_ptrs])
        _check(lib().s360_isp_process_footage_dev(ctx.h, self.h, n, a, bits, w, h, b, int(high_byte)))


class RawConverter:
    """camera_isp/RawConverter.h:19-30 on the GPU: packed footage samples -> (h, w) uint16."""

    @staticmethod
    def _convert(ctx, raw_frame, bits, width, height):
        packed = np.ascontiguousarray(raw_frame, np.uint8).reshape(-1)
        if packed.size < width * height * bits // 8:
            raise VrCamException("packed frame smaller than width * height * bits / 8")
        out = np.zeros((height, width), np.uint16)
        _check(lib().s360_raw_convert(ctx.h, _p(packed), bits, width, height, _p(out)))
        return out

    @staticmethod
    def convert8Frame(ctx, raw_frame, width, height):
        return RawConverter._convert(ctx, raw_frame, 8, width, height)

    @staticmethod
    def convert12Frame(ctx, raw_frame, width, height):
        return RawConverter._convert(ctx, raw_frame, 12, width, height)


class FootageHeader(C.Structure):   # BinaryFootageFile::MetadataHeader, BinaryFootageFile.h:18-27
    _fields_ = [(n, C.c_uint32) for n in ("magic", "timestamp", "fileIndex", "fileCount", "width", "height", "bitsPerPixel", "numberOfCameras")]


class BinaryFootageFile:
    """camera_isp/BinaryFootageFile.h:16-62 (host I/O; no device needed): the capture software's .bin container."""

    def __init__(self, file_path):
        self.h = C.c_void_p()
        _check(lib().s360_footage_open(str(file_path).encode(), C.byref(self.h)))
        self._hd = FootageHeader()
        frames, fbytes = C.c_uint64(), C.c_uint64()
        _check(lib().s360_footage_info(self.h, C.byref(self._hd), C.byref(frames), C.byref(fbytes)))
        self._frames, self._frame_bytes = frames.value, fbytes.value

    def close(self):
        if self.h:
            lib().s360_footage_close(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def getMetadata(self):
        return self._hd

    def getNumberOfCameras(self):
        return self._hd.numberOfCameras

    def getNumberOfFrames(self):
        return self._frames

    def getFrameSize(self):
        return self._frame_bytes

    def getBitsPerPixel(self):
        return self._hd.bitsPerPixel

    def getFrame(self, frame_number, camera_number):
        """The frame's packed bytes (a copy of the mapped range)."""
        ptr = C.POINTER(C.c_uint8)()
        _check(lib().s360_footage_frame(self.h, frame_number, camera_number, C.byref(ptr)))
        return np.ctypeslib.as_array(ptr, shape=(self._frame_bytes,)).copy()


def saveFlowToFile(flow, path):  # CvUtil.cpp:159-182
    f = _c(flow, np.float32)
    _check(lib().s360_flow_save(path.encode(), _p(f), f.shape[1], f.shape[0]))


def readFlowFromFile(path):  # CvUtil.cpp:184-199
    ptr = C.POINTER(C.c_float)()
    w, h = C.c_int(), C.c_int()
    _check(lib().s360_flow_load(path.encode(), C.byref(ptr), C.byref(w), C.byref(h)))
    a = np.ctypeslib.as_array(ptr, shape=(h.value, w.value, 2)).copy()
    lib().s360_free(ptr)
    return a
