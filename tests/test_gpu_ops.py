"""GPU parity of the per-operator C-ABI entry points against the CPU oracle (bit-exact for
integer/byte work; the float paths are evaluated op-for-op without FMA so they are bit-exact
too). Sizes are chosen so the oracle finishes in seconds."""
import numpy as np
import pytest

import oracle_bind as ob
from surround360_b200 import api

pytestmark = pytest.mark.gpu
rng = np.random.default_rng(1234)


def same(a, b):
    assert a.shape == b.shape
    assert np.array_equal(a, b), f"mismatch: {np.count_nonzero(a != b)} of {a.size}, max {np.abs(a.astype(np.float64) - b.astype(np.float64)).max()}"


@pytest.mark.parametrize("shape,dst", [((131, 203, 4), (101, 65)), ((120, 333, 3), (320, 120)), ((64, 64, 4), (32, 32)),
                                       ((77, 50, 3), (91, 60)), ((9, 11, 4), (4, 5))])
def test_resize_cubic(ctx, shape, dst):
    a = rng.integers(0, 256, shape, dtype=np.uint8)
    same(ctx.resizeCubic(a, dst[0], dst[1]), ob.resize_cubic_u8(a, dst[0], dst[1]))


@pytest.mark.parametrize("offset", [49.13, -49.13, 0.5, 1.5, 209.7, 0.0, -0.49])
def test_offset_wrap_index_map_bit_exact(ctx, offset):
    a = rng.integers(0, 256, (16, 210, 4), dtype=np.uint8)
    same(ctx.offsetHorizontalWrap(a, offset), ob.offset_wrap(a, offset))
    b = rng.integers(0, 256, (5, 33, 3), dtype=np.uint8)
    same(ctx.offsetHorizontalWrap(b, offset), ob.offset_wrap(b, offset))


def test_flatten_deghost(ctx):
    a = rng.integers(0, 256, (64, 210, 4), dtype=np.uint8)
    b = rng.integers(0, 256, (64, 210, 4), dtype=np.uint8)
    b[:, :50, 3] = 0
    b[:, 50:100, 3] = 255
    b[:, 100:150, :3] = a[:, 100:150, :3]   # equal colours: the truncation-sensitive case
    same(ctx.flattenLayersDeghostPreferBase(a, b), ob.flatten_deghost(a, b))


def test_flatten_all_alpha_and_diffs(ctx):
    # every top alpha x a sweep of colour differences
    base = np.zeros((256, 256, 4), np.uint8)
    top = np.zeros((256, 256, 4), np.uint8)
    base[..., :3] = 200
    base[..., 3] = 255
    top[..., 3] = np.arange(256)[:, None]
    top[..., 0] = np.arange(256)[None, :]
    top[..., 1] = 200
    top[..., 2] = 255 - np.arange(256)[None, :]
    same(ctx.flattenLayersDeghostPreferBase(base, top), ob.flatten_deghost(base, top))


def test_feather_alpha(ctx):
    f = rng.integers(0, 256, (150, 300, 4), dtype=np.uint8)
    f[..., 3] = (rng.uniform(0, 1, (150, 300)) > 0.01) * 255
    f[:40, :, 3] = 0
    same(ctx.featherAlphaChannel(f, 31), ob.feather_alpha(f, 31))
    g = rng.integers(0, 256, (40, 50, 4), dtype=np.uint8)   # smaller than the kernel radius: reflect101 wraps repeatedly
    same(ctx.featherAlphaChannel(g, 31), ob.feather_alpha(g, 31))


def test_sharpen(ctx):
    img = rng.integers(0, 256, (90, 140, 3), dtype=np.uint8)
    blocky = np.ascontiguousarray(np.repeat(np.repeat(img[::6, ::6], 6, 0), 6, 1)[:90, :140])
    for im in (img, blocky):
        same(ctx.sharpen(im, 0.25), ob.sharpen(im, 0.25))
    same(ctx.sharpen(blocky, 1.0), ob.sharpen(blocky, 1.0))


@pytest.mark.parametrize("photo", [False, True])
def test_cubemap(ctx, photo):
    eye = rng.integers(0, 256, (128, 256, 3), dtype=np.uint8)
    same(ctx.cubemap(eye, 64, 48, photo), ob.cubemap(eye, 64, 48, photo))


def test_project_side_and_pole(ctx, orig):
    src = rng.integers(0, 256, (2048, 2048, 3), dtype=np.uint8)
    hr = float(orig.geometry(ob.make_params())["hRad"])
    for i in (0, 5, 13):
        cam = ctx.side_index(i)
        d = np.float32(np.float32(-np.float32(i) / np.float32(14) * np.float32(2.0)) * np.pi)
        l, r = float(np.float32(d + np.float32(hr / 2))), float(np.float32(d - np.float32(hr / 2)))
        g = ctx.bicubicRemapToSpherical(cam, src, 227, 221, l, r, hr / 2, -hr / 2, feather=100)
        o = orig.project_side(cam, src, 227, 221, l, r, hr / 2, -hr / 2, 100)
        same(g, o)
    # fisheye (FTHETA) camera through the same entry point, BGRA input, no feather
    src4 = rng.integers(0, 256, (2048, 2048, 4), dtype=np.uint8)
    top = ctx.special_camera(0)
    g = ctx.bicubicRemapToSpherical(top, src4, 300, 77, float(np.float32(2 * np.pi)), 0.0, float(np.float32(np.pi / 2)), 0.3)
    o = orig.project_side(top, src4, 300, 77, float(np.float32(2 * np.pi)), 0.0, float(np.float32(np.pi / 2)), 0.3, 0)
    same(g, o)


def test_novel_view_chunk(ctx):
    from surround360_b200 import synth
    i0, i1 = synth.pair_512()
    a, b = i0[:300, :400].copy(), i1[:300, :400].copy()
    flr = ob.pixflow(a, b, "pixflow_low", hint=3)
    frl = ob.pixflow(b, a, "pixflow_low", hint=1)
    gl, gr = ctx.combineLazyNovelViews(a, b, flr, frl, 200, 600, 200, 60.5)
    ol, orr = ob.novel_view_chunk(a, b, flr, frl, 200, 600, 200, 60.5)
    same(gl, ol)
    same(gr, orr)
    # buffer wider than the number of novel views: unfilled columns sample (0,0,t=0)
    gl, gr = ctx.combineLazyNovelViews(a, b, flr, frl, 210, 600, 200, 60.5)
    ol, orr = ob.novel_view_chunk(a, b, flr, frl, 210, 600, 200, 60.5)
    same(gl, ol)
    same(gr, orr)


def test_errors_are_reported_not_swallowed(ctx):
    from surround360_b200 import api
    a = np.zeros((64, 64, 4), np.uint8)
    with pytest.raises(api.VrCamException, match="unrecognized flow algorithm name"):
        ctx.computeOpticalFlow("no_such_flow", a, a)
    with pytest.raises(api.VrCamException):
        ctx.bicubicRemapToSpherical(99, a, 8, 8, 0, 1, 0, 1)


def test_strided_and_device_pointer_variants(ctx, orig):
    """s360_flow_compute_strided / _dev, s360_novel_view_chunk_dev, s360_project_spherical_strided: same bits as the dense host calls."""
    import ctypes as C
    torch = pytest.importorskip("torch")
    from surround360_b200 import synth
    i0, i1 = synth.pair_512()
    wide0 = np.zeros((200, 400, 4), np.uint8); wide1 = np.zeros((200, 400, 4), np.uint8)
    wide0[:, 100:356] = i0[:200, :256]; wide1[:, 37:293] = i1[:200, :256]
    v0, v1 = wide0[:, 100:356], wide1[:, 37:293]            # views: row stride 1600 bytes, like cv::Mat ROIs
    dense = ctx.computeOpticalFlow("pixflow_low", np.ascontiguousarray(v0), np.ascontiguousarray(v1), hint=api.DirectionHint.LEFT)
    assert np.array_equal(ctx.computeOpticalFlowStrided("pixflow_low", v0, v1, hint=api.DirectionHint.LEFT), dense)
    assert np.array_equal(dense, ob.pixflow(np.ascontiguousarray(v0), np.ascontiguousarray(v1), "pixflow_low", hint=3))
    # device pointers with a stride
    d0 = torch.from_numpy(wide0).cuda(); d1 = torch.from_numpy(wide1).cuda()
    out = torch.zeros((200, 256, 2), dtype=torch.float32, device="cuda")
    for _ in range(2):   # the second call reuses the cached solver
        ctx.flow_compute_dev("pixflow_low", d0.data_ptr() + 100 * 4, 1600, d1.data_ptr() + 37 * 4, 1600, 256, 200, out.data_ptr(), hint=api.DirectionHint.LEFT)
        ctx.synchronize()
        assert np.array_equal(out.cpu().numpy(), dense)
    # novel view chunk on device pointers
    rng = np.random.default_rng(3)
    ovl = rng.integers(0, 256, (120, 90, 4), dtype=np.uint8); ovr = rng.integers(0, 256, (120, 90, 4), dtype=np.uint8)
    flr = rng.normal(size=(120, 90, 2)).astype(np.float32) * 3; frl = rng.normal(size=(120, 90, 2)).astype(np.float32) * 3
    hl, hr = ctx.combineLazyNovelViews(ovl, ovr, flr, frl, 50, 130, 45, 12.5)
    t = [torch.from_numpy(a).cuda() for a in (ovl, ovr, flr, frl)]
    ol = torch.zeros((120, 50, 4), dtype=torch.uint8, device="cuda"); orr = torch.zeros_like(ol)
    ctx.novel_view_chunk_dev(t[0].data_ptr(), t[1].data_ptr(), 0, t[2].data_ptr(), t[3].data_ptr(), 90, 120, 50, 130, 45, 12.5, ol.data_ptr(), orr.data_ptr())
    ctx.synchronize()
    assert np.array_equal(ol.cpu().numpy(), hl) and np.array_equal(orr.cpu().numpy(), hr)
    # projection with strides
    src = rng.integers(0, 256, (300, 320, 3), dtype=np.uint8)
    widesrc = np.zeros((300, 400, 3), np.uint8); widesrc[:, 40:360] = src
    dense_p = ctx.bicubicRemapToSpherical(1, src, 96, 80, 0.3, -0.3, 0.25, -0.25, feather=10)
    dst = np.zeros((80, 128, 4), np.uint8)
    sv = widesrc[:, 40:360]
    rc = api.lib().s360_project_spherical_strided(ctx.h, 1, C.c_void_p(sv.ctypes.data), C.c_size_t(sv.strides[0]), 320, 300, 3, C.c_float(0.3), C.c_float(-0.3),
                                                  C.c_float(0.25), C.c_float(-0.25), 10, C.c_void_p(dst.ctypes.data), C.c_size_t(dst.strides[0]), 96, 80)
    assert rc == 0 and np.array_equal(dst[:, :96], dense_p) and not dst[:, 96:].any()


def test_pole_to_side_operator(ctx, orig):
    """s360_pole_to_side (poleToSideFlowThread, TRSP:388-561) on its own against the oracle: feather, extension, flow, ramped warp,
    seam blend, alpha ramp, padding."""
    rng = np.random.default_rng(21)
    sw, sh, fh = 700, 350, 180
    from surround360_b200 import synth
    tex = synth.band_limited_noise(sh + 8, sw + 40, 77)
    tex = (255 * (tex - tex.min()) / max(tex.max() - tex.min(), 1e-6)).astype(np.uint8)
    side = np.zeros((sh, sw, 4), np.uint8)
    side[..., :3] = tex[:sh, :sw, None]; side[..., 3] = 255; side[:40, :, 3] = 0      # the side panorama has no data near the pole
    fish = np.zeros((fh, sw, 4), np.uint8)
    fish[..., :3] = tex[2:fh + 2, 5:sw + 5, None]; fish[..., 3] = 255
    p = api.default_params(eqr_width=sw, eqr_height=sh, enable_top=1, enable_bottom=1)
    warped, flow = ctx.poleToSideFlow(p, side, fish)
    ow, of = orig.pole_to_side(ob.make_params(eqr_width=sw, eqr_height=sh, enable_top=1, enable_bottom=1), side, fish)
    assert np.array_equal(flow, of)
    assert np.array_equal(warped, ow)
