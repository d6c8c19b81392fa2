"""GPU parity of the per-operator C-ABI entry points against the CPU oracle (bit-exact for
integer/byte work; the float paths are evaluated op-for-op without FMA so they are bit-exact
too). Sizes are chosen so the oracle finishes in seconds."""
import numpy as np
import pytest

import oracle_bind as ob

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
