"""GPU soft ISP (surround360_b200/csrc/isp_kernels.cu) against the REFERENCE'S OWN CameraIsp compiled unchanged
(oracle/_ref/libref_isp.so, camera_isp/CameraIsp.h) through the call sequence of Raw2Rgb's CPU branch (Raw2Rgb.cpp:445-456), and
against the committed golden checksums of tests/golden/isp_golden.json. Bar: bit-exact (the pipeline is float arithmetic in the
reference's order; every libm value comes from host tables or the exact expf port)."""
import json
import os
import zlib

import numpy as np
import pytest

import ref_bind as rb
from surround360_b200 import api, synth

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
CONFIGS = json.load(open(os.path.join(HERE, "golden", "isp_configs.json")))   # the reference's res/config/isp/*.json + variants, verbatim text
GOLDEN = json.load(open(os.path.join(HERE, "golden", "isp_golden.json")))


@pytest.fixture(scope="module")
def ctx():
    c = api.Context(0)
    yield c
    c.close()


def run_gpu(ctx, cfg, raw, demosaic=2, resize=1, bpp=8, tone=True, offset=0, swizzle=True):
    isp = api.CameraIsp(CONFIGS[cfg], bpp)
    isp.setBitsPerPixel(16)
    isp.setDemosaicFilter(demosaic)
    isp.setResize(resize)
    (isp.enableToneMap if tone else isp.disableToneMap)()
    isp.addBlackLevelOffset(offset)
    out = isp.getImage(ctx, raw, swizzle)
    isp.close()
    return out


def run_ref(cfg, raw, demosaic=2, resize=1, bpp=8, tone=True, offset=0, swizzle=True):
    return rb.isp_process(CONFIGS[cfg], raw, demosaic, resize, bpp, tone, offset, swizzle)


CASES = [  # (name, config, (h, w), dtype, demosaic, resize, bpp, tone, offset, swizzle)
    ("a_edge_16", "sensor_a", (136, 200), np.uint16, 2, 1, 16, True, 0, True),
    ("a_edge_8", "sensor_a", (136, 200), np.uint16, 2, 1, 8, True, 0, True),
    ("a_bilinear_16", "sensor_a", (136, 200), np.uint16, 0, 1, 16, True, 0, True),
    ("b_edge_16_rgb", "sensor_b", (97, 131), np.uint16, 2, 1, 16, True, 25, False),
    ("passthrough_bilinear_8", "passthrough", (64, 96), np.uint16, 0, 1, 8, False, 0, True),
    ("a_edge_resize2", "sensor_a", (272, 400), np.uint16, 2, 2, 16, True, 0, True),
    ("a_bilinear_resize4", "sensor_a", (256, 384), np.uint16, 0, 4, 8, True, 0, True),
    ("a_edge_8bit_input", "sensor_a", (120, 168), np.uint8, 2, 1, 8, True, 0, True),
    ("rggb_edge_16", "variant_rggb", (136, 200), np.uint16, 2, 1, 16, True, 0, True),
    ("grbg_bilinear_16", "variant_grbg", (136, 200), np.uint16, 0, 1, 16, True, 0, True),
    ("bggr_edge_clamp_8", "variant_bggr_clamp", (136, 200), np.uint16, 2, 1, 8, True, 0, True),
    ("nosharpen_edge_16", "variant_nosharpen", (136, 200), np.uint16, 2, 1, 16, True, 0, True),
    ("lowcore_edge_16", "variant_lowcore", (136, 200), np.uint16, 2, 1, 16, True, 0, True),
]


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_isp_matches_reference_build(ctx, case):
    name, cfg, (h, w), dt, dm, rs, bpp, tone, off, sw = case
    raw = synth.bayer_frame(h, w, dt, seed=zlib.crc32(name.encode()) & 0xffff)
    got = run_gpu(ctx, cfg, raw, dm, rs, bpp, tone, off, sw)
    assert format(zlib.crc32(got.tobytes()), "08x") == GOLDEN[name], "differs from the golden checksum generated with the reference build"
    if rb.available():
        ref = run_ref(cfg, raw, dm, rs, bpp, tone, off, sw)
        assert got.shape == ref.shape and got.dtype == ref.dtype
        assert np.array_equal(got, ref), f"{(got != ref).sum()} of {got.size} samples differ, max |d| {np.abs(got.astype(np.int64) - ref.astype(np.int64)).max()}"


def test_isp_camera_size_2048(ctx):
    """The size the rig's cameras deliver (2048 x 2048, 16 bit): full frame against the reference build + golden checksum."""
    raw = synth.bayer_frame(2048, 2048, np.uint16, seed=7)
    got = run_gpu(ctx, "sensor_a", raw, 2, 1, 16)
    assert format(zlib.crc32(got.tobytes()), "08x") == GOLDEN["a_edge_16_2048"]
    if rb.available():
        assert np.array_equal(got, run_ref("sensor_a", raw, 2, 1, 16))


def test_isp_strided_input(ctx):
    raw = synth.bayer_frame(136, 232, np.uint16, seed=3)
    view = raw[:, :200]
    assert not view.flags["C_CONTIGUOUS"]
    a = run_gpu(ctx, "sensor_a", view)
    b = run_gpu(ctx, "sensor_a", np.ascontiguousarray(view))
    assert np.array_equal(a, b)


def test_isp_errors(ctx):
    raw = synth.bayer_frame(64, 64, np.uint16, seed=1)
    with pytest.raises(api.VrCamException, match="frequency"):
        run_gpu(ctx, "sensor_a", raw, demosaic=1)
    isp = api.CameraIsp(CONFIGS["sensor_a"], 8)
    with pytest.raises(api.VrCamException, match="Demosaic filter"):   # setDemosaicFilter's range check, CameraIsp.h:968-976
        isp.setDemosaicFilter(3)
    with pytest.raises(api.VrCamException, match="resize value"):      # setResize, :978-989
        isp.setResize(3)
    with pytest.raises(api.VrCamException):
        api.CameraIsp("{ not json", 8)
    with pytest.raises(api.VrCamException, match="8 or 16"):
        api.CameraIsp(CONFIGS["sensor_a"], 12)
    with pytest.raises(api.VrCamException, match="too small"):
        run_gpu(ctx, "sensor_a", synth.bayer_frame(6, 64, np.uint16, seed=1))
