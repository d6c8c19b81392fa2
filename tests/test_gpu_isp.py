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


def mosaic16(bgr):
    """A 16-bit GBRG Bayer frame whose photosites sample the given 8-bit BGR image (what a camera of the rig would deliver)."""
    h, w = bgr.shape[:2]
    yy, xx = np.mgrid[0:h, 0:w]
    red = (yy % 2 == 1) & (xx % 2 == 0)
    blue = (yy % 2 == 0) & (xx % 2 == 1)
    m = np.where(red, bgr[..., 2], np.where(blue, bgr[..., 0], bgr[..., 1])).astype(np.uint16)
    return np.ascontiguousarray(m * 160 + 1400)   # above the black level, below saturation


def test_raw_to_panorama_on_device(frames17):
    """The fused chain: 17 raw frames -> s360_isp_process_frame_dev -> s360_render_frame_dev, nothing leaves the device
    in between. Against the reference's chain: its CameraIsp at 16 bits (what Unpacker writes as PNG), narrowed like the renderer's
    imread (high byte), rendered from host images."""
    torch = pytest.importorskip("torch")
    from test_gpu_frame import SMALL
    from conftest import RIG17
    c = api.Context(0)
    c.load_rig(RIG17)
    params = api.default_params(**SMALL)
    raws = [mosaic16(f) for f in frames17]
    h, w = raws[0].shape
    isp = api.CameraIsp(CONFIGS["sensor_a"], 16)
    isp.setBitsPerPixel(16); isp.setDemosaicFilter(2); isp.setResize(1); isp.enableToneMap()
    d_raw = [torch.from_numpy(r.view(np.int16)).cuda() for r in raws]
    d_bgr = [torch.zeros((h, w, 3), dtype=torch.uint8, device="cuda") for _ in raws]
    isp.frameImageDev(c, d_raw[0].data_ptr(), w, h, 16, d_bgr[0].data_ptr())                                         # one camera alone
    isp.frameImagesDev(c, [r.data_ptr() for r in d_raw[1:]], w, h, 16, [o.data_ptr() for o in d_bgr[1:]])            # the other 16 in shared launches
    eq_w, eq_h, cw, ch = c.output_size(params)
    d_eq = torch.zeros((eq_h, eq_w, 3), dtype=torch.uint8, device="cuda")
    d_cube = torch.zeros((ch, cw, 3), dtype=torch.uint8, device="cuda")
    c.render_frame_dev(params, [o.data_ptr() for o in d_bgr], w, h, d_eq.data_ptr(), d_cube.data_ptr())
    c.synchronize()
    # reference chain for the ISP part (3 cameras are enough for the bit check of the narrowing; all 17 feed the host render)
    host_frames = []
    for k, r in enumerate(raws):
        if rb.available() and k in (0, 7, 16):
            ref16 = run_ref("sensor_a", r, 2, 1, 16)
            assert np.array_equal(d_bgr[k].cpu().numpy(), (ref16 >> 8).astype(np.uint8)), f"camera {k}"
        host_frames.append(d_bgr[k].cpu().numpy())
    eq, cube = c.renderStereoPanorama(params, host_frames)
    assert np.array_equal(d_eq.cpu().numpy(), eq) and np.array_equal(d_cube.cpu().numpy(), cube)
    assert eq.std() > 10   # a real picture came out
    isp.close(); c.close()


@pytest.mark.skipif(not os.path.exists(rb.REF_ISP_B200_SO), reason="oracle/_ref/libref_isp_b200.so not built")
def test_reference_build_with_b200_isp_adapter():
    """integration/reference_tree/camera_isp/B200CameraIsp.h compiled INTO the reference build and driven through the reference's
    CameraIsp interface (virtual loadImage / getImage, the setters of Raw2Rgb.cpp:446-454): same bytes as the reference's own class."""
    raw = synth.bayer_frame(136, 200, np.uint16, seed=21)
    for dm, rs, bpp, tone, off, sw in ((2, 1, 16, True, 0, True), (0, 2, 8, False, 12, False)):
        a = rb.isp_process_b200(CONFIGS["sensor_a"], raw, dm, rs, bpp, tone, off, sw)
        b = rb.isp_process(CONFIGS["sensor_a"], raw, dm, rs, bpp, tone, off, sw)
        assert np.array_equal(a, b)
    with pytest.raises(RuntimeError, match="frequency"):
        rb.isp_process_b200(CONFIGS["sensor_a"], raw, 1)


def test_isp_batch_of_cameras(ctx):
    """s360_isp_process_frames_dev: the cameras of a frame through shared launches (blockIdx.z = image) == one call per camera."""
    torch = pytest.importorskip("torch")
    raws = [synth.bayer_frame(136, 200, np.uint16, seed=40 + k) for k in range(5)]
    isp = api.CameraIsp(CONFIGS["sensor_a"], 16)
    isp.setBitsPerPixel(16); isp.setDemosaicFilter(2); isp.setResize(1); isp.enableToneMap()
    single = [isp.getImage(ctx, r) for r in raws]
    if torch.cuda.is_available():
        d_raw = [torch.from_numpy(r.view(np.int16)).cuda() for r in raws]
        d_out = [torch.zeros((136, 200, 3), dtype=torch.uint8, device="cuda") for _ in raws]
        isp.frameImagesDev(ctx, [t.data_ptr() for t in d_raw], 200, 136, 16, [t.data_ptr() for t in d_out])
        ctx.synchronize()
        outs = [t.cpu().numpy() for t in d_out]
    else:   # tools/emu: "device" memory is host memory
        outs = [np.zeros((136, 200, 3), np.uint8) for _ in raws]
        isp.frameImagesDev(ctx, [r.ctypes.data for r in raws], 200, 136, 16, [o.ctypes.data for o in outs])
        ctx.synchronize()
    for o, s in zip(outs, single):
        assert np.array_equal(o, (s >> 8).astype(np.uint8))
    with pytest.raises(api.VrCamException, match="batch size"):
        isp.frameImagesDev(ctx, [0] * 40, 200, 136, 16, [0] * 40)
    isp.close()


def test_raw2rgb_cli(ctx, tmp_path):
    """surround360_b200/bin/Raw2Rgb with the reference's flags (Raw2Rgb.cpp:33-45): grey 16-bit PNG and headerless .raw input, 8- and
    16-bit PNG output; the files hold what the library call returns (and the reference build's result)."""
    import subprocess
    cv2 = pytest.importorskip("cv2")
    exe = os.path.join(os.path.dirname(HERE), "surround360_b200", "bin", "Raw2Rgb")
    raw = synth.bayer_frame(136, 200, np.uint16, seed=9)
    cfg = json.loads(CONFIGS["sensor_a"])
    cfg["CameraIsp"].update(width=200, height=136)
    cfg_path = str(tmp_path / "isp.json"); open(cfg_path, "w").write(json.dumps(cfg))
    png_in = str(tmp_path / "raw.png"); assert cv2.imwrite(png_in, raw)
    raw_in = str(tmp_path / "frame.raw"); raw.tofile(raw_in)
    for src, bpp, extra in ((png_in, 16, []), (raw_in, 8, ["--demosaic_filter", "0", "--disable_tone_curve"])):
        out = str(tmp_path / f"out{bpp}.png")
        p = subprocess.run([exe, "--input_image_path", src, "--output_image_path", out, "--isp_config_path", cfg_path, "--output_bpp", str(bpp)] + extra,
                           capture_output=True, text=True)
        assert p.returncode == 0, p.stderr[-400:]
        assert "Runtime = " in p.stderr
        got = cv2.imread(out, cv2.IMREAD_UNCHANGED)
        want = run_gpu(ctx, "sensor_a", raw, 0 if extra else 2, 1, bpp, not extra)
        assert got.dtype == want.dtype and np.array_equal(got, want)
        if rb.available():
            assert np.array_equal(got, run_ref("sensor_a", raw, 0 if extra else 2, 1, bpp, not extra))
    p = subprocess.run([exe, "--input_image_path", png_in, "--output_image_path", str(tmp_path / "x.png"), "--isp_config_path", cfg_path, "--demosaic_filter", "1"],
                       capture_output=True, text=True)
    assert p.returncode != 0 and "frequency" in p.stderr
    p = subprocess.run([exe, "--input_image_path", png_in, "--output_image_path", str(tmp_path / "x.png")], capture_output=True, text=True)
    assert p.returncode != 0 and "isp_config_path" in p.stderr
