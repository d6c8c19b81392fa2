"""Pins the oracle's hand loops to the REFERENCE'S OWN CODE: oracle/_ref is built by oracle/ref_build.sh from
the unmodified sources under /root/reference (PixFlow.h, NovelView.cpp, CvUtil.cpp, Filter.h, ImageWarper.cpp,
PoleRemoval.cpp, Camera.cpp, RigDescription.cpp and TestRenderStereoPanorama.cpp itself) against stand-ins for the
absent third-party headers (oracle/ref_shim/: OpenCV primitives forward to the cv2-pinned oracle/ocv.h). What a
compiler makes of the reference -- overload resolution of exp()/cos()/abs(), integer promotions, float->uchar
truncation, evaluation order -- is then compared with the restatement in oracle/*.h: bit for bit.

CPU tier. Skipped when neither /root/reference nor a prebuilt oracle/_ref is present."""
import json
import os
import subprocess

import numpy as np
import pytest

import oracle_bind as ob
import ref_bind as rb
from conftest import RIG17
from surround360_b200 import synth

pytestmark = pytest.mark.skipif(not rb.available(), reason="oracle/_ref not built (needs /root/reference)")

SMALL = dict(eqr_width=2100, eqr_height=1024, final_eqr_width=2048, final_eqr_height=2048, sharpening=0.25, enable_top=1,
             enable_bottom=1, want_cubemap=1, cubemap_width=256, cubemap_height=256)
TINY = dict(eqr_width=1400, eqr_height=700, final_eqr_width=1024, final_eqr_height=1024, sharpening=0.25, enable_top=1, enable_bottom=1)


def same(a, b, what):
    assert a.shape == b.shape, f"{what}: {a.shape} vs {b.shape}"
    if not np.array_equal(a, b):
        d = np.abs(a.astype(np.float64) - b.astype(np.float64))
        raise AssertionError(f"{what}: max diff {d.max()} on {np.count_nonzero(d)} of {d.size}")


def test_reference_camera_unit_test_runs():
    """Camera::unitTest() (render/Camera.cpp:291-412) of the compiled reference: every CHECK in it holds."""
    rb.camera_unit_test()


def test_rig_description_and_camera_model(orig):
    s = rb.rig_summary(RIG17)
    assert s["n_side"] == orig.n_side == 14
    assert (s["top"], s["bottom"], s["bottom2"]) == tuple(orig.camera_id(orig.special(k)) for k in (0, 1, 2))
    g = orig.geometry(ob.make_params(eqr_width=8400, eqr_height=4096))
    assert np.float32(2 * s["h_fov"]) == g["hRad"] and np.float32(2 * s["v_fov"]) == g["vRad"]
    rng = np.random.default_rng(5)
    for cam in (0, 3, 9, 15, 16):
        for _ in range(20):
            p = rng.normal(size=3) * 300
            same(rb.camera_pixel(RIG17, cam, p), orig.pixel(cam, p), f"Camera::pixel cam {cam}")


@pytest.mark.parametrize("alg", ["pixflow_low", "pixflow_search_20"])
def test_pixflow_config1(alg):
    """BASELINE config 1 (512x512 pair): PixFlow::computeOpticalFlow of the reference == oracle, both hints of NovelView.cpp:289,297."""
    i0, i1 = synth.pair_512()
    for a, b, hint in ((i0, i1, 3), (i1, i0, 1)):
        same(rb.pixflow(a, b, alg, hint), ob.pixflow(a, b, alg, hint=hint), f"{alg} hint {hint}")


def test_pixflow_ragged_hints_and_temporal():
    i0, i1 = synth.pair_512()
    a = np.ascontiguousarray(i0[:181, :233]); b = np.ascontiguousarray(i1[:181, :233])
    a[:, :40, 3] = 0                                        # a gated region
    for hint in (0, 1, 2, 3, 4):
        alg = "pixflow_search_20" if hint else "pixflow_low"
        same(rb.pixflow(a, b, alg, hint), ob.pixflow(a, b, alg, hint=hint), f"ragged hint {hint}")
    prev = ob.pixflow(a, b, "pixflow_low", hint=2)
    pa = np.ascontiguousarray(np.roll(a, 3, axis=1)); pb = np.ascontiguousarray(np.roll(b, 3, axis=1))
    same(rb.pixflow(a, b, "pixflow_low", 2, prev=(prev, pa, pb)), ob.pixflow(a, b, "pixflow_low", hint=2, prev=(prev, pa, pb)), "temporal")


def test_blend_and_filter_operators(orig):
    rng = np.random.default_rng(11)
    base = rng.integers(0, 256, (97, 131, 4), dtype=np.uint8); top = rng.integers(0, 256, (97, 131, 4), dtype=np.uint8)
    top[:20, :, 3] = 0; top[20:40, :, 3] = 255
    same(rb.flatten_deghost(base, top), ob.flatten_deghost(base, top), "flattenLayersDeghostPreferBase")
    # every alpha against a spread of colour differences (the float exp() / tanhf() readings)
    a = np.zeros((256, 64, 4), np.uint8); t = np.zeros_like(a)
    a[..., :3] = 200; a[..., 3] = 255
    t[..., 3] = np.arange(256, dtype=np.uint8)[:, None]
    t[..., 0] = 200 - np.arange(64, dtype=np.uint8) * 3; t[..., 1] = 200; t[..., 2] = 200 - np.arange(64, dtype=np.uint8)
    same(rb.flatten_deghost(a, t), ob.flatten_deghost(a, t), "flatten: all alphas x colour differences")
    img = rng.integers(0, 256, (150, 210, 4), dtype=np.uint8); img[:, :, 3] = 255; img[40:90, 60:150, 3] = 0
    same(rb.feather_alpha(img, 31), ob.feather_alpha(img, 31), "featherAlphaChannel")
    pano = rng.integers(0, 256, (40, 420, 4), dtype=np.uint8)
    for off in (0.0, 196.53, -196.53, 0.5, 1.5, 419.5, -0.5):
        same(rb.offset_wrap(pano, off), ob.offset_wrap(pano, off), f"offsetHorizontalWrap {off}")
    eye = rng.integers(0, 256, (96, 200, 3), dtype=np.uint8)
    same(rb.sharpen(eye, 0.25), ob.sharpen(eye, 0.25), "iirLowPass + sharpenWithIirLowPass")
    same(rb.cubemap(eye, 32, 32, photo=False), ob.cubemap(eye, 32, 32, photo=False), "cubemap video")
    same(rb.cubemap(eye, 24, 40, photo=True), ob.cubemap(eye, 24, 40, photo=True), "cubemap photo")


def test_project_side_to_spherical(orig, frames17):
    g = orig.geometry(ob.make_params(eqr_width=1400, eqr_height=700))
    for side in (0, 6):
        direction = -np.float32(side) / np.float32(14) * np.float32(2.0) * np.float32(np.pi)
        src = frames17[orig.side_index(side)]
        l = np.float32(direction + g["hRad"] / 2); r = np.float32(direction - g["hRad"] / 2)
        t = np.float32(g["vRad"] / 2); b = np.float32(-g["vRad"] / 2)
        ref = rb.project_side(RIG17, side, src, g["camW"], g["camH"], l, r, t, b, feather=100)
        same(ref, orig.project_side(orig.side_index(side), src, g["camW"], g["camH"], l, r, t, b, feather=100), f"projectSideToSpherical {side}")


@pytest.fixture(scope="module")
def small_pair(orig, frames17):
    o = orig.render_frame(frames17, ob.make_params(keep_intermediates=1, **SMALL))
    r = rb.Frame(RIG17, frames17, rb.make_params(**SMALL))
    yield o, r
    r.close()


def test_whole_frame_reference_equals_oracle(small_pair):
    """renderStereoPanorama() of the compiled reference (TRSP:716-972) == the oracle's frame: equirect, cubemap, all 32 flows,
    the overlap images and the extended pole images."""
    o, r = small_pair
    same(r.equirect, o.get("equirect"), "equirect")
    same(r.cubemap, o.get("cubemap"), "cubemap")
    for i in range(14):
        same(r.flow(f"flowLtoR_{i}"), o.get("flowLtoR", i), f"flowLtoR {i}")
        same(r.flow(f"flowRtoL_{i}"), o.get("flowRtoL", i), f"flowRtoL {i}")
        ov = r.image(f"debug/000000/flow_images/overlap_{i}_L.png")
        same(ov, np.ascontiguousarray(o.get("projection", i)[:, -ov.shape[1]:]), f"overlap L {i}")
    for k, n in enumerate(("top_left", "top_right", "bottom_left", "bottom_right")):
        same(r.flow("flow_" + n), o.get("poleFlow", k), "pole flow " + n)


def test_temporal_chain_reference_equals_oracle(orig, frames17):
    """Two frames with --prev_frame_data_dir (TRSP:210-235, 418-436; PixFlow.h:101-118, 185-193)."""
    import tempfile
    work = tempfile.mkdtemp(prefix="s360_refchain_")
    kw = dict(eqr_width=1400, eqr_height=700, final_eqr_width=1024, final_eqr_height=1024, enable_top=1, enable_bottom=1, sharpening=0.25)
    seq = [frames17, [np.ascontiguousarray(np.roll(im, 5, axis=1)) for im in frames17]]
    state = ob.FrameState()
    try:
        prev = None
        for n, fr in enumerate(seq):
            name = f"{n:06d}"
            r = rb.Frame(RIG17, fr, rb.make_params(**kw), work_dir=work, frame=name, prev_frame=prev)
            # the next frame reads this frame's overlap / extended images back through imread(-1): keep them as PNG files
            for rel in [f"debug/{name}/flow_images/overlap_{i}_{s}.png" for i in range(14) for s in "LR"] + \
                       [f"debug/{name}/flow_images/extended{k}Spherical_{e}.png" for k in ("Side", "Fisheye")
                        for e in ("top_left", "top_right", "bottom_left", "bottom_right")]:
                import cv2
                assert cv2.imwrite(os.path.join(work, "out", rel), r.image(rel))
            o = orig.render_frame(fr, ob.make_params(**kw), state=state)
            same(r.equirect, o.get("equirect"), f"equirect of chained frame {n}")
            prev = name
    finally:
        import shutil
        shutil.rmtree(work, ignore_errors=True)


def test_pole_removal_reference_equals_oracle(orig, frames17):
    m1 = np.zeros((2048, 2048, 3), np.uint8); m1[900:1300, 980:1080] = (0, 0, 255)
    m2 = np.zeros((2048, 2048, 3), np.uint8); m2[700:1100, 1000:1100] = (0, 0, 255)
    m2[10:20, 10:20] = (0, 1, 255)
    kw = dict(eqr_width=1400, eqr_height=700, final_eqr_width=1024, final_eqr_height=1024, enable_bottom=1, enable_pole_removal=1)
    r = rb.Frame(RIG17, frames17, rb.make_params(**kw), pole_masks=(m1, m2))
    o = orig.render_frame(frames17, ob.make_params(**kw), (m1, m2))
    same(r.equirect, o.get("equirect"), "equirect (pole removal)")
    r.close()


def test_reference_binary_through_files(orig, frames17, tmp_path):
    """The reference's own main(): gflags command line, PNG files in, PNG + flow .bin out -- same pixels as the oracle."""
    cv2 = pytest.importorskip("cv2")
    rig = json.load(open(RIG17))
    imgs = tmp_path / "rgb"
    for cam, im in zip(rig["cameras"], frames17):
        (imgs / cam["id"]).mkdir(parents=True)
        assert cv2.imwrite(str(imgs / cam["id"] / "000000.png"), im)
    out = tmp_path / "out"
    (out / "flow" / "000000").mkdir(parents=True)
    (out / "debug" / "000000" / "flow_images").mkdir(parents=True)
    cmd = [rb.REF_BIN, "--rig_json_file", RIG17, "--imgs_dir", str(imgs), "--frame_number", "000000", "--output_data_dir", str(out),
           "--prev_frame_data_dir", "NONE", "--output_equirect_path", str(out / "eqr.png"), "--enable_top", "--enable_bottom=true",
           "--sharpening", "0.25", "--eqr_width", "1400", "--eqr_height", "700", "--final_eqr_width", "1024", "--final_eqr_height", "1024",
           "--v", "1", "--logbuflevel", "-1", "--stderrthreshold", "0"]
    p = subprocess.run(cmd, capture_output=True, text=True)
    assert p.returncode == 0, p.stderr[-2000:]
    assert "--- Runtime breakdown (sec) ---" in p.stderr and "Side optical flow:" in p.stderr
    o = orig.render_frame(frames17, ob.make_params(**TINY))
    same(cv2.imread(str(out / "eqr.png"), cv2.IMREAD_COLOR), o.get("equirect"), "equirect PNG of the reference binary")
    # error behaviour of the reference: VrCamException -> terminate handler -> abort (non-zero exit)
    p = subprocess.run(cmd + ["--side_flow_alg", "nope"], capture_output=True, text=True)
    assert p.returncode != 0 and "unrecognized flow algorithm name: nope" in p.stderr
