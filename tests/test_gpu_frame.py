"""Whole-frame parity (s360_render_frame) against the CPU oracle on the 17-camera rig at a
size the oracle renders in seconds, stage by stage, plus size-independent properties at the
full 8K configuration."""
import numpy as np
import pytest

import oracle_bind as ob
from surround360_b200 import api

pytestmark = pytest.mark.gpu

SMALL = dict(eqr_width=2100, eqr_height=1024, final_eqr_width=2048, final_eqr_height=2048, sharpening=0.25, enable_top=1,
             enable_bottom=1, want_cubemap=1, cubemap_width=256, cubemap_height=256)
# Stated tolerance of the 8-bit outputs: bit-exact. (Every float op on the path is evaluated in
# the CPU path's order without FMA, and the libm-dependent deghost terms are host tables.)
MAX_LSB = 0


def close(g, o, what):
    assert g.shape == o.shape, what
    d = np.abs(g.astype(np.int32) - o.astype(np.int32)) if g.dtype == np.uint8 else np.abs(g - o)
    assert d.max() <= MAX_LSB, f"{what}: max diff {d.max()} on {np.count_nonzero(d)} of {d.size}"


@pytest.fixture(scope="module")
def rendered(ctx, orig, frames17):
    ctx.keep_intermediates(True)
    eq, cube = ctx.renderStereoPanorama(api.default_params(**SMALL), frames17)
    ref = orig.render_frame(frames17, ob.make_params(keep_intermediates=1, **SMALL))
    return eq, cube, ref


def test_geometry_matches(ctx, orig):
    for name in ("3k", "4k", "6k", "8k"):
        g = ctx.geometry(api.default_params(**api.PRESETS[name]))
        o = orig.geometry(ob.make_params(**api.PRESETS[name]))
        assert (g.cam_w, g.cam_h, g.overlap_w, g.num_novel) == (o["camW"], o["camH"], o["overlapW"], o["numNovel"])
        assert g.slab_disp == o["slabDisp"] and g.shift_px == o["shiftPx"] and g.h_radians == o["hRad"]
    g = ctx.geometry(api.default_params(**api.PRESETS["8k"]))
    assert (g.cam_w, g.cam_h, g.overlap_w, g.num_novel, g.pole_h) == (1814, 1769, 1214, 600, 2104)


def test_stage_projections(ctx, rendered):
    for i in range(14):
        close(ctx.intermediate("projection", i), rendered[2].get("projection", i), f"projection {i}")


def test_stage_side_flows(ctx, rendered):
    for i in range(14):
        close(ctx.intermediate("flowLtoR", i), rendered[2].get("flowLtoR", i), f"flowLtoR {i}")
        close(ctx.intermediate("flowRtoL", i), rendered[2].get("flowRtoL", i), f"flowRtoL {i}")


def test_stage_panoramas(ctx, rendered):
    close(ctx.intermediate("panoL"), rendered[2].get("panoL"), "panoL")
    close(ctx.intermediate("panoR"), rendered[2].get("panoR"), "panoR")


def test_stage_poles(ctx, rendered):
    close(ctx.intermediate("topSpherical"), rendered[2].get("topSpherical"), "topSpherical")
    close(ctx.intermediate("bottomSpherical"), rendered[2].get("bottomSpherical"), "bottomSpherical")
    for k in range(4):
        close(ctx.intermediate("poleFlow", k), rendered[2].get("poleFlow", k), f"poleFlow {k}")
        close(ctx.intermediate("poleWarped", k), rendered[2].get("poleWarped", k), f"poleWarped {k}")


def test_stage_eyes_and_outputs(ctx, rendered):
    eq, cube, ref = rendered
    close(ctx.intermediate("eyeL")[..., :3], ref.get("eyeL"), "eyeL")
    close(ctx.intermediate("eyeR")[..., :3], ref.get("eyeR"), "eyeR")
    close(eq, ref.get("equirect"), "equirect")
    close(cube, ref.get("cubemap"), "cubemap")


def test_outputs_match_the_reference_build(ctx, rendered, frames17):
    """The CUDA frame against the REFERENCE'S OWN CODE (oracle/_ref: TestRenderStereoPanorama.cpp & co. compiled unchanged,
    oracle/ref_build.sh), not only against the restatement: equirect, cubemap and flows, 0 LSB."""
    import ref_bind as rb
    from conftest import RIG17
    if not rb.available():
        pytest.skip("oracle/_ref not built")
    eq, cube, _ = rendered
    r = rb.Frame(RIG17, frames17, rb.make_params(**SMALL))
    try:
        close(eq, r.equirect, "equirect vs reference build")
        close(cube, r.cubemap, "cubemap vs reference build")
        for i in (0, 7, 13):
            close(ctx.intermediate("flowLtoR", i), r.flow(f"flowLtoR_{i}"), f"flowLtoR {i} vs reference build")
            close(ctx.intermediate("flowRtoL", i), r.flow(f"flowRtoL_{i}"), f"flowRtoL {i} vs reference build")
        for k, n in enumerate(("top_left", "top_right", "bottom_left", "bottom_right")):
            close(ctx.intermediate("poleFlow", k), r.flow("flow_" + n), f"pole flow {n} vs reference build")
    finally:
        r.close()


def test_side_only_no_sharpen_no_resize(ctx, orig, frames17):
    kw = dict(eqr_width=1400, eqr_height=700, final_eqr_width=1400, final_eqr_height=1400)   # resize guard of TRSP:938-941 -> skipped
    eq = ctx.renderStereoPanorama(api.default_params(**kw), frames17)
    ref = orig.render_frame(frames17, ob.make_params(**kw))
    close(eq, ref.get("equirect"), "equirect (side only)")
    assert eq.shape == (1400, 1400, 3)


def test_top_only_search20(ctx, orig, frames17):
    kw = dict(eqr_width=1400, eqr_height=700, final_eqr_width=1024, final_eqr_height=1024, enable_top=1)
    eq = ctx.renderStereoPanorama(api.default_params(side_flow_alg="pixflow_search_20", polar_flow_alg="pixflow_search_20", **kw), frames17)
    ref = orig.render_frame(frames17, ob.make_params(side_search20=1, polar_search20=1, **kw))
    close(eq, ref.get("equirect"), "equirect (top only, search_20)")


def test_eqr_width_must_divide_by_cameras(ctx, frames17):
    with pytest.raises(api.VrCamException, match="evenly divisible"):
        ctx.renderStereoPanorama(api.default_params(eqr_width=1401, eqr_height=700), frames17)


def test_full_8k_properties(frames17):
    """BASELINE's full size: properties that do not need the (minute-long) CPU render."""
    c = api.Context(0)
    from conftest import RIG17
    c.load_rig(RIG17)
    p = api.default_params(sharpening=0.25, enable_top=1, enable_bottom=1, **api.PRESETS["8k"])
    eq1 = c.renderStereoPanorama(p, frames17).copy()
    eq2 = c.renderStereoPanorama(p, frames17)
    assert eq1.shape == (8192, 8192, 3)
    assert np.array_equal(eq1, eq2)                      # deterministic despite the concurrent wavefront bands
    left, right = eq1[:4096].astype(np.int32), eq1[4096:].astype(np.int32)
    assert np.abs(left - right).mean() > 0.5             # the two eyes differ (parallax) ...
    assert np.abs(left - right).mean() < 40              # ... but show the same scene
    assert eq1.std() > 10
    # the batched side flows equal the same pairs solved alone through the operator entry point
    g = c.geometry(p)
    pa = c.intermediate("projection", 3)
    pb = c.intermediate("projection", 4)
    ov_l = np.ascontiguousarray(pa[:, g.cam_w - g.overlap_w:])
    ov_r = np.ascontiguousarray(pb[:, :g.overlap_w])
    single = c.computeOpticalFlow("pixflow_low", ov_l, ov_r, hint=api.DirectionHint.LEFT)
    assert np.array_equal(single, c.intermediate("flowLtoR", 3))
    # latency mode (s360_ctx_set_latency_mode: the sweep variant with the cp.async rings) renders the same bytes: through the plain
    # launch path (first frame of a new plan) and through the re-captured graph
    c.set_latency_mode(True)
    assert np.array_equal(c.renderStereoPanorama(p, frames17), eq1)
    assert np.array_equal(c.renderStereoPanorama(p, frames17), eq1)
    c.set_latency_mode(False)
    assert np.array_equal(c.renderStereoPanorama(p, frames17), eq1)
    c.close()


def test_cli_is_a_drop_in(ctx, frames17, tmp_path, rig_path):
    """The TestRenderStereoPanorama binary: reference flags in, PNG + flow .bin files out, same pixels
    as the library call."""
    import json
    import os
    import subprocess
    cv2 = pytest.importorskip("cv2")
    from conftest import ROOT
    exe = os.path.join(ROOT, "surround360_b200", "bin", "TestRenderStereoPanorama")
    if not os.path.exists(exe):
        pytest.skip("CLI not built")
    rig = json.load(open(rig_path))
    imgs = tmp_path / "rgb"
    for cam, im in zip(rig["cameras"], frames17):
        (imgs / cam["id"]).mkdir(parents=True)
        assert cv2.imwrite(str(imgs / cam["id"] / "000000.png"), im)
    out = tmp_path / "out"
    (out / "flow" / "000000").mkdir(parents=True)
    (out / "debug" / "000000" / "flow_images").mkdir(parents=True)      # both created by the driver script (batch_process_video.py:116-118)
    (out / "debug" / "000000" / "projections").mkdir(parents=True)
    kw = dict(eqr_width=1400, eqr_height=700, final_eqr_width=1024, final_eqr_height=1024, sharpening=0.25, enable_top=1, enable_bottom=1)
    cmd = [exe, "--rig_json_file", rig_path, "--imgs_dir", str(imgs), "--frame_number", "000000", "--output_data_dir", str(out),
           "--prev_frame_data_dir", "NONE", "--output_equirect_path", str(out / "eqr.png"), "--output_cubemap_path", str(out / "cube.png"),
           "--cubemap_width=128", "--cubemap_height", "128", "--cubemap_format", "photo", "--enable_top", "--enable_bottom=true",
           "--sharpening", "0.25", "--eqr_width", "1400", "--eqr_height", "700", "--final_eqr_width", "1024", "--final_eqr_height", "1024",
           "--v", "1", "--logbuflevel", "-1", "--stderrthreshold", "0", "--save_debug_images"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert "--- Runtime breakdown (sec) ---" in r.stderr and "Side optical flow:" in r.stderr
    crop = cv2.imread(str(out / "debug" / "000000" / "projections" / f"crop_{[c for c in rig['cameras'] if 'side' in c.get('group', '')][0]['id']}.png"), cv2.IMREAD_UNCHANGED)
    assert crop is not None and crop.shape[2] == 4 and os.path.exists(out / "debug" / "000000" / "sphericalImgL.png")
    eq_cli = cv2.imread(str(out / "eqr.png"), cv2.IMREAD_COLOR)
    cube_cli = cv2.imread(str(out / "cube.png"), cv2.IMREAD_COLOR)
    eq, cube = ctx.renderStereoPanorama(api.default_params(want_cubemap=1, cubemap_photo=1, cubemap_width=128, cubemap_height=128, **kw), frames17)
    assert np.array_equal(eq_cli, eq) and np.array_equal(cube_cli, cube)
    f = api.readFlowFromFile(str(out / "flow" / "000000" / "flowLtoR_3.bin"))
    assert np.array_equal(f, ctx.intermediate("flowLtoR", 3))
    assert os.path.exists(out / "flow" / "000000" / "flow_bottom_right.bin")
    # error behaviour: unknown flow algorithm -> VrCamException text, non-zero exit
    r = subprocess.run(cmd + ["--side_flow_alg", "nope"], capture_output=True, text=True)
    assert r.returncode != 0 and "unrecognized flow algorithm name: nope" in r.stderr
    # missing output directories fail like imwriteExceptionOnFail / saveFlowToFile do in the reference
    import shutil
    shutil.rmtree(out / "flow")
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode != 0 and "file not found" in r.stderr
    r = subprocess.run(cmd + ["--save_flow_data=false", "--nosave_debug_images"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


def test_pole_removal(ctx, orig, frames17):
    """--enable_pole_removal (PoleRemoval.cpp:32-188): red-mask cut, feather, flip180, flow, warp, merge, re-cut."""
    m1 = np.zeros((2048, 2048, 3), np.uint8); m1[900:1300, 980:1080] = (0, 0, 255)
    m2 = np.zeros((2048, 2048, 3), np.uint8); m2[700:1100, 1000:1100] = (0, 0, 255)
    m2[10:20, 10:20] = (0, 1, 255)    # not pure red: must not be cut
    kw = dict(eqr_width=1400, eqr_height=700, final_eqr_width=1024, final_eqr_height=1024, enable_bottom=1, enable_pole_removal=1)
    ctx.set_pole_masks(m1, m2)
    ctx.keep_intermediates(True)
    eq = ctx.renderStereoPanorama(api.default_params(**kw), frames17)
    ref = orig.render_frame(frames17, ob.make_params(keep_intermediates=1, **kw), (m1, m2))
    close(ctx.intermediate("bottomMerged"), ref.get("bottomMerged"), "bottomMerged")
    close(ctx.intermediate("bottomSpherical"), ref.get("bottomSpherical"), "bottomSpherical (pole removal)")
    close(eq, ref.get("equirect"), "equirect (pole removal)")
    c2 = api.Context(0)
    from conftest import RIG17
    c2.load_rig(RIG17)
    with pytest.raises(api.VrCamException, match="missing or bad pole mask"):
        c2.renderStereoPanorama(api.default_params(**kw), frames17)
    c2.close()


def test_temporal_chain_of_frames(orig, frames17):
    """Video mode: frame i regularises against frame i-1 (flows + images kept on the device)."""
    from conftest import RIG17
    c = api.Context(0)
    c.load_rig(RIG17)
    c.set_temporal(True)
    kw = dict(eqr_width=1400, eqr_height=700, final_eqr_width=1024, final_eqr_height=1024, enable_top=1, enable_bottom=1, sharpening=0.25)
    seq = [frames17, [np.ascontiguousarray(np.roll(im, 5, axis=1)) for im in frames17], frames17]
    state = ob.FrameState()
    outs = []
    for fr in seq:
        eq = c.renderStereoPanorama(api.default_params(**kw), fr)
        ref = orig.render_frame(fr, ob.make_params(**kw), state=state)
        close(eq, ref.get("equirect"), "equirect (temporal chain)")
        outs.append(eq.copy())
    assert not np.array_equal(outs[0], outs[2])      # same input, different history -> different result
    c.reset_temporal()
    eq = c.renderStereoPanorama(api.default_params(**kw), seq[0])
    assert np.array_equal(eq, outs[0])               # after a reset the first frame renders like NONE again
    c.close()



def test_two_interleaved_temporal_chains(orig, frames17):
    """Video mode with chains in flight (bench.py --temporal, SURVEY 8e option A): two contexts render two different frame ranges
    interleaved; each chain must equal the oracle's chain over its own range (the previous-frame state is per context)."""
    from conftest import RIG17
    kw = dict(eqr_width=1400, eqr_height=700, final_eqr_width=1024, final_eqr_height=1024, enable_top=1, enable_bottom=1, sharpening=0.25)
    roll = lambda k: [np.ascontiguousarray(np.roll(im, k, axis=1)) for im in frames17]
    seqs = [[frames17, roll(5), roll(9)], [roll(3), frames17, roll(7)]]
    ctxs, states = [], []
    for _ in seqs:
        c = api.Context(0)
        c.load_rig(RIG17)
        c.set_temporal(True)
        ctxs.append(c)
        states.append(ob.FrameState())
    try:
        for i in range(3):
            for c, st, seq in zip(ctxs, states, seqs):      # frame i of chain A, then frame i of chain B, ...
                eq = c.renderStereoPanorama(api.default_params(**kw), seq[i])
                ref = orig.render_frame(seq[i], ob.make_params(**kw), state=st)
                close(eq, ref.get("equirect"), f"equirect, frame {i} of an interleaved chain")
    finally:
        for c in ctxs:
            c.close()
