"""Video mode through files (TRSP:200-255, 413-452): the state a frame hands to its successor -- flow fields and flow input
images -- imported into a fresh context (s360_temporal_import_*) and carried between two runs of the command-line tool
(--prev_frame_data_dir) must give the pixels of the in-process chain (s360_ctx_set_temporal)."""
import numpy as np
import pytest

from surround360_b200 import api

pytestmark = pytest.mark.gpu


def test_temporal_hand_over_through_host_buffers_and_files(frames17, tmp_path, rig_path):
    """The reference hands a frame's flows and flow input images to the next PROCESS through .bin / PNG files
    (TRSP:200-255, 413-452). Importing that state (s360_temporal_import_*) into a fresh context, and the command-line
    tool with --prev_frame_data_dir, must give the pixels of the in-process chain."""
    import json
    import os
    import subprocess
    from conftest import RIG17, ROOT
    cv2 = pytest.importorskip("cv2")
    kw = dict(eqr_width=1400, eqr_height=700, final_eqr_width=1024, final_eqr_height=1024, enable_top=1, enable_bottom=1, sharpening=0.25)
    f0, f1 = frames17, [np.ascontiguousarray(np.roll(im, 5, axis=1)) for im in frames17]
    P = api.default_params(**kw)
    # in-process chain
    a = api.Context(0); a.load_rig(RIG17); a.set_temporal(True)
    a.renderStereoPanorama(P, f0)
    n_side = 14
    side = [(a.intermediate("flowLtoR", i), a.intermediate("flowRtoL", i), a.intermediate("overlapL", i), a.intermediate("overlapR", i))
            for i in range(n_side)]
    pole = [(a.intermediate("poleFlow", k), a.intermediate("extFish", k)) for k in range(4)]
    chained = a.renderStereoPanorama(P, f1).copy()
    a.close()
    # state imported into a fresh context
    b = api.Context(0); b.load_rig(RIG17); b.set_temporal(True)
    b.frame_prepare(P, f1[0].shape[1], f1[0].shape[0])
    ow, oh, xw, xh = b.temporal_sizes()
    assert side[0][0].shape[:2] == (oh, ow) and pole[0][0].shape[:2] == (xh, xw)
    for i, (lr, rl, ol, orr) in enumerate(side):
        b.temporal_import_side(i, lr, rl, ol, orr)
    for k, (fl, xf) in enumerate(pole):
        b.temporal_import_pole(k, fl, xf)
    b.temporal_commit()
    assert np.array_equal(b.renderStereoPanorama(P, f1), chained)
    with pytest.raises(api.VrCamException, match="pair index out of range"):
        b.temporal_import_side(14, *side[0])
    b.close()
    c = api.Context(0); c.load_rig(RIG17)
    with pytest.raises(api.VrCamException, match="no frame plan"):
        c.temporal_commit()
    c.frame_prepare(P, f1[0].shape[1], f1[0].shape[0])
    with pytest.raises(api.VrCamException, match="temporal mode is off"):
        c.temporal_commit()
    c.close()
    # the command-line tool: frame 000000 with NONE, frame 000001 with --prev_frame_data_dir 000000
    exe = os.path.join(ROOT, "surround360_b200", "bin", "TestRenderStereoPanorama")
    if not os.path.exists(exe):
        pytest.skip("CLI not built")
    rig = json.load(open(rig_path))
    imgs, out = tmp_path / "rgb", tmp_path / "out"
    for name, fr in (("000000", f0), ("000001", f1)):
        for cam, im in zip(rig["cameras"], fr):
            (imgs / cam["id"]).mkdir(parents=True, exist_ok=True)
            assert cv2.imwrite(str(imgs / cam["id"] / f"{name}.png"), im)
        (out / "flow" / name).mkdir(parents=True)
        (out / "debug" / name / "flow_images").mkdir(parents=True)
    def run(frame, prev):
        cmd = [exe, "--rig_json_file", rig_path, "--imgs_dir", str(imgs), "--frame_number", frame, "--output_data_dir", str(out),
               "--prev_frame_data_dir", prev, "--output_equirect_path", str(out / f"eqr_{frame}.png"), "--enable_top", "--enable_bottom",
               "--sharpening", "0.25", "--eqr_width", "1400", "--eqr_height", "700", "--final_eqr_width", "1024", "--final_eqr_height", "1024"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
    run("000000", "NONE")
    assert np.array_equal(cv2.imread(str(out / "debug" / "000000" / "flow_images" / "overlap_3_L.png"), cv2.IMREAD_UNCHANGED), side[3][2])
    assert np.array_equal(cv2.imread(str(out / "debug" / "000000" / "flow_images" / "extendedFisheyeSpherical_bottom_left.png"),
                                     cv2.IMREAD_UNCHANGED), pole[2][1])
    run("000001", "000000")
    assert np.array_equal(cv2.imread(str(out / "eqr_000001.png"), cv2.IMREAD_COLOR), chained)
    os.remove(out / "flow" / "000000" / "flowRtoL_5.bin")
    r = subprocess.run([exe, "--rig_json_file", rig_path, "--imgs_dir", str(imgs), "--frame_number", "000001", "--output_data_dir", str(out),
                        "--prev_frame_data_dir", "000000", "--output_equirect_path", str(out / "x.png"), "--enable_top", "--enable_bottom",
                        "--eqr_width", "1400", "--eqr_height", "700", "--final_eqr_width", "1024", "--final_eqr_height", "1024"],
                       capture_output=True, text=True)
    assert r.returncode != 0 and "flowRtoL_5.bin" in r.stderr
