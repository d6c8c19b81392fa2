"""The drop-in proof for the operator boundary: the REFERENCE'S OWN renderer (TestRenderStereoPanorama.cpp & co. compiled unchanged from
/root/reference by oracle/ref_build.sh) with the adapter of integration/reference_tree/optical_flow/B200Flow.h registered in its flow
factory. `--side_flow_alg b200_pixflow_low --polar_flow_alg b200_pixflow_low` then makes the reference compute all 32 flows of a frame on
the B200 through the C ABI (s360_flow_compute_strided), everything else stays the reference's CPU code. The panorama must not change."""
import json
import os
import subprocess

import numpy as np
import pytest

import ref_bind as rb
from conftest import RIG17

BIN = os.path.join(rb.ORACLE_DIR, "_ref", "TestRenderStereoPanorama_b200")
pytestmark = pytest.mark.skipif(not os.path.exists(BIN), reason="oracle/_ref/TestRenderStereoPanorama_b200 not built")


def _cmd(imgs, out, side_alg, polar_alg):
    return [BIN, "--rig_json_file", RIG17, "--imgs_dir", str(imgs), "--frame_number", "000000", "--output_data_dir", str(out),
            "--prev_frame_data_dir", "NONE", "--output_equirect_path", str(out / "eqr.png"), "--enable_top", "--enable_bottom",
            "--sharpening", "0.25", "--eqr_width", "1400", "--eqr_height", "700", "--final_eqr_width", "1024", "--final_eqr_height", "1024",
            "--side_flow_alg", side_alg, "--polar_flow_alg", polar_alg]


def _write_inputs(tmp_path, frames17):
    cv2 = pytest.importorskip("cv2")
    rig = json.load(open(RIG17))
    imgs = tmp_path / "rgb"
    for cam, im in zip(rig["cameras"], frames17):
        (imgs / cam["id"]).mkdir(parents=True)
        assert cv2.imwrite(str(imgs / cam["id"] / "000000.png"), im)
    return imgs


def _outdir(tmp_path, name):
    out = tmp_path / name
    (out / "flow" / "000000").mkdir(parents=True)
    (out / "debug" / "000000" / "flow_images").mkdir(parents=True)
    return out


@pytest.mark.gpu
def test_reference_renderer_with_b200_flow_adapter(orig, frames17, tmp_path):
    import cv2
    import oracle_bind as ob
    imgs = _write_inputs(tmp_path, frames17)
    out_gpu = _outdir(tmp_path, "out_b200")
    p = subprocess.run(_cmd(imgs, out_gpu, "b200_pixflow_low", "b200_pixflow_low"), capture_output=True, text=True)
    assert p.returncode == 0, p.stderr[-3000:]
    kw = dict(eqr_width=1400, eqr_height=700, final_eqr_width=1024, final_eqr_height=1024, sharpening=0.25, enable_top=1, enable_bottom=1)
    o = orig.render_frame(frames17, ob.make_params(**kw))
    assert np.array_equal(cv2.imread(str(out_gpu / "eqr.png"), cv2.IMREAD_COLOR), o.get("equirect"))
    # the flow files the reference wrote came from the GPU: same bits as the CPU solver's
    from surround360_b200 import api
    ref_flow = rb.Frame(RIG17, frames17, rb.make_params(**kw))
    try:
        assert np.array_equal(api.readFlowFromFile(str(out_gpu / "flow" / "000000" / "flowLtoR_5.bin")), ref_flow.flow("flowLtoR_5"))
        assert np.array_equal(api.readFlowFromFile(str(out_gpu / "flow" / "000000" / "flow_bottom_right.bin")), ref_flow.flow("flow_bottom_right"))
    finally:
        ref_flow.close()
    # the reference's other algorithms and its error behaviour are untouched by the two added factory lines
    p = subprocess.run(_cmd(imgs, _outdir(tmp_path, "out_bad"), "nope", "pixflow_low"), capture_output=True, text=True)
    assert p.returncode != 0 and "unrecognized flow algorithm name: nope" in p.stderr


def test_adapter_fails_loudly_without_a_gpu(frames17, tmp_path):
    """CPU tier: with no CUDA device the adapter's constructor throws VrCamException (there is no CPU fallback behind the ABI)."""
    import shutil
    if shutil.which("nvidia-smi") and subprocess.run(["nvidia-smi", "-L"], capture_output=True).returncode == 0:
        pytest.skip("a GPU is present")
    imgs = _write_inputs(tmp_path, frames17)
    p = subprocess.run(_cmd(imgs, _outdir(tmp_path, "out"), "b200_pixflow_low", "pixflow_low"), capture_output=True, text=True)
    assert p.returncode != 0
    assert "VrCamException" in p.stderr or "CUDA" in p.stderr
