"""CPU tier for the product: the C-ABI library loads without a GPU and exports every symbol
include/s360.h declares; host-side logic (flow files, params, rig generator); no compute calls."""
import ctypes as C
import json
import os
import re

import numpy as np
import pytest

from conftest import ROOT
from surround360_b200 import api, rig_gen


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "s360.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(s360_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = api.lib()
    names = declared_symbols()
    assert len(names) >= 35
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing


def test_version_and_default_params():
    assert b"sm_100a" in api.lib().s360_version()
    p = api.default_params()
    # defaults of the gflags in TestRenderStereoPanorama.cpp:44-70
    assert (p.interpupilary_dist, p.sharpening, p.zero_parallax_dist) == (6.4, 0.0, 10000.0)
    assert (p.side_alpha_feather_size, p.std_alpha_feather_size) == (100, 31)
    assert (p.eqr_width, p.eqr_height, p.final_eqr_width, p.final_eqr_height) == (256, 128, 3480, 960)
    assert (p.cubemap_width, p.cubemap_height, p.enable_top, p.enable_bottom) == (1536, 1536, 0, 0)
    assert p.side_flow_alg == b"pixflow_low" and p.polar_flow_alg == b"pixflow_low"
    with pytest.raises(KeyError):
        api.default_params(no_such_flag=1)


def test_no_cpu_fallback_without_a_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(api.VrCamException, match="no CUDA device|CUDA"):
        api.Context(0)


def test_flow_file_format_roundtrip(tmp_path):
    # CvUtil.cpp:159-199: int32 rows, int32 cols, then (f32 x, f32 y) row-major, native endian
    rng = np.random.default_rng(5)
    flow = rng.standard_normal((7, 5, 2)).astype(np.float32)
    path = str(tmp_path / "flowLtoR_0.bin")
    api.saveFlowToFile(flow, path)
    raw = open(path, "rb").read()
    assert len(raw) == 8 + 7 * 5 * 8
    assert np.frombuffer(raw[:8], np.int32).tolist() == [7, 5]
    assert np.array_equal(np.frombuffer(raw[8:], np.float32).reshape(7, 5, 2), flow)
    assert np.array_equal(api.readFlowFromFile(path), flow)
    with pytest.raises(api.VrCamException, match="file not found"):
        api.readFlowFromFile(str(tmp_path / "missing.bin"))
    open(path, "wb").write(raw[:20])
    with pytest.raises(api.VrCamException, match="truncated"):
        api.readFlowFromFile(path)


def test_rig_generator_matches_format(tmp_path):
    rig = rig_gen.make_rig()
    assert len(rig["cameras"]) == 17
    side = [c for c in rig["cameras"] if "side" in c["group"]]
    assert len(side) == 14 and all(c["type"] == "RECTILINEAR" for c in side)
    for c in rig["cameras"]:
        f, u, r = (np.array(c[k], float) for k in ("forward", "up", "right"))
        assert np.dot(np.cross(r, u), f) < 0          # right-handed as Camera::setRotation checks
        assert {"version", "type", "origin", "resolution", "focal", "id"} <= set(c)
    p = tmp_path / "rig.json"
    rig_gen.write_rig(str(p), n_side=8)
    assert len(json.load(open(p))["cameras"]) == 11


def test_presets_follow_batch_process_video():
    assert api.PRESETS["8k"] == dict(eqr_width=8400, eqr_height=4096, final_eqr_width=8192, final_eqr_height=8192)
    assert api.PRESETS["6k"]["eqr_width"] % 14 == 0 and api.PRESETS["3k"]["eqr_width"] % 14 == 0
