"""CPU tier: the TestHyperPreview tool (surround360_b200/csrc/cli/TestHyperPreview.cpp: the reference's flags and .bin footage layout,
test/TestHyperPreview.cpp:40-55, 187-424) linked against the host emulation of the library (tools/emu): two footage files with the
cameras interleaved, 12-bit samples, the three preview cameras picked by their position in the ascending serial order; the PNG of every
frame equals saturate_cast<uchar>(preview * 255 / 65535) of the library call on the same samples (which tests/test_gpu_preview.py holds
bit-identical to the reference's own TestHyperPreview.cpp on the GPU). No GPU needed; test infrastructure like the oracle."""
import os
import shutil
import subprocess
import sys

import numpy as np
import pytest

from conftest import RIG17, ROOT

sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "tools", "emu"))
cv2 = pytest.importorskip("cv2")


@pytest.mark.skipif(shutil.which("g++") is None, reason="needs g++")
def test_hyper_preview_tool_under_host_emulation(tmp_path):
    import ounpack
    from surround360_b200 import synth
    os.environ["EMU_BUILD"] = str(tmp_path / "emu")
    import run_tests
    run_tests.BUILD = os.environ["EMU_BUILD"]
    so = run_tests.build()
    exe = os.path.join(run_tests.BUILD, "TestHyperPreview_emu")
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-o", exe, os.path.join(ROOT, "surround360_b200", "csrc", "cli", "TestHyperPreview.cpp"),
                           "-L", run_tests.BUILD, "-ls360b200_emu", "-lz", "-pthread", "-Wl,-rpath," + run_tests.BUILD])
    # 5 cameras in 2 files (camera c in file c % 2), 2 frames; serial order != footage order
    w = h = 2048
    serials = [16400500, 16400100, 16400900, 16400300, 16400700]
    rng = np.random.default_rng(4)
    yy, xx = np.mgrid[0:h, 0:w]

    def image(k, f):
        base = 2000 + 1500 * np.sin(xx / (37.0 + 11 * k)) * np.cos(yy / (53.0 - 3 * k + f))
        return np.clip(base + rng.normal(0, 60, base.shape), 0, 4095).astype(np.uint16)
    frames = [[image(c, f) for c in range(5)] for f in range(2)]
    prefix = tmp_path / "bin"; prefix.mkdir()
    packed = {}
    for i in range(2):
        cams = [c for c in range(5) if c % 2 == i]
        p = synth.write_footage_bin(str(prefix / f"{i}.bin"), [[fr[c] for c in cams] for fr in frames], 12, [serials[c] for c in cams], file_index=i, file_count=2)
        for f in range(2):
            for j, c in enumerate(cams):
                packed[(f, c)] = p[f][j]
    dest = tmp_path / "preview"; dest.mkdir()
    order = sorted(range(5), key=lambda c: str(serials[c]))    # positions in the ascending serial order -> footage camera
    top, bottom, bottom2 = 0, 3, 4                              # flags: indices into the sorted order
    r = subprocess.run([exe, "--binary_prefix", str(prefix), "--file_count", "2", "--rig_json_file", RIG17, "--preview_dest", str(dest), "--eqr_width", "256",
                        "--eqr_height", "128", "--top_cam_index", str(top), "--bottom_cam_index", str(bottom), "--bottom_cam2_index", str(bottom2)],
                       capture_output=True, text=True, timeout=1500)
    assert r.returncode == 0, r.stderr[-800:]
    # expected frames: the library call (the same emulated library, in a process of its own) on the converter's output
    for f in range(2):
        for i, role in enumerate((top, bottom, bottom2)):
            np.save(tmp_path / f"raw_{f}_{i}.npy", ounpack.convert12_frame(packed[(f, order[role])], w, h))
    helper = f'''
import sys, numpy as np
sys.path.insert(0, {ROOT!r})
from surround360_b200 import api
api.LIB_PATH = {so!r}
ctx = api.Context(0); ctx.load_rig({RIG17!r})
for f in range(2):
    raws = [np.load({str(tmp_path)!r} + f"/raw_{{f}}_{{i}}.npy") for i in range(3)]
    np.save({str(tmp_path)!r} + f"/want_{{f}}.npy", ctx.hyperPreview(*raws, 256, 128, gamma=1.0, softmax_coef=30.0))
'''
    subprocess.check_call([sys.executable, "-c", helper], timeout=1500)
    for f in range(2):
        want = np.load(tmp_path / f"want_{f}.npy")
        want8 = np.nan_to_num(np.rint(want * np.float32(255.0 / 65535.0)), nan=0.0).clip(0, 255).astype(np.uint8)
        got = cv2.imread(str(dest / f"{f:06d}.png"), cv2.IMREAD_COLOR)
        assert got is not None and np.array_equal(got, want8), f"frame {f}"
        assert want8.max() > 0
    assert not (dest / "000002.png").exists()
    r = subprocess.run([exe, "--binary_prefix", str(prefix), "--file_count", "3", "--rig_json_file", RIG17, "--preview_dest", str(dest)], capture_output=True, text=True)
    assert r.returncode != 0 and "error opening binary file" in r.stderr
    r = subprocess.run([exe, "--binary_prefix", str(prefix), "--file_count", "1", "--rig_json_file", RIG17, "--preview_dest", str(dest)], capture_output=True, text=True)
    assert r.returncode != 0 and "binary file count in metadata" in r.stderr
