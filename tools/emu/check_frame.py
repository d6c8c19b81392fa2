#!/usr/bin/env python
"""tools/emu: a tiny 17-camera frame through the emulated C ABI against the oracle, then the video-mode hand-over:
in-process chain == state imported into a fresh context (s360_temporal_import_*). DEVELOPMENT TOOL ONLY."""
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import run_tests  # noqa: E402

so = run_tests.build()
from surround360_b200 import api  # noqa: E402
api.LIB_PATH = so
import oracle_bind as ob  # noqa: E402
from conftest import RIG17  # noqa: E402

KW = dict(eqr_width=700, eqr_height=350, final_eqr_width=512, final_eqr_height=512, enable_top=1, enable_bottom=1, sharpening=0.25)


def ctx(temporal=False):
    c = api.Context(0)
    c.set_graphs(False)
    c.load_rig(RIG17)
    if temporal:
        c.set_temporal(True)
    return c


def frames():
    cache = "/tmp/s360_frames17_v1.npy"
    if os.path.exists(cache):
        return list(np.load(cache))
    from surround360_b200 import synth
    return synth.render_rig_frames(RIG17)


def latency_frame():
    """--latency: one frame with every context in latency mode (k_sweep_q inside the whole pipeline: side flows in column bands, pole flows
    with their gated regions in row bands) against the oracle and against the default kernels."""
    f0 = frames()
    P = api.default_params(**KW)
    t = time.time()
    a = ctx()
    a.set_latency_mode(True)
    eq = a.renderStereoPanorama(P, f0).copy()
    flows = [a.intermediate("flowLtoR", 3).copy(), a.intermediate("poleFlow", 1).copy()]
    print(f"latency-mode frame emulated in {time.time() - t:.0f} s")
    ref = ob.Rig.load(RIG17).render_frame(f0, ob.make_params(**KW))
    ok = np.array_equal(eq, ref.get("equirect"))
    print("latency-mode frame vs oracle:", "bit-exact" if ok else "MISMATCH")
    a.set_latency_mode(False)
    eq_t = a.renderStereoPanorama(P, f0)
    ok2 = np.array_equal(eq, eq_t) and np.array_equal(flows[0], a.intermediate("flowLtoR", 3)) and np.array_equal(flows[1], a.intermediate("poleFlow", 1))
    print("latency mode vs throughput mode (frame, a side flow, a pole flow):", "bit-exact" if ok2 else "MISMATCH")
    a.close()
    sys.exit(0 if ok and ok2 else 1)


def main():
    if "--latency" in sys.argv:
        latency_frame()
    f0 = frames()
    f1 = [np.ascontiguousarray(np.roll(im, 5, axis=1)) for im in f0]
    P = api.default_params(**KW)
    orig = ob.Rig.load(RIG17)
    t = time.time()
    a = ctx(temporal=True)
    eq0 = a.renderStereoPanorama(P, f0).copy()
    print(f"frame 0 emulated in {time.time() - t:.0f} s")
    state = ob.FrameState()
    ref0 = orig.render_frame(f0, ob.make_params(**KW), state=state)
    ok = np.array_equal(eq0, ref0.get("equirect"))
    print("frame 0 vs oracle:", "bit-exact" if ok else "MISMATCH")
    side = [(a.intermediate("flowLtoR", i), a.intermediate("flowRtoL", i), a.intermediate("overlapL", i), a.intermediate("overlapR", i)) for i in range(14)]
    pole = [(a.intermediate("poleFlow", k), a.intermediate("extFish", k)) for k in range(4)]
    chained = a.renderStereoPanorama(P, f1).copy()
    ref1 = orig.render_frame(f1, ob.make_params(**KW), state=state)
    ok1 = np.array_equal(chained, ref1.get("equirect"))
    print("frame 1 (in-process chain) vs oracle chain:", "bit-exact" if ok1 else "MISMATCH")
    a.close()
    b = ctx(temporal=True)
    b.frame_prepare(P, f1[0].shape[1], f1[0].shape[0])
    ow, oh, xw, xh = b.temporal_sizes()
    assert side[0][0].shape[:2] == (oh, ow) and pole[0][0].shape[:2] == (xh, xw), (side[0][0].shape, ow, oh, pole[0][0].shape, xw, xh)
    for i, s in enumerate(side):
        b.temporal_import_side(i, *s)
    for k, p in enumerate(pole):
        b.temporal_import_pole(k, *p)
    b.temporal_commit()
    imported = b.renderStereoPanorama(P, f1)
    ok2 = np.array_equal(imported, chained)
    print("frame 1 from imported state vs in-process chain:", "bit-exact" if ok2 else "MISMATCH")
    ok3 = not np.array_equal(chained, eq0)
    b.close()
    if "--cli" in sys.argv:   # the command-line tool linked against the emulated library: two processes, files in between
        import json
        import subprocess
        import tempfile
        import cv2
        exe = os.path.join(run_tests.BUILD, "TestRenderStereoPanorama_emu")
        subprocess.check_call(["g++", "-O1", "-std=c++17", "-o", exe, os.path.join(ROOT, "surround360_b200", "csrc", "cli", "TestRenderStereoPanorama.cpp"),
                               "-L", run_tests.BUILD, "-ls360b200_emu", "-lz", "-pthread", "-Wl,-rpath," + run_tests.BUILD])
        tmp = tempfile.mkdtemp(prefix="s360emu")
        rig = json.load(open(RIG17))
        imgs, out = os.path.join(tmp, "rgb"), os.path.join(tmp, "out")
        for name, fr in (("000000", f0), ("000001", f1)):
            for cam, im in zip(rig["cameras"], fr):
                os.makedirs(os.path.join(imgs, cam["id"]), exist_ok=True)
                assert cv2.imwrite(os.path.join(imgs, cam["id"], name + ".png"), im)
            os.makedirs(os.path.join(out, "flow", name))
            os.makedirs(os.path.join(out, "debug", name, "flow_images"))
        for frame, prev in (("000000", "NONE"), ("000001", "000000")):
            cmd = [exe, "--rig_json_file", RIG17, "--imgs_dir", imgs, "--frame_number", frame, "--output_data_dir", out, "--prev_frame_data_dir", prev,
                   "--output_equirect_path", os.path.join(out, f"eqr_{frame}.png"), "--enable_top", "--enable_bottom", "--sharpening", "0.25",
                   "--eqr_width", "700", "--eqr_height", "350", "--final_eqr_width", "512", "--final_eqr_height", "512"]
            r = subprocess.run(cmd, capture_output=True, text=True)
            assert r.returncode == 0, r.stderr
        ok4 = np.array_equal(cv2.imread(os.path.join(out, "eqr_000000.png"), cv2.IMREAD_COLOR), eq0)
        ok5 = np.array_equal(cv2.imread(os.path.join(out, "eqr_000001.png"), cv2.IMREAD_COLOR), chained)
        ok6 = np.array_equal(cv2.imread(os.path.join(out, "debug", "000000", "flow_images", "overlap_3_L.png"), cv2.IMREAD_UNCHANGED), side[3][2])
        print("CLI frame 0 vs library:", "bit-exact" if ok4 else "MISMATCH")
        print("CLI frame 1 with --prev_frame_data_dir 000000 vs in-process chain:", "bit-exact" if ok5 else "MISMATCH")
        print("CLI overlap_3_L.png vs intermediate:", "bit-exact" if ok6 else "MISMATCH")
        ok = ok and ok4 and ok5 and ok6
    sys.exit(0 if (ok and ok1 and ok2 and ok3) else 1)


if __name__ == "__main__":
    main()
