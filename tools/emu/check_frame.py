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


def main():
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
    sys.exit(0 if (ok and ok1 and ok2 and ok3) else 1)


if __name__ == "__main__":
    main()
