#!/usr/bin/env python
"""tools/emu: surround360_b200.batch_process_video against the host emulation of the library, on a tiny preset -- a two-frame temporal
chain and two independent frames in two chains, compared with the oracle (FrameState chain / single frames). DEVELOPMENT TOOL ONLY
(the driver itself needs a GPU; this checks its planning, file layout and chain handling end to end without one)."""
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import run_tests  # noqa: E402

so = run_tests.build()
from surround360_b200 import api, batch_process_video as bpv  # noqa: E402
api.LIB_PATH = so
_init = api.Context.__init__


def init(self, *a, **kw):
    _init(self, *a, **kw)
    self.set_graphs(False)   # graphs are not modelled


api.Context.__init__ = init
TINY = dict(eqr_width=700, eqr_height=350, final_eqr_width=512, final_eqr_height=512)
api.PRESETS["tiny"] = TINY
import cv2  # noqa: E402
import oracle_bind as ob  # noqa: E402
from conftest import RIG17  # noqa: E402
from check_frame import frames  # noqa: E402


def main():
    import json
    f0 = frames()
    f1 = [np.ascontiguousarray(np.roll(im, 5, axis=1)) for im in f0]
    root = tempfile.mkdtemp(prefix="s360batch")
    ids = [c["id"] for c in json.load(open(RIG17))["cameras"]]
    for n, fr in ((0, f0), (1, f1)):
        for cid, im in zip(ids, fr):
            os.makedirs(os.path.join(root, "rgb", cid), exist_ok=True)
            assert cv2.imwrite(os.path.join(root, "rgb", cid, f"{n:06d}.png"), im)
    common = ["--root_dir", root, "--start_frame", "0", "--end_frame", "1", "--quality", "tiny", "--rig_json_file", RIG17, "--flow_alg", "pixflow_low",
              "--enable_top", "--enable_bottom", "--gpus", "1"]
    kw = dict(TINY, enable_top=1, enable_bottom=1, sharpening=0.25)
    orig = ob.Rig.load(RIG17)
    # 1. one temporal chain = the reference's behaviour (frame 1 regularised against frame 0)
    assert bpv.main(common + ["--chains", "1"]) == 0
    state = ob.FrameState()
    ref0 = orig.render_frame(f0, ob.make_params(**kw), state=state).get("equirect")
    ref1 = orig.render_frame(f1, ob.make_params(**kw), state=state).get("equirect")
    got0 = cv2.imread(os.path.join(root, "eqr_frames", "eqr_000000.png"), cv2.IMREAD_COLOR)
    got1 = cv2.imread(os.path.join(root, "eqr_frames", "eqr_000001.png"), cv2.IMREAD_COLOR)
    ok = np.array_equal(got0, ref0) and np.array_equal(got1, ref1)
    print("temporal chain of two frames vs oracle chain:", "bit-exact" if ok else "MISMATCH")
    # 2. independent frames in two chains
    assert bpv.main(common + ["--chains", "2", "--independent"]) == 0
    ind1 = orig.render_frame(f1, ob.make_params(**kw)).get("equirect")
    got0 = cv2.imread(os.path.join(root, "eqr_frames", "eqr_000000.png"), cv2.IMREAD_COLOR)
    got1 = cv2.imread(os.path.join(root, "eqr_frames", "eqr_000001.png"), cv2.IMREAD_COLOR)
    ok2 = np.array_equal(got0, ref0) and np.array_equal(got1, ind1) and not np.array_equal(ind1, ref1)
    print("independent frames in two chains vs oracle:", "bit-exact" if ok2 else "MISMATCH")
    sys.exit(0 if ok and ok2 else 1)


if __name__ == "__main__":
    main()
