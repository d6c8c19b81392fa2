"""Pins the oracle's PixFlow (oracle/opixflow.h) against the golden flow produced by the second,
independent restatement (python cv2 calls + numba loops, tests/golden/make_golden.py)."""
import os

import numpy as np

import oracle_bind as ob
from conftest import GOLDEN


def test_oracle_flow_equals_cv2_numba_restatement():
    g = np.load(os.path.join(GOLDEN, "pixflow_o1.npz"))
    flow = ob.pixflow(g["i0"], g["i1"], "pixflow_low", hint=3)
    d = np.abs(flow - g["flow"])
    assert d.max() == 0.0, f"max {d.max()} on {np.count_nonzero(d)} of {d.size}"


def test_level_dump_and_pyramid_sizes():
    g = np.load(os.path.join(GOLDEN, "pixflow_o1.npz"))
    flow, levels = ob.pixflow(g["i0"], g["i1"], "pixflow_low", hint=3, dump_levels=True)
    # 200x160 -> working 100x80; buildPyramid (PixFlow.h:477-491) stops before a side <= 24
    sizes = [(l.shape[1], l.shape[0]) for l in levels][::-1]
    w, h, expect = 100, 80, [(100, 80)]
    while True:
        w, h = int(np.float32(w) * np.float32(0.9) + np.float32(0.5)), int(np.float32(h) * np.float32(0.9) + np.float32(0.5))
        if w <= 24 or h <= 24:
            break
        expect.append((w, h))
    assert sizes == expect
    assert flow.shape == (160, 200, 2)


def test_search20_differs_only_through_initialisation():
    g = np.load(os.path.join(GOLDEN, "pixflow_o1.npz"))
    a = ob.pixflow(g["i0"], g["i1"], "pixflow_search_20", hint=0)   # UNKNOWN hint disables the search (PixFlow.h:372)
    b = ob.pixflow(g["i0"], g["i1"], "pixflow_low", hint=0)
    assert np.array_equal(a, b)
    c = ob.pixflow(g["i0"], g["i1"], "pixflow_search_20", hint=3)
    assert np.isfinite(c).all()


