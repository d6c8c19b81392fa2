#!/usr/bin/env python
"""tools/emu: build the host emulation of surround360_b200/csrc/flow_kernels.cu and compare PixFlow
(FlowBatch::run) bit for bit with the oracle on small image pairs. DEVELOPMENT TOOL ONLY: it lets a
kernel change be checked before GPU time is spent on it; nothing in the product, tests or bench uses it.

  python tools/emu/check_flow.py [--sizes 200x272,272x200] [--alg pixflow_low]"""
import argparse
import ctypes as C
import os
import subprocess
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def build(src=None):
    bdir = os.environ.get("EMU_BUILD", os.path.join(HERE, "build"))
    os.makedirs(bdir, exist_ok=True)
    gen = os.path.join(bdir, "flow_kernels_emu.cpp")
    subprocess.check_call([sys.executable, os.path.join(HERE, "emu_prep.py"), src or os.path.join(ROOT, "surround360_b200", "csrc", "flow_kernels.cu"), gen],
                          stdout=subprocess.DEVNULL)
    so = os.path.join(bdir, "libemu_flow.so")
    subprocess.check_call(["g++", "-std=c++20", "-O1", "-ffp-contract=off", "-fPIC", "-shared", "-pthread", "-I", os.path.join(HERE, "stub"),
                           "-I", os.path.join(ROOT, "surround360_b200", "csrc"), "-o", so, os.path.join(HERE, "emu_rt.cpp"),
                           os.path.join(HERE, "emu_flow.cpp"), gen])
    return C.CDLL(so)


def make_pair(w, h, seed):
    from surround360_b200 import synth
    rng = np.random.default_rng(seed)
    base = synth.band_limited_noise(h + 16, w + 16, seed)
    if base.ndim == 2:
        base = np.stack([base, np.roll(base, 3, 0), np.roll(base, 5, 1)], -1)
    base = (255 * (base - base.min()) / max(base.max() - base.min(), 1e-6)).astype(np.uint8) if base.dtype != np.uint8 else base
    i0 = np.dstack([base[8:8 + h, 8:8 + w, :3], np.full((h, w), 255, np.uint8)])
    i1 = np.dstack([base[6:6 + h, 11:11 + w, :3], np.full((h, w), 255, np.uint8)])
    # an alpha hole and a feathered edge so that the gate and the diffusion paths are exercised
    i0[h // 3:h // 3 + 20, w // 4:w // 4 + 30, 3] = 0
    ramp = np.clip(np.arange(w) * 16, 0, 255).astype(np.uint8)
    i1[:, :, 3] = np.minimum(i1[:, :, 3], ramp[None, :])
    return np.ascontiguousarray(i0), np.ascontiguousarray(i1)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sizes", default="200x272,272x200")
    ap.add_argument("--alg", default="pixflow_low")
    ap.add_argument("--src", default=None, help="alternative flow_kernels.cu (mutation tests, experiments)")
    a = ap.parse_args()
    import oracle_bind as ob
    L = build(a.src)
    L.emu_flow_batch.restype = C.c_int
    ok = True
    for si, s in enumerate(a.sizes.split(",")):
        w, h = map(int, s.split("x"))
        pairs = [make_pair(w, h, 10 + si), make_pair(w, h, 20 + si)]
        i0s = [p[0] for p in pairs] + [pairs[0][1]]     # third flow: the reciprocal direction of the first pair
        i1s = [p[1] for p in pairs] + [pairs[0][0]]
        hints = [3, 2, 1]
        n = len(i0s)
        out = np.zeros((n, h, w, 2), np.float32)
        arr0 = (C.c_void_p * n)(*[x.ctypes.data for x in i0s])
        arr1 = (C.c_void_p * n)(*[x.ctypes.data for x in i1s])
        hh = (C.c_int * n)(*hints)
        t = time.time()
        rc = L.emu_flow_batch(a.alg.encode(), n, arr0, arr1, hh, w, h, out.ctypes.data_as(C.c_void_p))
        dt = time.time() - t
        assert rc == 0, rc
        for k in range(n):
            ref = ob.pixflow(i0s[k], i1s[k], a.alg, hints[k])
            same = np.array_equal(out[k].view(np.uint32), np.asarray(ref, np.float32).view(np.uint32))
            d = np.abs(out[k] - ref).max()
            print(f"{s} flow {k} hint {hints[k]}: {'bit-exact' if same else 'MISMATCH'} (max |diff| {d:.3g}, |flow| max {np.abs(ref).max():.2f})  [{dt:.1f} s emulated]")
            ok &= same
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
