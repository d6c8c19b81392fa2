#!/usr/bin/env python
"""tools/emu: run GPU-tier tests of this repository against the host emulation of the whole C ABI
(libs360b200_emu.so = every .cu of surround360_b200/csrc rewritten by emu_prep.py + rig.cpp). DEVELOPMENT TOOL ONLY:
a way to check a change bit-for-bit before GPU time is spent; it is far too slow for real frame sizes, it does not
model CUDA graphs (contexts are switched to plain launches), and nothing in the product, tests/ or bench.py uses it.

  python tools/emu/run_tests.py tests/test_gpu_ops.py [-k expr] [pytest args...]"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "surround360_b200", "csrc")
BUILD = os.environ.get("EMU_BUILD", os.path.join(HERE, "build"))


def build():
    os.makedirs(BUILD, exist_ok=True)
    objs = []
    common = ["g++", "-std=c++20", "-O1", "-ffp-contract=off", "-fPIC", "-c", "-I", os.path.join(HERE, "stub"), "-I", CSRC,
              "-I", os.path.join(ROOT, "include")]
    for name in ("warp_kernels", "finish_kernels", "flow_kernels", "preview_kernels", "isp_kernels", "pipeline", "capi"):
        gen = os.path.join(BUILD, name + "_emu.cpp")
        subprocess.check_call([sys.executable, os.path.join(HERE, "emu_prep.py"), os.path.join(CSRC, name + ".cu"), gen], stdout=subprocess.DEVNULL)
        obj = os.path.join(BUILD, name + "_emu.o")
        if not os.path.exists(obj) or os.path.getmtime(obj) < max(os.path.getmtime(os.path.join(CSRC, name + ".cu")), os.path.getmtime(os.path.join(HERE, "emu_cuda.h"))):
            subprocess.check_call(common + ["-o", obj, gen])
        objs.append(obj)
    for src, obj in ((os.path.join(CSRC, "rig.cpp"), "rig_emu.o"), (os.path.join(CSRC, "footage.cpp"), "footage_emu.o"), (os.path.join(HERE, "emu_rt.cpp"), "emu_rt.o")):
        o = os.path.join(BUILD, obj)
        subprocess.check_call(common + ["-o", o, src])
        objs.append(o)
    so = os.path.join(BUILD, "libs360b200_emu.so")
    subprocess.check_call(["g++", "-shared", "-pthread", "-o", so] + objs)
    return so


def main():
    so = build()
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from surround360_b200 import api
    api.LIB_PATH = so
    orig_init = api.Context.__init__

    def init(self, *a, **kw):
        orig_init(self, *a, **kw)
        self.set_graphs(False)   # graphs are not modelled

    api.Context.__init__ = init
    import pytest
    sys.exit(pytest.main(["-m", "gpu", "-x", "-q", "-p", "no:cacheprovider"] + sys.argv[1:]))


if __name__ == "__main__":
    main()
