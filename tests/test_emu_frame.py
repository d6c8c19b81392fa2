"""CPU tier: a small 17-camera frame (top + bottom cameras, sharpening) through the whole pipeline with the latency-mode sweeps (k_sweep_q:
four lanes per pixel, 8-line bands, shared-memory rings), the CUDA sources executed on host threads by tools/emu: the equirect equals the
oracle's and the throughput-mode kernels' bit for bit. Source-level regression check without a GPU (test infrastructure; the product
library still requires a CUDA device)."""
import os
import shutil
import subprocess
import sys

import pytest

from conftest import ROOT


@pytest.mark.skipif(shutil.which("g++") is None, reason="needs g++")
def test_latency_mode_frame_under_host_emulation(tmp_path, frames17):   # frames17: makes sure the synthetic frames are cached in /tmp
    env = dict(os.environ, EMU_BUILD=str(tmp_path / "emu"))
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "emu", "check_frame.py"), "--latency"], capture_output=True, text=True, env=env,
                       cwd=ROOT, timeout=1700)
    tail = r.stdout[-1500:] + r.stderr[-1500:]
    assert r.returncode == 0, tail
    assert r.stdout.count("bit-exact") == 2 and "MISMATCH" not in r.stdout, tail
