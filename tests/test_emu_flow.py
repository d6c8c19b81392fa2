"""CPU tier: the CUDA source of the flow solver (surround360_b200/csrc/flow_kernels.cu), rewritten for host threads by
tools/emu (kernel launches and inline PTX only) and executed there, gives the oracle's flow bit for bit on a small
pair. This is a source-level regression check for the kernels that needs no GPU; it is test infrastructure like the
oracle -- the product library is not involved and still requires a CUDA device."""
import os
import shutil
import subprocess
import sys

import pytest

from conftest import ROOT


@pytest.mark.skipif(shutil.which("g++") is None, reason="needs g++")
@pytest.mark.parametrize("mode", ["throughput", "latency"])
def test_flow_kernels_source_matches_oracle_under_host_emulation(tmp_path, mode):
    """throughput: k_sweep<DIR, TR, 0>; latency: k_sweep_q (four lanes per pixel, cp.async rings copied synchronously by the emulator)."""
    env = dict(os.environ, EMU_BUILD=str(tmp_path / "emu"))
    env.pop("S360_EMU_LATENCY", None)
    if mode == "latency":
        env["S360_EMU_LATENCY"] = "1"
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "emu", "check_flow.py"), "--sizes", "150x211"],
                       capture_output=True, text=True, env=env, timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert r.stdout.count("bit-exact") == 3 and "MISMATCH" not in r.stdout
