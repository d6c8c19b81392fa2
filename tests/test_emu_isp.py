"""CPU tier: the CUDA sources of the soft ISP, the Unpacker front end and the NV12 hand-off (isp_kernels.cu, finish_kernels.cu behind the C
ABI), executed on host threads by tools/emu, pass their GPU-tier tests: bit-identical to the reference build's golden checksums
(tests/golden/isp_golden.json, unpack_golden.json), to the reference's own code where /root/reference is present, and to the numpy
restatements. A source-level regression check that needs no GPU; test infrastructure like the oracle -- the product library is not
involved and still requires a CUDA device."""
import os
import shutil
import subprocess
import sys

import pytest

from conftest import ROOT


@pytest.mark.skipif(shutil.which("g++") is None, reason="needs g++")
def test_isp_unpack_nv12_sources_under_host_emulation(tmp_path):
    env = dict(os.environ, EMU_BUILD=str(tmp_path / "emu"))
    select = "(footage or raw_convert or matches_port or a_edge_16 or bilinear_resize4 or rggb or errors) and not on_device and not 2048"
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "emu", "run_tests.py"), "tests/test_gpu_unpack.py", "tests/test_nv12.py",
                        "tests/test_gpu_isp.py", "-k", select], capture_output=True, text=True, env=env, cwd=ROOT, timeout=1500)
    tail = r.stdout[-1500:] + r.stderr[-1500:]
    assert r.returncode == 0, tail
    assert " passed" in r.stdout and "failed" not in r.stdout, tail
