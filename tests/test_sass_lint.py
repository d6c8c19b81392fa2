"""CPU tier: the built library's SASS. The path is bit-exact because no float multiply-add is contracted
(-fmad=false; the reference build has no FMA). ptxas 12.9 contracts the PACKED FP32 forms (mul.rn.f32x2 + add.rn.f32x2 ->
FFMA2) even with --fmad=false and explicit .rn, which no host emulation can see -- so the library must not contain them."""
import os
import re
import shutil
import subprocess

import pytest

from conftest import ROOT

LIB = os.path.join(ROOT, "surround360_b200", "libs360b200.so")


@pytest.mark.skipif(shutil.which("cuobjdump") is None or not os.path.exists(LIB), reason="needs cuobjdump and the built library")
def test_no_packed_fp32_contraction_and_sm100a_only():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    archs = set(re.findall(r"arch = (sm_\w+)", out))
    assert archs == {"sm_100a"}, archs
    ops = re.findall(r"^\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", out, flags=re.M)
    assert len(ops) > 10000
    packed = sorted({o for o in ops if o.split(".")[0] in ("FFMA2", "FADD2", "FMUL2")})
    assert not packed, f"packed FP32 instructions in the library: {packed}"
    # fused multiply-adds may only come from the compiler's division / square-root / exp expansions (which start from a MUFU
    # approximation) or from their written-out copies: a kernel without MUFU must not contain FFMA or DFMA at all
    for f in re.split(r"\n\s*Function : ", out)[1:]:
        name = f.split("\n", 1)[0].strip()
        fops = re.findall(r"^\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", f, flags=re.M)
        fma = sum(o.split(".")[0] in ("FFMA", "DFMA") for o in fops)
        mufu = sum(o.startswith("MUFU") for o in fops)
        assert fma == 0 or mufu > 0, f"{name}: {fma} fused multiply-adds without any division/sqrt expansion"
