"""The caller of the renderer (surround360_b200/batch_process_video.py; reference: scripts/batch_process_video.py): argument surface,
quality presets, and the plan that assigns frames to ranks and chains -- host logic, no GPU. (End to end the driver is checked against
the oracle under host emulation by tools/emu/run_batch.py: a two-frame temporal chain and independent frames in two chains.)"""
import subprocess
import sys

import pytest

from conftest import ROOT
from surround360_b200 import api, batch_process_video as bpv

BASE = ["--root_dir", "/nonexistent", "--rig_json_file", "rig.json", "--flow_alg", "pixflow_low", "--quality", "8k"]


def plan(*extra):
    return bpv.make_plan(bpv.parse_args(BASE + list(extra)))


def test_video_mode_cuts_the_sequence_into_contiguous_chains():
    p = plan("--start_frame", "10", "--end_frame", "109", "--gpus", "4", "--chains", "3")
    assert len(p) == 4 and all(len(ch) == 3 for ch in p)
    seen = [f for r in p for ch in r for f in ch["frames"]]
    assert seen == list(range(10, 110))                                   # every frame once, in order over (rank, chain)
    for r in p:
        for ch in r:
            assert ch["temporal"] and ch["frames"] == list(range(ch["frames"][0], ch["frames"][-1] + 1))
    sizes = [len(ch["frames"]) for r in p for ch in r]
    assert max(sizes) - min(sizes) <= 1
    one = plan("--start_frame", "0", "--end_frame", "9", "--gpus", "1", "--chains", "1")   # the reference's single chain
    assert one == [[dict(frames=list(range(10)), temporal=True)]]
    few = plan("--start_frame", "0", "--end_frame", "1", "--gpus", "2", "--chains", "4")   # more chains than frames: empty chains are dropped
    assert [len(r) for r in few] == [2, 0]


def test_independent_mode_shards_frames_over_ranks():
    p = plan("--start_frame", "5", "--end_frame", "14", "--gpus", "2", "--chains", "2", "--independent")
    assert [sorted(f for ch in r for f in ch["frames"]) for r in p] == [[5, 7, 9, 11, 13], [6, 8, 10, 12, 14]]   # frame f -> rank f mod world
    assert all(not ch["temporal"] for r in p for ch in r)


def test_argument_checks_follow_the_reference_script():
    with pytest.raises(SystemExit, match="enable_pole_removal"):          # scripts/batch_process_video.py:164-166
        plan("--start_frame", "0", "--end_frame", "1", "--enable_pole_removal")
    with pytest.raises(SystemExit, match="resume"):
        plan("--start_frame", "0", "--end_frame", "1", "--resume")
    with pytest.raises(SystemExit, match="Unrecognized quality"):
        bpv.quality_params("5k")
    q = bpv.quality_params("6k")
    assert q == dict(eqr_width=6300, eqr_height=3072, final_eqr_width=6144, final_eqr_height=6144, sharpening=0.25)   # :188-193
    assert set(api.PRESETS) >= {"3k", "4k", "6k", "8k"}


def test_dryrun_prints_the_plan_without_a_gpu():
    r = subprocess.run([sys.executable, "-m", "surround360_b200.batch_process_video"] + BASE + ["--start_frame", "0", "--end_frame", "29", "--gpus", "2",
                       "--chains", "3", "--dryrun"], capture_output=True, text=True, cwd=ROOT)
    assert r.returncode == 0, r.stderr
    lines = r.stdout.strip().splitlines()
    assert len(lines) == 6 and lines[0].startswith("rank 0 chain 0: 5 frame(s) 000000..000004") and lines[-1].startswith("rank 1 chain 2: 5 frame(s) 000025..000029")
