#!/bin/bash
# GPU call 3 of round 2: GPU tier on the rewritten flow kernels, bench (20 frames in flight and 1), the same with busy polling
# (A/B of the hand-off sleep), then the per-kernel counter pass over one frame.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r2c3_tests.log
cat gpurun_out/r2c3_tests.log
B="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-extras"
$B > gpurun_out/r2c3_bench.json 2> gpurun_out/r2c3_bench.err
$B --inflight 1 > gpurun_out/r2c3_bench_f1.json 2>> gpurun_out/r2c3_bench.err
cp surround360_b200/libs360b200.so /tmp/lib_default.so
cp build_variants/libs360b200_nosleep.so surround360_b200/libs360b200.so
$B > gpurun_out/r2c3_bench_nosleep.json 2>> gpurun_out/r2c3_bench.err
$B --inflight 1 > gpurun_out/r2c3_bench_nosleep_f1.json 2>> gpurun_out/r2c3_bench.err
cp /tmp/lib_default.so surround360_b200/libs360b200.so
SMALL="python bench.py --inflight 1 --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-extras"
M=gpu__time_duration.sum,smsp__inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread
timeout 900 ncu --metrics $M --clock-control none --launch-skip 2700 -c 1100 --csv --log-file gpurun_out/r2c3_frame_counters.csv $SMALL > gpurun_out/r2c3_ncu1.log 2>&1
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r2c3_bench*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d["value"], 2), round(d["ms_per_step"], 1), round(d["roofline"]["ns_per_wavefront_step"]), d.get("stage_ms_single_frame"))
    except Exception as e:
        print(f, "unreadable", e)
PY
tail -3 gpurun_out/r2c3_bench.err
