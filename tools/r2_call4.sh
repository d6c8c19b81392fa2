#!/bin/bash
# GPU call 4 of round 2: GPU tier, the full default bench line (roofline_all, extras, cpu_baseline), A/B of the footprint
# prefetch in the sweep, compute-sanitizer memcheck (flow config 1 + a whole small frame) and racecheck (the shared-memory kernels).
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r2c4_tests.log
cat gpurun_out/r2c4_tests.log
python bench.py > gpurun_out/r2c4_bench_default.json 2> gpurun_out/r2c4_bench.err
B="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-extras"
$B --inflight 1 > gpurun_out/r2c4_bench_f1.json 2>> gpurun_out/r2c4_bench.err
cp surround360_b200/libs360b200.so /tmp/lib_default.so
cp build_variants/libs360b200_prefetch.so surround360_b200/libs360b200.so
$B > gpurun_out/r2c4_bench_prefetch.json 2>> gpurun_out/r2c4_bench.err
$B --inflight 1 > gpurun_out/r2c4_bench_prefetch_f1.json 2>> gpurun_out/r2c4_bench.err
cp /tmp/lib_default.so surround360_b200/libs360b200.so
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 3 python -m pytest tests/test_gpu_flow.py -q -k "config1" -x > gpurun_out/r2c4_memcheck_flow.log 2>&1; echo "memcheck flow rc=$?" >> gpurun_out/r2c4_memcheck_flow.log
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 3 python -m pytest tests/test_gpu_frame.py -q -k "stage_eyes_and_outputs or stage_poles" -x > gpurun_out/r2c4_memcheck_frame.log 2>&1; echo "memcheck frame rc=$?" >> gpurun_out/r2c4_memcheck_frame.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 3 python -m pytest tests/test_gpu_ops.py -q -x > gpurun_out/r2c4_racecheck_ops.log 2>&1; echo "racecheck ops rc=$?" >> gpurun_out/r2c4_racecheck_ops.log
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r2c4_bench*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d["value"], 2), round(d["ms_per_step"], 1), round(d["roofline"]["ns_per_wavefront_step"]), d.get("stage_ms_single_frame", {}).get("total"))
    except Exception as e:
        print(f, "unreadable", e)
PY
tail -3 gpurun_out/r2c4_bench.err; tail -4 gpurun_out/r2c4_memcheck_flow.log gpurun_out/r2c4_memcheck_frame.log gpurun_out/r2c4_racecheck_ops.log
