#!/bin/bash
# GPU call 6 of round 2: TMA probe, not-yet-run GPU tests, sweep variants (software pipeline, re-used quotients), frames-in-flight curve of the new default.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for v in 0 1 2 3; do echo "== tma_probe $v"; timeout 60 tools/tma_probe $v; echo "rc=$?"; done > gpurun_out/r2c6_tma_probe.log 2>&1
timeout 120 compute-sanitizer --tool memcheck tools/tma_probe 0 > gpurun_out/r2c6_tma_probe_sanitizer.log 2>&1
cat gpurun_out/r2c6_tma_probe.log; tail -15 gpurun_out/r2c6_tma_probe_sanitizer.log
timeout 900 python -m pytest tests/test_gpu_preview.py tests/test_gpu_reference_adapter.py tests/test_gpu_ops.py tests/test_gpu_flow.py -m gpu -x -q 2>&1 | tail -8 > gpurun_out/r2c6_tests.log
cat gpurun_out/r2c6_tests.log
B="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-extras"
for F in 20 12 16 24; do $B --inflight $F > gpurun_out/r2c6_bench_default_n$F.json 2>> gpurun_out/r2c6_bench.err; done
$B --inflight 1 > gpurun_out/r2c6_bench_default_f1.json 2>> gpurun_out/r2c6_bench.err
cp surround360_b200/libs360b200.so /tmp/lib_default.so
for name in pipe vh vhpipe vh4 pipe5; do
  cp build_variants/libs360b200_$name.so surround360_b200/libs360b200.so
  $B > gpurun_out/r2c6_bench_$name.json 2>> gpurun_out/r2c6_bench.err
  case $name in pipe|vh|vhpipe) $B --inflight 1 > gpurun_out/r2c6_bench_${name}_f1.json 2>> gpurun_out/r2c6_bench.err;; esac
done
cp /tmp/lib_default.so surround360_b200/libs360b200.so
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r2c6_bench*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d["value"], 2), round(d["ms_per_step"], 1), d["config"]["frames_per_step"], round(d["roofline"]["ns_per_wavefront_step"]), round(d.get("stage_ms_single_frame", {}).get("total", 0), 1), d["output_crc32"])
    except Exception as e:
        print(f, "unreadable", e)
PY
tail -5 gpurun_out/r2c6_bench.err
