#!/bin/bash
# GPU call 7 of round 2: TMA descriptor walk, soft-ISP tests and timing, sweep variants (footprint coefficients, truncation), frames in flight.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{
for a in "tensor 4 32 1" "tensor 4 32 0" "tensor 4 64 0" "tensor 8 16 0" "tensor 8 32 0" "tensor 4 256 0" "tensor 8 142 0" "bulk 144" "bulk 16"; do
  echo "== tma_probe $a"; timeout 60 tools/tma_probe $a; echo "rc=$?"
done
} > gpurun_out/r2c7_tma_probe.log 2>&1
cat gpurun_out/r2c7_tma_probe.log
timeout 900 python -m pytest tests/test_gpu_isp.py tests/test_gpu_flow.py -m gpu -x -q 2>&1 | tail -12 > gpurun_out/r2c7_tests.log
cat gpurun_out/r2c7_tests.log
timeout 600 python bench.py --isp-only > gpurun_out/r2c7_isp.json 2> gpurun_out/r2c7_isp.err; cat gpurun_out/r2c7_isp.json; tail -3 gpurun_out/r2c7_isp.err
B="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-extras"
cp surround360_b200/libs360b200.so /tmp/lib_default.so
for name in default vh trunc coef4; do
  if [ $name != default ]; then cp build_variants/libs360b200_$name.so surround360_b200/libs360b200.so; fi
  $B > gpurun_out/r2c7_bench_$name.json 2>> gpurun_out/r2c7_bench.err
  $B --inflight 1 > gpurun_out/r2c7_bench_${name}_f1.json 2>> gpurun_out/r2c7_bench.err
  $B --inflight 24 > gpurun_out/r2c7_bench_${name}_n24.json 2>> gpurun_out/r2c7_bench.err
done
cp /tmp/lib_default.so surround360_b200/libs360b200.so
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r2c7_bench*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d["value"], 2), round(d["ms_per_step"], 1), d["config"]["frames_per_step"], round(d["roofline"]["ns_per_wavefront_step"]), round(d.get("stage_ms_single_frame", {}).get("total", 0), 1), d["output_crc32"]["graph_replay_frame0"])
    except Exception as e:
        print(f, "unreadable", e)
PY
tail -5 gpurun_out/r2c7_bench.err
