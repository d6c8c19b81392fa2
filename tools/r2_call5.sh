#!/bin/bash
# GPU call 5 of round 2: new operator / chain tests, TMA row pass (tests + A/B), sweep register / branch variants, frames-in-flight curve.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -m pytest tests/test_gpu_ops.py tests/test_gpu_frame.py -m gpu -x -q -k "strided or pole_to_side or interleaved or cli" 2>&1 | tail -8 > gpurun_out/r2c5_tests_new.log
cat gpurun_out/r2c5_tests_new.log
S360_TMA=1 timeout 600 python -m pytest tests/test_gpu_flow.py tests/test_gpu_frame.py -m gpu -x -q -k "config1 or many_bands or ragged or stage_side_flows or stage_poles" 2>&1 | tail -8 > gpurun_out/r2c5_tests_tma.log
cat gpurun_out/r2c5_tests_tma.log
B="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-extras"
$B > gpurun_out/r2c5_bench_default.json 2> gpurun_out/r2c5_bench.err
S360_TMA=1 timeout 600 $B > gpurun_out/r2c5_bench_tma.json 2>> gpurun_out/r2c5_bench.err
for F in 8 12 16; do $B --inflight $F > gpurun_out/r2c5_bench_inflight$F.json 2>> gpurun_out/r2c5_bench.err; done
cp surround360_b200/libs360b200.so /tmp/lib_default.so
for lib in build_variants/libs360b200_*.so; do
  name=$(basename "$lib" .so); name=${name#libs360b200_}
  cp "$lib" surround360_b200/libs360b200.so
  $B > gpurun_out/r2c5_bench_$name.json 2>> gpurun_out/r2c5_bench.err
  $B --inflight 1 > gpurun_out/r2c5_bench_${name}_f1.json 2>> gpurun_out/r2c5_bench.err
done
cp /tmp/lib_default.so surround360_b200/libs360b200.so
# ncu: the row pass with and without TMA at the finest side level (k_blur15_rows is launched twice per level: skip to the last levels)
SMALL="python bench.py --inflight 1 --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-extras"
timeout 600 ncu --set full --clock-control none -k regex:"k_blur15_rows" -s 120 -c 4 -f -o gpurun_out/r2c5_rows_manual $SMALL > gpurun_out/r2c5_ncu_a.log 2>&1
S360_TMA=1 timeout 600 ncu --set full --clock-control none -k regex:"k_blur15_rows" -s 120 -c 4 -f -o gpurun_out/r2c5_rows_tma $SMALL > gpurun_out/r2c5_ncu_b.log 2>&1
for n in manual tma; do ncu -i gpurun_out/r2c5_rows_$n.ncu-rep --page raw --csv > gpurun_out/r2c5_rows_${n}_raw.csv 2>/dev/null; done
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r2c5_bench*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d["value"], 2), round(d["ms_per_step"], 1), d["config"]["frames_per_step"], round(d["roofline"]["ns_per_wavefront_step"]), round(d.get("stage_ms_single_frame", {}).get("total", 0), 1))
    except Exception as e:
        print(f, "unreadable", e)
PY
tail -3 gpurun_out/r2c5_bench.err
