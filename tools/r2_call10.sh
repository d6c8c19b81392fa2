#!/bin/bash
# GPU call 10 of round 2: A/B of the footprint ring (own-flow footprint of the pixel 1-3 steps ahead copied with cp.async from the exact
# address, inputs through the cp.async ring) against the default and the L1-prefetch build; parity tests and a source-level capture of the winner candidate.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
B="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-extras"
cp surround360_b200/libs360b200.so /tmp/lib_default.so
for name in default aq8 aq4 aq8l3; do
  if [ $name != default ]; then cp build_variants/libs360b200_$name.so surround360_b200/libs360b200.so; fi
  timeout 300 $B --inflight 1 > gpurun_out/r2c10_bench_${name}_f1.json 2>> gpurun_out/r2c10_bench.err
  timeout 400 $B > gpurun_out/r2c10_bench_$name.json 2>> gpurun_out/r2c10_bench.err
  if [ $name = aq8 ]; then
    timeout 600 python -m pytest tests/test_gpu_flow.py tests/test_gpu_frame.py -m gpu -x -q 2>&1 | tail -4 > gpurun_out/r2c10_tests_$name.log
    cat gpurun_out/r2c10_tests_$name.log
    SMALL="python bench.py --inflight 1 --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-extras"
    timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_sweep" -s 266 -c 2 -f -o gpurun_out/r2c10_sweep_pole $SMALL > gpurun_out/r2c10_ncu.log 2>&1
    ncu -i gpurun_out/r2c10_sweep_pole.ncu-rep --page raw --csv > gpurun_out/r2c10_sweep_pole_raw.csv 2>/dev/null
    ncu -i gpurun_out/r2c10_sweep_pole.ncu-rep --page source --print-source sass --csv > gpurun_out/r2c10_sweep_pole_sass.csv 2>/dev/null
    rm -f gpurun_out/r2c10_sweep_pole.ncu-rep
  fi
done
cp /tmp/lib_default.so surround360_b200/libs360b200.so
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r2c10_bench*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        st = d.get("stage_ms_single_frame", {})
        print(f, round(d["value"], 2), round(d["ms_per_step"], 1), d["config"]["frames_per_step"], round(d["roofline"]["ns_per_wavefront_step"]), round(st.get("total", 0), 1), round(st.get("side_flow", 0), 1), round(st.get("poles", 0), 1), d["output_crc32"]["graph_replay_frame0"])
    except Exception as e:
        print(f, "unreadable", e)
PY
tail -5 gpurun_out/r2c10_bench.err
