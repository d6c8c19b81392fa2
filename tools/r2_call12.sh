#!/bin/bash
# GPU call 12 of round 2: latency mode with four lanes per pixel (k_sweep_q) against the one-lane ring kernel (S360_SWEEP_QUADLANE=0):
# parity, one frame alone, and a capture of the finest pole level.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_flow.py tests/test_gpu_frame.py -m gpu -x -q -k "latency or full_8k or config1 or many_bands" 2>&1 | tail -5 > gpurun_out/r2c12_tests.log
cat gpurun_out/r2c12_tests.log
B="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-extras"
timeout 300 $B --inflight 1 --latency-mode > gpurun_out/r2c12_bench_quadlane_f1.json 2>> gpurun_out/r2c12_bench.err
S360_SWEEP_QUADLANE=0 timeout 300 $B --inflight 1 --latency-mode > gpurun_out/r2c12_bench_ring_f1.json 2>> gpurun_out/r2c12_bench.err
timeout 300 $B --inflight 4 --latency-mode > gpurun_out/r2c12_bench_quadlane_n4.json 2>> gpurun_out/r2c12_bench.err
SMALL="python bench.py --inflight 1 --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-extras"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_sweep" -s 266 -c 2 -f -o gpurun_out/r2c12_sweep_pole_q $SMALL --latency-mode > gpurun_out/r2c12_ncu.log 2>&1
ncu -i gpurun_out/r2c12_sweep_pole_q.ncu-rep --page raw --csv > gpurun_out/r2c12_sweep_pole_q_raw.csv 2>/dev/null
ncu -i gpurun_out/r2c12_sweep_pole_q.ncu-rep --page source --print-source sass --csv > gpurun_out/r2c12_sweep_pole_q_sass.csv 2>/dev/null
rm -f gpurun_out/r2c12_sweep_pole_q.ncu-rep
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r2c12_bench*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        st = d.get("stage_ms_single_frame", {})
        print(f, round(d["value"], 2), round(d["ms_per_step"], 1), d["config"]["frames_per_step"], round(d["roofline"]["ns_per_wavefront_step"]), round(st.get("total", 0), 1), round(st.get("side_flow", 0), 1), round(st.get("poles", 0), 1), d["output_crc32"]["graph_replay_frame0"], {k: (round(v, 1) if isinstance(v, float) else v) for k, v in d.get("single_frame_latency", {}).items() if k.endswith("_ms")})
    except Exception as e:
        print(f, "unreadable", e)
PY
tail -5 gpurun_out/r2c12_bench.err
