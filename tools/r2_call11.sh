#!/bin/bash
# GPU call 11 of round 2: the two sweep modes in one library (template parameter RING): parity of the latency mode, frames/s and ms alone
# of both modes, shared-memory carve-out of the latency kernel (64 KB asked for / 32 KB / driver default).
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_flow.py tests/test_gpu_frame.py -m gpu -x -q -k "not cli" 2>&1 | tail -5 > gpurun_out/r2c11_tests.log
cat gpurun_out/r2c11_tests.log
B="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-extras"
timeout 300 $B --inflight 1 --latency-mode > gpurun_out/r2c11_bench_lat_f1.json 2>> gpurun_out/r2c11_bench.err
S360_SWEEP_CARVEOUT=14 timeout 300 $B --inflight 1 --latency-mode > gpurun_out/r2c11_bench_lat32k_f1.json 2>> gpurun_out/r2c11_bench.err
S360_SWEEP_CARVEOUT=-1 timeout 300 $B --inflight 1 --latency-mode > gpurun_out/r2c11_bench_latauto_f1.json 2>> gpurun_out/r2c11_bench.err
timeout 300 $B --inflight 1 > gpurun_out/r2c11_bench_default_f1.json 2>> gpurun_out/r2c11_bench.err
timeout 400 $B > gpurun_out/r2c11_bench_default.json 2>> gpurun_out/r2c11_bench.err
timeout 400 $B --latency-mode > gpurun_out/r2c11_bench_lat.json 2>> gpurun_out/r2c11_bench.err
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r2c11_bench*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        st = d.get("stage_ms_single_frame", {})
        print(f, round(d["value"], 2), round(d["ms_per_step"], 1), d["config"]["frames_per_step"], round(d["roofline"]["ns_per_wavefront_step"]), round(st.get("total", 0), 1), round(st.get("side_flow", 0), 1), round(st.get("poles", 0), 1), d["output_crc32"]["graph_replay_frame0"], {k: (round(v, 1) if isinstance(v, float) else v) for k, v in d.get("single_frame_latency", {}).items() if k.endswith("_ms")})
    except Exception as e:
        print(f, "unreadable", e)
PY
tail -5 gpurun_out/r2c11_bench.err
