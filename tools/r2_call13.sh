#!/bin/bash
# GPU call 13 of round 2: k_isp_prep with four pixels per thread (ISP + Unpacker tests, ISP timing), k_sweep_q with the warp vote that
# skips both phases when no pixel of the warp-step is un-gated (parity, one frame alone).
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_isp.py tests/test_gpu_unpack.py tests/test_gpu_flow.py -m gpu -x -q 2>&1 | tail -5 > gpurun_out/r2c13_tests.log
cat gpurun_out/r2c13_tests.log
timeout 600 python -m pytest tests/test_gpu_frame.py -m gpu -x -q -k "full_8k" 2>&1 | tail -3 >> gpurun_out/r2c13_tests.log
tail -3 gpurun_out/r2c13_tests.log
timeout 600 python bench.py --isp-only --no-cpu-baseline > gpurun_out/r2c13_isp.json 2> gpurun_out/r2c13_isp.err; cut -c1-1500 gpurun_out/r2c13_isp.json; tail -3 gpurun_out/r2c13_isp.err
B="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-extras"
timeout 300 $B --inflight 1 --latency-mode > gpurun_out/r2c13_bench_quadlane_vote_f1.json 2>> gpurun_out/r2c13_bench.err
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r2c13_bench*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        st = d.get("stage_ms_single_frame", {})
        print(f, round(d["value"], 2), round(d["ms_per_step"], 1), d["config"]["frames_per_step"], round(d["roofline"]["ns_per_wavefront_step"]), round(st.get("total", 0), 1), round(st.get("side_flow", 0), 1), round(st.get("poles", 0), 1), d["output_crc32"]["graph_replay_frame0"], {k: (round(v, 1) if isinstance(v, float) else v) for k, v in d.get("single_frame_latency", {}).items() if k.endswith("_ms")})
    except Exception as e:
        print(f, "unreadable", e)
PY
tail -3 gpurun_out/r2c13_bench.err
