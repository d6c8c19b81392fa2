#!/bin/bash
# GPU call 9 of round 2: the Unpacker front end on the GPU (tests), and the A/B of the sweep's input fetch: L1 prefetch of the row record /
# old flow several positions ahead (rec4, rec8), footprint prefetch distance (q2), cp.async ring in shared memory (async4, async8).
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_unpack.py tests/test_gpu_isp.py -m gpu -x -q 2>&1 | tail -8 > gpurun_out/r2c9_tests_unpack.log
cat gpurun_out/r2c9_tests_unpack.log
B="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-extras"
cp surround360_b200/libs360b200.so /tmp/lib_default.so
for name in default rec4 rec8 rec8q2 async4 async8 async8q2; do
  if [ $name != default ]; then cp build_variants/libs360b200_$name.so surround360_b200/libs360b200.so; fi
  timeout 400 $B > gpurun_out/r2c9_bench_$name.json 2>> gpurun_out/r2c9_bench.err
  timeout 300 $B --inflight 1 > gpurun_out/r2c9_bench_${name}_f1.json 2>> gpurun_out/r2c9_bench.err
  if [ $name = async8 ] || [ $name = rec8 ]; then
    timeout 600 python -m pytest tests/test_gpu_flow.py -m gpu -x -q -k "config1 or many_bands or ragged or alpha_gated" 2>&1 | tail -4 > gpurun_out/r2c9_tests_$name.log
    cat gpurun_out/r2c9_tests_$name.log
  fi
done
cp /tmp/lib_default.so surround360_b200/libs360b200.so
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r2c9_bench*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        st = d.get("stage_ms_single_frame", {})
        print(f, round(d["value"], 2), round(d["ms_per_step"], 1), d["config"]["frames_per_step"], round(d["roofline"]["ns_per_wavefront_step"]), round(st.get("total", 0), 1), round(st.get("side_flow", 0), 1), round(st.get("poles", 0), 1), d["output_crc32"]["graph_replay_frame0"])
    except Exception as e:
        print(f, "unreadable", e)
PY
tail -5 gpurun_out/r2c9_bench.err
