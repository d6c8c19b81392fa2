#!/bin/bash
# Final GPU call of round 2, at HEAD, most important first: GPU tier, the bench line as the driver invokes it, the ncu passes of the same
# tree (launch list = per-kernel counters of one frame; --set full of the finest side and pole sweeps, throughput and latency kernels).
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( time python -m pytest tests -m gpu -x -q ) 2>&1 | tail -18 > gpurun_out/r2f_tests.log
cat gpurun_out/r2f_tests.log
( time python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2f_bench_driver_style.json 2> gpurun_out/r2f_bench.err ) 2>&1 | tail -3
tail -2 gpurun_out/r2f_bench.err
SMALL="python bench.py --inflight 1 --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-extras"
M=gpu__time_duration.sum,smsp__inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread
timeout 900 ncu --metrics $M --clock-control none --launch-skip 2700 -c 1100 --csv --log-file gpurun_out/r2f_frame_counters.csv $SMALL > gpurun_out/r2f_ncu1.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_sweep" -s 194 -c 2 -f -o gpurun_out/r2f_sweep_side $SMALL > gpurun_out/r2f_ncu2.log 2>&1
ncu -i gpurun_out/r2f_sweep_side.ncu-rep --page raw --csv > gpurun_out/r2f_sweep_side_raw.csv 2>/dev/null
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_sweep" -s 266 -c 2 -f -o gpurun_out/r2f_sweep_pole $SMALL > gpurun_out/r2f_ncu3.log 2>&1
ncu -i gpurun_out/r2f_sweep_pole.ncu-rep --page raw --csv > gpurun_out/r2f_sweep_pole_raw.csv 2>/dev/null
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_sweep" -s 266 -c 2 -f -o gpurun_out/r2f_sweep_pole_latency $SMALL --latency-mode > gpurun_out/r2f_ncu4.log 2>&1
ncu -i gpurun_out/r2f_sweep_pole_latency.ncu-rep --page raw --csv > gpurun_out/r2f_sweep_pole_latency_raw.csv 2>/dev/null
ncu -i gpurun_out/r2f_sweep_pole_latency.ncu-rep --page source --print-source sass --csv > gpurun_out/r2f_sweep_pole_latency_sass.csv 2>/dev/null
rm -f gpurun_out/r2f_sweep_side.ncu-rep gpurun_out/r2f_sweep_pole.ncu-rep gpurun_out/r2f_sweep_pole_latency.ncu-rep
python bench.py --inflight 1 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-extras > gpurun_out/r2f_bench_f1.json 2>> gpurun_out/r2f_bench.err
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r2f_bench*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d["value"], 2), round(d["ms_per_step"], 1), d["config"]["frames_per_step"], d.get("e2e"), d.get("cpu_baseline", {}) and d["cpu_baseline"].get("value"), round(d.get("stage_ms_single_frame", {}).get("total", 0), 1), d.get("single_frame_latency"))
        if d.get("extra"): print(json.dumps({k: {kk: vv for kk, vv in v.items() if kk != "kernels"} for k, v in d["extra"].items()})[:2400])
    except Exception as e:
        print(f, "unreadable", e)
PY
