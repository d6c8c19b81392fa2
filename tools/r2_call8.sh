#!/bin/bash
# GPU call 8 of round 2: TMA alignment hypothesis + the fixed TMA row pass (tests, A/B, ncu), soft ISP with the line kernel (tests, batch
# timing, memcheck), preview exactness, and a source-level capture of the finest pole / side sweep (stall reasons per SASS line).
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{
for a in "tensor 8 142 0 968" "tensor 8 142 0 960" "tensor 8 144 0 960" "tensor 4 36 0 8" "tensor 4 36 0 16"; do
  echo "== tma_probe $a"; timeout 60 tools/tma_probe $a; echo "rc=$?"
done
} > gpurun_out/r2c8_tma_probe.log 2>&1
cat gpurun_out/r2c8_tma_probe.log
timeout 900 python -m pytest tests/test_gpu_isp.py tests/test_gpu_preview.py -m gpu -x -q 2>&1 | tail -12 > gpurun_out/r2c8_tests_isp.log
cat gpurun_out/r2c8_tests_isp.log
S360_TMA=1 timeout 900 python -m pytest tests/test_gpu_flow.py tests/test_gpu_frame.py -m gpu -x -q -k "config1 or many_bands or ragged or stage_side_flows or stage_poles or equirect" 2>&1 | tail -8 > gpurun_out/r2c8_tests_tma.log
cat gpurun_out/r2c8_tests_tma.log
timeout 400 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_isp.py -m gpu -x -q -k "a_edge_16 or a_bilinear_resize4 or batch_of_cameras" 2>&1 | tail -8 > gpurun_out/r2c8_memcheck_isp.log
cat gpurun_out/r2c8_memcheck_isp.log
timeout 600 python bench.py --isp-only > gpurun_out/r2c8_isp.json 2> gpurun_out/r2c8_isp.err; cat gpurun_out/r2c8_isp.json; tail -3 gpurun_out/r2c8_isp.err
B="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-extras"
timeout 600 $B > gpurun_out/r2c8_bench_default.json 2>> gpurun_out/r2c8_bench.err
S360_TMA=1 timeout 600 $B > gpurun_out/r2c8_bench_tma.json 2>> gpurun_out/r2c8_bench.err
timeout 600 $B --inflight 1 > gpurun_out/r2c8_bench_default_f1.json 2>> gpurun_out/r2c8_bench.err
S360_TMA=1 timeout 600 $B --inflight 1 > gpurun_out/r2c8_bench_tma_f1.json 2>> gpurun_out/r2c8_bench.err
SMALL="python bench.py --inflight 1 --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-extras"
timeout 600 ncu --set full --clock-control none -k regex:"k_blur15_rows" -s 120 -c 4 -f -o gpurun_out/r2c8_rows_manual $SMALL > gpurun_out/r2c8_ncu_a.log 2>&1
S360_TMA=1 timeout 600 ncu --set full --clock-control none -k regex:"k_blur15_rows" -s 120 -c 4 -f -o gpurun_out/r2c8_rows_tma $SMALL > gpurun_out/r2c8_ncu_b.log 2>&1
for n in manual tma; do ncu -i gpurun_out/r2c8_rows_$n.ncu-rep --page raw --csv > gpurun_out/r2c8_rows_${n}_raw.csv 2>/dev/null; done
rm -f gpurun_out/r2c8_rows_manual.ncu-rep gpurun_out/r2c8_rows_tma.ncu-rep
# source-level view of the sweep: launches 266/267 of a frame = finest pole level (forward, backward), 132/133 = finest side level
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_sweep" -s 266 -c 2 -f -o gpurun_out/r2c8_sweep_pole $SMALL > gpurun_out/r2c8_ncu_c.log 2>&1
ncu -i gpurun_out/r2c8_sweep_pole.ncu-rep --page raw --csv > gpurun_out/r2c8_sweep_pole_raw.csv 2>/dev/null
ncu -i gpurun_out/r2c8_sweep_pole.ncu-rep --page source --print-source sass --csv > gpurun_out/r2c8_sweep_pole_sass.csv 2>/dev/null
ncu -i gpurun_out/r2c8_sweep_pole.ncu-rep --page source --print-source cuda,sass --csv > gpurun_out/r2c8_sweep_pole_cuda.csv 2>/dev/null
rm -f gpurun_out/r2c8_sweep_pole.ncu-rep
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r2c8_bench*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        rows = [k for k in d.get("roofline_all", []) if k["kernel"].startswith("k_blur15_rows")]
        print(f, round(d["value"], 2), round(d["ms_per_step"], 1), d["config"]["frames_per_step"], round(d.get("stage_ms_single_frame", {}).get("total", 0), 1), d["output_crc32"]["graph_replay_frame0"], rows)
    except Exception as e:
        print(f, "unreadable", e)
PY
tail -5 gpurun_out/r2c8_bench.err
ls -la gpurun_out | head -40
