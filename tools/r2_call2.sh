#!/bin/bash
# GPU call 2 of round 2 (one GPU): GPU tier incl. the full-size parity tests, then a per-kernel counter pass over one frame
# (time, instructions, DRAM bytes, issue utilisation) and a source-level capture of the finest side-level sweep.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r2c2_tests.log
cat gpurun_out/r2c2_tests.log
SMALL="python bench.py --inflight 1 --steps 1 --warmup 3 --no-cpu-baseline --no-e2e"
M=gpu__time_duration.sum,smsp__inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,launch__grid_size,launch__block_size,launch__registers_per_thread,lts__t_bytes.sum
timeout 900 ncu --metrics $M --clock-control none --launch-skip 2700 -c 1100 --csv --log-file gpurun_out/r2c2_frame_counters.csv $SMALL > gpurun_out/r2c2_ncu1.log 2>&1
# sweeps per frame: 31 side levels x 2 + 36 pole levels x 2 = 134; finest side level of the 2nd frame = #194,195, finest pole level = #266,267
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_sweep" -s 194 -c 2 -f -o gpurun_out/r2c2_sweep_side $SMALL > gpurun_out/r2c2_ncu2.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_sweep" -s 266 -c 2 -f -o gpurun_out/r2c2_sweep_pole $SMALL > gpurun_out/r2c2_ncu3.log 2>&1
for n in side pole; do
  ncu -i gpurun_out/r2c2_sweep_$n.ncu-rep --page raw --csv > gpurun_out/r2c2_sweep_${n}_raw.csv 2>/dev/null
  ncu -i gpurun_out/r2c2_sweep_$n.ncu-rep --page source --csv > gpurun_out/r2c2_sweep_${n}_source.csv 2>/dev/null
done
ls -la gpurun_out | tail -8
