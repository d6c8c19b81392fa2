#!/bin/bash
# Profiling call for the next round (one GPU, run from the repo root under gpurun; ≈3-4 minutes of box time):
#   1. plain bench.py (default flags)                      -> gpurun_out/bench_r2.json      (never under ncu)
#   2. ncu launch list of one frame                        -> gpurun_out/launches_r2.csv
#   3. ncu --set full of the finest side level's kernels   -> gpurun_out/level_r2.ncu-rep + level_r2_raw.csv
#      (per level the solver launches: k_level_prep, k_blur15_rows, k_blur15_cols<0>, k_sweep2<+1>, k_median5,
#       k_sweep2<-1>, k_median5, k_blur15_rows, k_blur15_cols<1>, k_resize_cubic_f2 = 10 matches; the side batch has 31
#       levels, so skipping 300 matches lands on its finest level, which has no upsample: 9 launches)
#   4. the frame-level kernels once                        -> gpurun_out/frame_r2_raw.csv
# Summaries: python tools/ncu_table.py gpurun_out/level_r2_raw.csv ; python tools/ncu_table.py --launches gpurun_out/launches_r2.csv
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python bench.py > gpurun_out/bench_r2.json 2> gpurun_out/bench_r2.err
SMALL="python bench.py --inflight 1 --steps 1 --warmup 3 --no-cpu-baseline --no-e2e"
ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 2700 -c 1100 --csv --log-file gpurun_out/launches_r2.csv $SMALL > gpurun_out/ncu_r2_1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_level_prep|k_blur15|k_sweep2|k_median5|k_resize_cubic_f2" -s 300 -c 9 -f -o gpurun_out/level_r2 $SMALL > gpurun_out/ncu_r2_2.log 2>&1
ncu -i gpurun_out/level_r2.ncu-rep --page raw --csv > gpurun_out/level_r2_raw.csv 2>/dev/null
ncu --set full --clock-control none -k regex:"k_project_spherical|k_novel_view|k_pole_warp|k_flatten_poles|k_resize_cubic8|k_blur3_f2|k_resize_linear|k_flow_downscale|k_gray_alpha_blur5|k_erode|k_blur8" -c 40 -f -o gpurun_out/frame_r2 $SMALL > gpurun_out/ncu_r2_3.log 2>&1
ncu -i gpurun_out/frame_r2.ncu-rep --page raw --csv > gpurun_out/frame_r2_raw.csv 2>/dev/null
ls -la gpurun_out | tail -8
