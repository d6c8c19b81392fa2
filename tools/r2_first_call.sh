#!/bin/bash
# First GPU call after merging r2-prep (run from the repo root under gpurun, one GPU):
#   1. parity: the whole GPU tier on the merged tree (column-band sweep, temporal hand-over, launch checks)
#   2. A/B: bench (20 frames in flight and 1) with the default build (k_sweep2) and every library in build_variants/
# Outputs land in gpurun_out/. Build the row-band variant BEFORE the call (no nvcc time on the GPU box):
#   make -C surround360_b200/csrc                                    # default
#   tools/r2_first_call.sh --build-variants                          # build_variants/libs360b200_<variant>.so
set -u
cd "$(dirname "$0")/.."
if [ "${1:-}" = "--build-variants" ]; then
  # A/B libraries: name=flags (the default build is k_sweep2, 96 registers, bands over the shorter side)
  mkdir -p build_variants /tmp/s360_variant
  rm -rf /tmp/s360_variant/csrc && cp -r surround360_b200/csrc /tmp/s360_variant/
  for v in "sweep2_128regs=-DS360_SWEEP2_MINBLOCKS=4" "sweep2_80regs=-DS360_SWEEP2_MINBLOCKS=6" "sweep1_fast=-DS360_SWEEP_KERNEL=1" "sweep_generic=-DS360_SWEEP_KERNEL=0" "rowbands=-DS360_SWEEP_TRANSPOSE=0" "median8=-DS360_MEDIAN_ROWS=8"; do
    name=${v%%=*}; flags=${v#*=}
    (cd /tmp/s360_variant/csrc && mkdir -p build &&
     nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -fmad=false -std=c++17 -Xcompiler -fPIC $flags -c flow_kernels.cu -o build/flow_kernels_$name.o) || exit 1
    nvcc -gencode arch=compute_100a,code=sm_100a -shared -o build_variants/libs360b200_$name.so \
      surround360_b200/csrc/build/warp_kernels.o surround360_b200/csrc/build/finish_kernels.o /tmp/s360_variant/csrc/build/flow_kernels_$name.o \
      surround360_b200/csrc/build/pipeline.o surround360_b200/csrc/build/capi.o surround360_b200/csrc/build/rig.o -lcudart || exit 1
  done
  ls -la build_variants
  exit 0
fi
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/r2_tests.log
cat gpurun_out/r2_tests.log
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2_bench_default.json 2> gpurun_out/r2_bench.err
python bench.py --steps 2 --warmup 3 --inflight 1 --no-cpu-baseline --no-e2e > gpurun_out/r2_bench_default_f1.json 2>> gpurun_out/r2_bench.err
cp surround360_b200/libs360b200.so /tmp/libs360b200_default.so
for lib in build_variants/libs360b200_*.so; do
  [ -f "$lib" ] || continue
  name=$(basename "$lib" .so); name=${name#libs360b200_}
  cp "$lib" surround360_b200/libs360b200.so
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2_bench_$name.json 2>> gpurun_out/r2_bench.err
  python bench.py --steps 2 --warmup 3 --inflight 1 --no-cpu-baseline --no-e2e > gpurun_out/r2_bench_${name}_f1.json 2>> gpurun_out/r2_bench.err
done
cp /tmp/libs360b200_default.so surround360_b200/libs360b200.so
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r2_bench_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, d["value"], d["ms_per_step"], d["roofline"].get("ns_per_wavefront_step"))
    except Exception as e:
        print(f, "unreadable", e)
PY
