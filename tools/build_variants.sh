#!/bin/bash
# A/B builds of the sweep kernel: build_variants/libs360b200_<name>.so = the product library with flow_kernels.cu compiled under
# the given macros (the other objects are the default build's). build_variants/ is git-ignored but travels to the GPU box.
#   tools/build_variants.sh name1 "-DMACRO=..." name2 "-D..." ...
set -e
cd "$(dirname "$0")/.."
C=surround360_b200/csrc
make -C $C -j8 -s
mkdir -p build_variants
NVFLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -fmad=false -std=c++17 -Xcompiler -fPIC"
OTHERS=$(ls $C/build/*.o | grep -v flow_kernels.o)
while [ $# -ge 2 ]; do
  name=$1; defs=$2; shift 2
  ( nvcc $NVFLAGS $defs -Xptxas -v -c $C/flow_kernels.cu -o build_variants/flow_$name.o 2> build_variants/ptxas_$name.log
    nvcc -gencode arch=compute_100a,code=sm_100a -shared -o build_variants/libs360b200_$name.so build_variants/flow_$name.o $OTHERS -lcudart
    echo "$name: $(grep -A1 'k_sweepILi1ELb0' build_variants/ptxas_$name.log | grep -o 'Used [0-9]* registers' | head -1)" ) &
done
wait
