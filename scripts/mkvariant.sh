#!/bin/bash
# scripts/mkvariant.sh NAME [-DFLAG ...]: builds csrc/variants/lib_NAME.so with extra defines (A/B on one box)
cd "$(dirname "$0")/../lowcost3dreconstruction_b200/csrc" || exit 1
name=$1; shift
mkdir -p variants
/usr/local/cuda/bin/nvcc -O3 -std=c++17 -lineinfo -fmad=false -gencode arch=compute_100a,code=sm_100a \
  -Xcompiler -fPIC,-Wall,-Wno-unused-function -Xptxas -v "$@" -shared -o variants/lib_$name.so capi.cu 2> variants/build_$name.log \
  || { cat variants/build_$name.log; exit 1; }
grep -A1 "icp_iteration_kernelILi1ELb0" variants/build_$name.log | grep -E "registers|spill" | head -3
