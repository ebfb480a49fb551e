#!/bin/bash
# usage: scratch/mk.sh <name> [extra nvcc flags...]  -> scratch/libs/libfhp_<name>.so (headline kernel variant only)
mkdir -p scratch/libs
name=$1; shift
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -fmad=false -Xcompiler -fPIC,-ffp-contract=off \
  -shared -cudart static -DFHP_BENCH_BUILD "$@" -Xptxas -v \
  quadrupedal_foothold_planner_b200/csrc/fhp_kernels.cu quadrupedal_foothold_planner_b200/csrc/fhp_capi.cu \
  -o scratch/libs/libfhp_$name.so -ldl 2>&1 | grep -A3 "plan_kernelILb0ELi2ELb0ELb0ELb0E" | grep -E "stack|Used" | tr '\n' ' '
echo " -> $name"
