#!/bin/bash
# usage: scratch/prof2.sh <tag> [bench args]  -> gpurun_out/prof_<tag>.ncu-rep + launch list
tag=$1; shift
ncu --set full --clock-control none --import-source on -k regex:plan_kernel -s 1 -c 1 -o gpurun_out/prof_$tag python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-others --e2e-steps 1 "$@" > gpurun_out/ncu_$tag.log 2>&1
tail -1 gpurun_out/ncu_$tag.log | cut -c1-160
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$tag.csv python bench.py --steps 3 --warmup 1 --no-cpu-baseline --no-others --e2e-steps 1 "$@" > gpurun_out/launches_$tag.log 2>&1
tail -1 gpurun_out/launches_$tag.log | cut -c1-100
