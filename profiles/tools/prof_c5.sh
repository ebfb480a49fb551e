#!/bin/bash
ncu --set full --clock-control none --import-source on -k regex:plan_kernel -s 1 -c 1 -o gpurun_out/prof_$1 python bench.py --workload c5 --steps 2 --warmup 1 --no-cpu-baseline --e2e-steps 1 > gpurun_out/ncu_$1.log 2>&1
tail -1 gpurun_out/ncu_$1.log | cut -c1-120
