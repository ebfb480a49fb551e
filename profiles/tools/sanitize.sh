#!/bin/bash
# compute-sanitizer over a parity subset (mask staging = the default path, plus one tma/ldg case each)
S=/usr/local/cuda/bin/compute-sanitizer
K='rough_map_trot or streamed or compact_response or edges_border or heading_rough or heading_border or adopted or c1_flat or wide_foot or maps_smaller'
for tool in ${TOOLS:-memcheck racecheck synccheck}; do
  timeout 1500 $S --tool $tool --error-exitcode 7 python -m pytest tests/test_gpu_parity.py tests/test_gpu_heading.py -m gpu -q -x -k "($K) and (mask or ldg or tma)" > gpurun_out/sanitizer_$tool.log 2>&1
  echo "$tool rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed" gpurun_out/sanitizer_$tool.log | tail -3
done
