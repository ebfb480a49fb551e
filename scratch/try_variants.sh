#!/bin/bash
# usage: try_variants.sh lib1.so lib2.so ...   (each copied over the in-tree lib, bench run, restored)
cp quadrupedal_foothold_planner_b200/libfhp.so /tmp/libfhp_orig.so
for so in "$@"; do
  cp "$so" quadrupedal_foothold_planner_b200/libfhp.so
  touch quadrupedal_foothold_planner_b200/libfhp.so
  echo "== $so"
  python bench.py --steps 10 --warmup 3 --no-cpu-baseline --e2e-steps 2 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('value %.4g ms %.3f e2e %.4g frac %.3f' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac']))"
done
cp /tmp/libfhp_orig.so quadrupedal_foothold_planner_b200/libfhp.so
