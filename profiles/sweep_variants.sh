#!/bin/bash
# A/B of compile-time scan variants (built into .variants/ by hand): batch size x CTAs per SM.
cp muscato_b200/libmuscato_b200.so /tmp/lib_orig.so
for v in .variants/lib_*.so; do
  cp $v muscato_b200/libmuscato_b200.so
  python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readlines()[-1]); print('$v', 'scan_ms', round(d['roofline']['kernel_ms'],4), 'step_ms', round(d['ms_per_step'],4), d['stage_ms_per_step'])"
done
cp /tmp/lib_orig.so muscato_b200/libmuscato_b200.so
