#!/bin/bash
# Sweep the minimiser length of the Bloom addressing (MSC_MINIMIZER_M) on the bench workload.
for m in 9 10 11 12 13; do
  MSC_MINIMIZER_M=$m python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readlines()[-1]); print('m=$m', 'scan_ms', round(d['roofline']['kernel_ms'],4), 'bloom_pass', d['counts']['bloom_pass'], 'step_ms', round(d['ms_per_step'],4))"
done
