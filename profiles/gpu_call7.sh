#!/bin/bash
# round 2, call 7: scan v3 (group record fetched at hit time, per-candidate records written by the scan, line-granular
# Bloom blocks): parity suite, then scan time vs Bloom block size / minimiser length at S2 and S2/4
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
O=gpurun_out
timeout 2400 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -5 $O/pytest_gpu.log
show() { python - "$1" <<'PY'
import json,sys
l=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
s=l["stage_ms"]; print("   step %.2f ms | " % l["ms_per_step"] + " ".join(f"{k[3:]}={v:.2f}" for k,v in s.items()))
PY
}
for sc in 0.25 1.0; do
for blk in 4 2; do for m in 10 11 12 13; do
  MSC_BLOOM_LG_BLK=$blk MSC_MINIMIZER_M=$m timeout 400 python profiles/scale_step.py --scale $sc --steps 2 > $O/sweep_${sc}_${blk}_${m}.log 2>&1
  echo "== scale $sc lg_blk $blk m $m rc=$?"; show $O/sweep_${sc}_${blk}_${m}.log
done; done; done
MSC_TRACE=1 timeout 600 python profiles/scale_step.py --scale 1.0 --steps 1 > $O/trace_s2_v3.log 2>&1; echo "trace full rc=$?"
