#!/bin/bash
# round 2, call 9: scan v4 (L2 prefetch of the tile's bucket lines before the drain, group records of a round fetched together)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
O=gpurun_out
timeout 2400 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -3 $O/pytest_gpu.log
show() { python - "$1" <<'PY'
import json,sys
l=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
s=l["stage_ms"]; print("   step %.2f ms | " % l["ms_per_step"] + " ".join(f"{k[3:]}={v:.2f}" for k,v in s.items()))
PY
}
for sc in 0.25 1.0; do
  for pf in 1 0; do
  MSC_SCAN_PREFETCH=$pf timeout 400 python profiles/scale_step.py --scale $sc --steps 2 > $O/step_${sc}_pf$pf.log 2>&1; echo "== scale $sc prefetch $pf rc=$?"; show $O/step_${sc}_pf$pf.log
  done
  MSC_MINIMIZER_M=11 timeout 400 python profiles/scale_step.py --scale $sc --steps 2 > $O/step_${sc}_m11.log 2>&1; echo "== scale $sc m=11 rc=$?"; show $O/step_${sc}_m11.log
done
timeout 1200 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 120 --csv --log-file $O/launches_r02_s2_v4.csv \
    python profiles/scale_step.py --scale 1.0 --steps 1 > $O/ncu_l4.log 2>&1
echo "launchlist rc=$?"
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:"scan_targets_kernel" -s 2 -c 1 \
    -o $O/prof_r02_s2_scan_v4 python profiles/scale_step.py --scale 1.0 --steps 1 > $O/ncu_f4.log 2>&1
echo "ncufull rc=$?"
