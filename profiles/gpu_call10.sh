#!/bin/bash
# round 2, call 10: scan v5 (per-tile target-offset slice in shared memory, flush per tile, no prefetch)
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
  timeout 400 python profiles/scale_step.py --scale $sc --steps 2 > $O/step_${sc}_v5.log 2>&1; echo "== scale $sc v5 rc=$?"; show $O/step_${sc}_v5.log
  MSC_BLOOM_LG_BLK=2 timeout 400 python profiles/scale_step.py --scale $sc --steps 2 > $O/step_${sc}_v5b2.log 2>&1; echo "== scale $sc v5 lg_blk 2 rc=$?"; show $O/step_${sc}_v5b2.log
done
timeout 400 python profiles/scale_step.py --scale 1.0 --steps 2 --window-width 20 > $O/step_1.0_w20_v5.log 2>&1; echo "== scale 1.0 W=20 rc=$?"; show $O/step_1.0_w20_v5.log
timeout 1200 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 120 --csv --log-file $O/launches_r02_s2_v5.csv \
    python profiles/scale_step.py --scale 1.0 --steps 1 > $O/ncu_l5.log 2>&1
echo "launchlist rc=$?"
timeout 1200 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 120 --csv --log-file $O/launches_r02_s2_w20_v5.csv \
    python profiles/scale_step.py --scale 1.0 --steps 1 --window-width 20 > $O/ncu_l5w.log 2>&1
echo "launchlist w20 rc=$?"
