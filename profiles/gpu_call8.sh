#!/bin/bash
# round 2, call 8: sketch pre-filter in confirm (parity + time), ncu --set full of scan v3 and confirm at S2/4
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
O=gpurun_out
timeout 2400 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -5 $O/pytest_gpu.log
show() { python - "$1" <<'PY'
import json,sys
l=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
s=l["stage_ms"]; print("   step %.2f ms | " % l["ms_per_step"] + " ".join(f"{k[3:]}={v:.2f}" for k,v in s.items())); print("   ", l["counts"])
PY
}
for sc in 0.25 1.0; do
  timeout 400 python profiles/scale_step.py --scale $sc --steps 2 > $O/step_${sc}_sk.log 2>&1; echo "== scale $sc sketch on rc=$?"; show $O/step_${sc}_sk.log
  MSC_SKETCH=0 timeout 400 python profiles/scale_step.py --scale $sc --steps 2 > $O/step_${sc}_nosk.log 2>&1; echo "== scale $sc sketch off rc=$?"; show $O/step_${sc}_nosk.log
done
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:"scan_targets_kernel|confirm_pairs_kernel" -s 4 -c 2 \
    -o $O/prof_r02_s2q_v3 python profiles/scale_step.py --scale 0.25 --steps 1 > $O/ncu_f2.log 2>&1
echo "ncufull rc=$?"
