#!/bin/bash
# round 2, call 39: combine scatter with one per-read cursor
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
O=gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -m gpu -x -q > $O/pytest_front.log 2>&1; echo "pytest rc=$?"
tail -3 $O/pytest_front.log
show() { python - "$1" <<'PY'
import json,sys
try:
    l=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    s=l["stage_ms"]; print("   step %.2f ms | " % l["ms_per_step"] + " ".join(f"{k[3:]}={v:.2f}" for k,v in s.items()))
except Exception as e:
    print("   failed:", e)
PY
}
timeout 400 python profiles/scale_step.py --scale 1.0 --steps 2 > $O/step_1.0_c1.log 2>&1; echo "== scale 1.0 rc=$?"; show $O/step_1.0_c1.log
timeout 400 python profiles/scale_step.py --scale 1.0 --steps 2 --window-width 20 > $O/step_1.0_w20_c1.log 2>&1; echo "== scale 1.0 W=20 rc=$?"; show $O/step_1.0_w20_c1.log
