#!/bin/bash
# round 2, call 36: the window pass writes the item records, the scatter only partitions them
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
O=gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_readprep.py -m gpu -x -q > $O/pytest_front.log 2>&1; echo "pytest rc=$?"
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
run() { # name scale env...
  name=$1; sc=$2; shift; shift
  env "$@" timeout 400 python profiles/scale_step.py --scale $sc --steps 2 $EXTRA > $O/step_${sc}_$name.log 2>&1; echo "== scale $sc $name rc=$?"; show $O/step_${sc}_$name.log
}
run b3 1.0 MSC_X=1
run b3 0.25 MSC_X=1
MSC_TRACE=1 timeout 400 python profiles/scale_step.py --scale 1.0 --steps 1 > $O/trace_s2_b3.log 2>&1; python profiles/trace_names.py $O/trace_s2_b3.log 1 | grep -E "build_|table_clear|pack_" 
