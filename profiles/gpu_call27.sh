#!/bin/bash
# round 2, call 28: bucket-line prefetch through the TMA engine (cp.async.bulk.prefetch.L2, 128 B)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
O=gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "random_cases or golden or gendat_medium" > $O/pytest_front.log 2>&1; echo "pytest rc=$?"
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
  env "$@" timeout 400 python profiles/scale_step.py --scale $sc --steps 2 > $O/step_${sc}_$name.log 2>&1; echo "== scale $sc $name rc=$?"; show $O/step_${sc}_$name.log
}
run pf1 1.0 MSC_SCAN_PREFETCH=1
run pf0 1.0 MSC_SCAN_PREFETCH=0
run pf1_16 1.0 MSC_SCAN_PREFETCH=1 MSC_FRONT_PASS_MB=16
run pf1_64 1.0 MSC_SCAN_PREFETCH=1 MSC_FRONT_PASS_MB=64
run pf1 0.25 MSC_SCAN_PREFETCH=1
run pf1 0.125 MSC_SCAN_PREFETCH=1
