#!/bin/bash
# round 2, call 26: what bounds the exact-front scan -- the same step without candidate records / without table look-ups
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
O=gpurun_out
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
run dbg0 1.0 MSC_SCAN_DEBUG=0
run dbg1 1.0 MSC_SCAN_DEBUG=1
run dbg2 1.0 MSC_SCAN_DEBUG=2
run dbg2_64 1.0 MSC_SCAN_DEBUG=2 MSC_FRONT_PASS_MB=64
run dbg2_16 1.0 MSC_SCAN_DEBUG=2 MSC_FRONT_PASS_MB=16
run dbg2_128 1.0 MSC_SCAN_DEBUG=2 MSC_FRONT_PASS_MB=128
