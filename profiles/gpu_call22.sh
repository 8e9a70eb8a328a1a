#!/bin/bash
# round 2, call 22: where the exact front starts to pay (keys per context) and its slice size at low hit density
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
O=gpurun_out
show() { python - "$1" <<'PY'
import json,sys
try:
    l=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    s=l["stage_ms"]; print("   step %.2f ms | " % l["ms_per_step"] + " ".join(f"{k[3:]}={v:.2f}" for k,v in s.items()))
    print("   ", {k: l["counts"][k] for k in ("n_keys", "target_bases", "bloom_bytes", "n_candidates", "bloom_pass", "n_matches")})
except Exception as e:
    print("   failed:", e)
PY
}
run() { # name scale env...
  name=$1; sc=$2; shift; shift
  env "$@" timeout 400 python profiles/scale_step.py --scale $sc --steps 2 > $O/step_${sc}_$name.log 2>&1; echo "== scale $sc $name rc=$?"; show $O/step_${sc}_$name.log
}
for sc in 0.125 0.05 0.02; do
  run bloom $sc MSC_FRONT_DIRECT=0
  for mb in 32 64 128; do run direct$mb $sc MSC_FRONT_DIRECT=1 MSC_FRONT_PASS_MB=$mb; done
done
