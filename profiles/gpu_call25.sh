#!/bin/bash
# round 2, call 25: exact front v6 (four tile buffers per warp, target offsets requested one tile ahead)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
O=gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_s2.py -m gpu -x -q > $O/pytest_front.log 2>&1; echo "pytest rc=$?"
tail -3 $O/pytest_front.log
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
for mb in 32 16 8 4; do run d6_$mb 1.0 MSC_FRONT_PASS_MB=$mb; done
run d6 0.25 MSC_X=1
for mb in 32 16 8; do run d6_$mb 0.125 MSC_FRONT_PASS_MB=$mb; done
run d6 0.05 MSC_X=1
