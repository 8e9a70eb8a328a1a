#!/bin/bash
# round 2, call 16: exact front (4^W-bit bitmap tested slice by slice, L2 eviction priorities) against the Bloom front at S2
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "random_cases or golden or gendat_medium" > $O/pytest_front.log 2>&1; echo "pytest rc=$?"
tail -3 $O/pytest_front.log
show() { python - "$1" <<'PY'
import json,sys
try:
    l=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    s=l["stage_ms"]; print("   step %.2f ms | " % l["ms_per_step"] + " ".join(f"{k[3:]}={v:.2f}" for k,v in s.items()))
    print("   ", {k: l["counts"][k] for k in ("bloom_bytes", "n_candidates", "bloom_pass", "n_matches")})
except Exception as e:
    print("   failed:", e)
PY
}
run() { # name, env..., then args after --
  name=$1; shift
  env "$@" timeout 400 python profiles/scale_step.py --scale 1.0 --steps 2 > $O/step_1.0_$name.log 2>&1; echo "== scale 1.0 $name rc=$?"; show $O/step_1.0_$name.log
}
run direct32 MSC_X=0
run direct64 MSC_FRONT_PASS_MB=64
run direct128 MSC_FRONT_PASS_MB=128
run direct16 MSC_FRONT_PASS_MB=16
run direct32_nostream MSC_SCAN_STREAM_TAB=0
run bloom_stream MSC_FRONT_DIRECT=0
run bloom_nostream MSC_FRONT_DIRECT=0 MSC_SCAN_STREAM_TAB=0
