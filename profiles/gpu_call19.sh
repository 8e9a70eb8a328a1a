#!/bin/bash
# round 2, call 19: exact front v4 (own kernel scan_direct.cuh: survivor queue that outlives the tile, no stage, slots reserved ahead of the look-ups): slice sweep at S2
# passes; last-use L2 hint on the record load; PDL only on small workloads): slice sweep at S2
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
for mb in 32 16 8 4 2 1; do run d4_$mb MSC_FRONT_PASS_MB=$mb; done
run d4_8_nostream MSC_FRONT_PASS_MB=8 MSC_SCAN_STREAM_TAB=0
run d4_16_nostream MSC_FRONT_PASS_MB=16 MSC_SCAN_STREAM_TAB=0
run bloom MSC_FRONT_DIRECT=0
run bloom_nostream MSC_FRONT_DIRECT=0 MSC_SCAN_STREAM_TAB=0
env MSC_X=1 timeout 400 python profiles/scale_step.py --scale 1.0 --steps 2 --window-width 20 > $O/step_1.0_w20_v8.log 2>&1; echo "== scale 1.0 W=20 rc=$?"; show $O/step_1.0_w20_v8.log
env MSC_X=1 timeout 400 python profiles/scale_step.py --scale 0.25 --steps 2 > $O/step_0.25_v8.log 2>&1; echo "== scale 0.25 rc=$?"; show $O/step_0.25_v8.log
