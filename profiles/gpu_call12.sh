#!/bin/bash
# round 2, call 12: random-line HBM microbenchmark; scan v7 (target offsets of a tile in registers, arithmetic Bloom masks)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
O=gpurun_out
timeout 300 ./profiles/microbench/random_lines > $O/random_lines_b200.txt 2>&1; echo "microbench rc=$?"; cat $O/random_lines_b200.txt
timeout 2400 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -3 $O/pytest_gpu.log
show() { python - "$1" <<'PY'
import json,sys
l=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
s=l["stage_ms"]; print("   step %.2f ms | " % l["ms_per_step"] + " ".join(f"{k[3:]}={v:.2f}" for k,v in s.items()))
PY
}
timeout 400 python profiles/scale_step.py --scale 0.25 --steps 2 > $O/step_0.25_v7.log 2>&1; echo "== scale 0.25 v7 rc=$?"; show $O/step_0.25_v7.log
for am in 1 0; do
MSC_SCAN_ALU_MASKS=$am timeout 400 python profiles/scale_step.py --scale 1.0 --steps 2 > $O/step_1.0_v7_am$am.log 2>&1; echo "== scale 1.0 v7 alu_masks $am rc=$?"; show $O/step_1.0_v7_am$am.log
MSC_SCAN_ALU_MASKS=$am timeout 400 python profiles/scale_step.py --scale 1.0 --steps 2 --window-width 20 > $O/step_1.0_w20_v7_am$am.log 2>&1; echo "== scale 1.0 W=20 alu_masks $am rc=$?"; show $O/step_1.0_w20_v7_am$am.log
done
