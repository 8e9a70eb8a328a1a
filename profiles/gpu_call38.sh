#!/bin/bash
# round 2, call 38: exact-front scan with 2 CTAs per SM (is it bound by the number of warps?)
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
cp muscato_b200/libmuscato_b200.so /tmp/lib_base.so
for v in base c2; do
  if [ $v = base ]; then cp /tmp/lib_base.so muscato_b200/libmuscato_b200.so; else cp muscato_b200/libmuscato_b200_$v.so muscato_b200/libmuscato_b200.so; fi
  timeout 400 python profiles/scale_step.py --scale 1.0 --steps 2 > $O/step_1.0_$v.log 2>&1; echo "== scale 1.0 $v rc=$?"; show $O/step_1.0_$v.log
done
cp /tmp/lib_base.so muscato_b200/libmuscato_b200.so
