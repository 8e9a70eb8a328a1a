#!/bin/bash
# round 2, call 11: scan v6 (offset-slice loads pipelined, no X-summary look-ups without X), scatter with a 5120-record
# stage, table partition size sweep; ncu --set full of the scan at S2 (W=15 and W=20)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
O=gpurun_out
timeout 2400 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -3 $O/pytest_gpu.log
show() { python - "$1" <<'PY'
import json,sys
l=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
s=l["stage_ms"]; print("   step %.2f ms | " % l["ms_per_step"] + " ".join(f"{k[3:]}={v:.2f}" for k,v in s.items()))
PY
}
timeout 400 python profiles/scale_step.py --scale 0.25 --steps 2 > $O/step_0.25_v6.log 2>&1; echo "== scale 0.25 v6 rc=$?"; show $O/step_0.25_v6.log
for bpp in 17 18 19; do
  MSC_TRACE=1 MSC_TABLE_LG_BPP=$bpp timeout 400 python profiles/scale_step.py --scale 1.0 --steps 1 > $O/step_1.0_v6_bpp$bpp.log 2>&1; echo "== scale 1.0 v6 lg_bpp $bpp rc=$?"; show $O/step_1.0_v6_bpp$bpp.log
  python profiles/trace_names.py $O/step_1.0_v6_bpp$bpp.log 1 | grep -E "build_|table_clear"
done
timeout 400 python profiles/scale_step.py --scale 1.0 --steps 2 --window-width 20 > $O/step_1.0_w20_v6.log 2>&1; echo "== scale 1.0 W=20 rc=$?"; show $O/step_1.0_w20_v6.log
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:"scan_targets_kernel" -s 2 -c 1 \
    -o $O/prof_r02_s2_scan_v6 python profiles/scale_step.py --scale 1.0 --steps 1 > $O/ncu_f6.log 2>&1
echo "ncufull rc=$?"
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:"scan_targets_kernel" -s 2 -c 1 \
    -o $O/prof_r02_s2_scan_w20_v6 python profiles/scale_step.py --scale 1.0 --steps 1 --window-width 20 > $O/ncu_f6w.log 2>&1
echo "ncufull w20 rc=$?"
