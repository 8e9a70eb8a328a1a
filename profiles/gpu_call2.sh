#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
for g in 128 64 32; do
  MSC_L2_FETCH=$g timeout 300 python profiles/scale_step.py --scale 0.25 --steps 3 > $O/l2fetch_$g.log 2>&1
  echo "== L2 fetch $g"; tail -1 $O/l2fetch_$g.log | cut -c300-900
done
timeout 300 python profiles/scale_step.py --scale 0.25 --steps 3 > $O/l2fetch_default.log 2>&1; echo "== default"; tail -1 $O/l2fetch_default.log | cut -c300-900
python - <<'PY'
import ctypes
rt = ctypes.CDLL("libcudart.so.12") if False else None
PY
MSC_L2_FETCH=32 timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
   -k regex:"scan_targets_kernel|build_keys_insert_kernel|cand_prepare_kernel|confirm_pairs_kernel" -c 8 --csv --log-file $O/l2fetch32_dram.csv \
   python profiles/scale_step.py --scale 0.25 --steps 1 > $O/l2fetch32_ncu.log 2>&1
echo "ncu rc=$?"
