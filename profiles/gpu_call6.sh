#!/bin/bash
# round 2, call 6: GPU test suite; launch list (time + DRAM bytes) of one FULL-size S2 step; ncu --set full of the
# scan kernel at full size
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
O=gpurun_out
timeout 2400 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -5 $O/pytest_gpu.log
timeout 1200 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 120 --csv --log-file $O/launches_r02_s2.csv \
    python profiles/scale_step.py --scale 1.0 --steps 1 > $O/ncu_l1.log 2>&1
echo "launchlist rc=$?"
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:"scan_targets_kernel" -s 2 -c 1 \
    -o $O/prof_r02_s2_scan python profiles/scale_step.py --scale 1.0 --steps 1 > $O/ncu_f1.log 2>&1
echo "ncufull rc=$?"
ls -la $O
