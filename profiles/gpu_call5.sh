#!/bin/bash
# round 2, call 5: full GPU test suite; per-launch device times of one S2/4 and one S2 step (MSC_TRACE);
# launch list + ncu --set full of the large-scale kernels at S2/4 (table 4.3 GB, Bloom 128 MiB: beyond the L2)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
O=gpurun_out
timeout 2400 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -5 $O/pytest_gpu.log
MSC_TRACE=1 timeout 300 python profiles/scale_step.py --scale 0.25 --steps 1 > $O/trace_s2q.log 2>&1; echo "trace q rc=$?"
MSC_TRACE=1 timeout 600 python profiles/scale_step.py --scale 1.0 --steps 1 > $O/trace_s2.log 2>&1; echo "trace full rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv --log-file $O/launches_r02_s2q.csv \
    python profiles/scale_step.py --scale 0.25 --steps 1 > $O/ncu_l.log 2>&1
echo "launchlist rc=$?"
timeout 1500 ncu --set full --clock-control none --import-source on \
    -k regex:"scan_targets_kernel|build_insert_kernel|build_scatter_kernel|build_windows_kernel|cand_prepare_kernel|confirm_pairs_kernel" -c 12 \
    -o $O/prof_r02_s2q python profiles/scale_step.py --scale 0.25 --steps 1 > $O/ncu_f.log 2>&1
echo "ncufull rc=$?"
ls -la $O
