#!/bin/bash
# round 2, call 1: hygiene + new parity tests, the S2 bench with the round-1 kernels (baseline at scale),
# launch list + ncu --set full of the four large-scale kernels at 1/4 of S2
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
O=gpurun_out
(free -g; nproc; nvidia-smi --query-gpu=name,memory.total --format=csv; df -h /dev/shm /tmp | tail -2) > $O/box.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/box.txt
timeout 900 python bench.py --steps 3 --warmup 3 > $O/bench_s2_base.log 2> $O/bench_s2_base.err; echo "bench rc=$?" >> $O/box.txt
timeout 600 python profiles/scale_step.py --scale 0.25 --steps 2 > $O/step_q.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_r02_base_s2q.csv \
    python profiles/scale_step.py --scale 0.25 --steps 2 > $O/ncu_l.log 2>&1
echo "launchlist rc=$?" >> $O/box.txt
timeout 1500 ncu --set full --clock-control none --import-source on \
    -k regex:"scan_targets_kernel|build_keys_insert_kernel|cand_prepare_kernel|confirm_pairs_kernel" -c 12 \
    -o $O/prof_r02_base_s2q python profiles/scale_step.py --scale 0.25 --steps 1 > $O/ncu_f.log 2>&1
echo "ncufull rc=$?" >> $O/box.txt
tail -3 $O/pytest_gpu.log; tail -c 600 $O/bench_s2_base.log; cat $O/box.txt
