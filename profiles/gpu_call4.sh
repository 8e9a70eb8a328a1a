#!/bin/bash
# round 2 (re-entry), call 4: state of HEAD on the GPU: parity tests, the S2/4 step, the S2 bench
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
O=gpurun_out
(free -g; nproc; nvidia-smi --query-gpu=name,memory.total --format=csv; df -h /dev/shm /tmp | tail -2) > $O/box.txt 2>&1
timeout 1800 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $O/box.txt
tail -15 $O/pytest_gpu.log
timeout 300 python profiles/scale_step.py --scale 0.25 --steps 3 > $O/step_q_v2.log 2>&1; echo "step rc=$?"; tail -1 $O/step_q_v2.log | cut -c300-1100
timeout 900 python bench.py --steps 3 --warmup 3 > $O/bench_s2_v2.log 2> $O/bench_s2_v2.err; echo "bench rc=$?" | tee -a $O/box.txt; tail -c 3000 $O/bench_s2_v2.log; tail -5 $O/bench_s2_v2.err
cat $O/box.txt
