#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -15 $O/pytest_gpu.log
timeout 300 python profiles/scale_step.py --scale 0.25 --steps 3 > $O/step_q_v2.log 2>&1; tail -1 $O/step_q_v2.log | cut -c300-1000
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $O/bench_s2_v2.log 2> $O/bench_s2_v2.err; echo "bench rc=$?"; tail -c 1500 $O/bench_s2_v2.log; tail -5 $O/bench_s2_v2.err
