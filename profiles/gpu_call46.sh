#!/bin/bash
# round 2, call 46: MaxMatches truncation and the hit-explosion stress cases under both fronts of the scan
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_shard_maxmatches.py -m gpu -x -q -k "maxmatches or high_multiplicity" > gpurun_out/pytest_mm.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/pytest_mm.log
