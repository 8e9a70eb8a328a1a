#!/bin/bash
# round 2, call 13 (2 GPUs): the N>1 bench path (rank grids, NCCL MIN all-reduce, verification against the unsharded
# run), the two NCCL tests, S3 / S4 bench modes at reduced scale
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
O=gpurun_out
nvidia-smi --query-gpu=index,name,memory.total --format=csv > $O/box2.txt; free -g >> $O/box2.txt; nproc >> $O/box2.txt
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
brief() { python - "$1" <<'PY'
import json,sys
try:
    l=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print("   value %.3e %s | %.2f ms/step | e2e %.3e (%.2f ms) | grid %s | verify %s (%s matches) | launches %s" % (l["value"], l["unit"], l["ms_per_step"], l["e2e"]["value"], l["e2e"]["ms_per_step"], l["setup"]["grid"], l.get("verify_sharded_equals_unsharded"), l.get("verify_matches"), l["gpu_launches"]))
    print("   stage", {k[3:]: round(v,2) for k,v in l["stage_ms_per_step"].items()}, "roofline frac %.3f" % l["roofline"]["frac"])
except Exception as e:
    print("   no JSON line:", e)
PY
}
timeout 600 $TR bench.py --gpus 2 --steps 3 --warmup 3 --scale 0.1 > $O/bench_n2_s2_010.log 2> $O/bench_n2_s2_010.err; echo "== N=2 S2/10 auto grid rc=$?"; brief $O/bench_n2_s2_010.log; tail -3 $O/bench_n2_s2_010.err
timeout 600 $TR bench.py --gpus 2 --steps 3 --warmup 3 --scale 0.1 --grid 1x2 > $O/bench_n2_s2_010_1x2.log 2> $O/bench_n2_s2_010_1x2.err; echo "== N=2 S2/10 1x2 rc=$?"; brief $O/bench_n2_s2_010_1x2.log; tail -3 $O/bench_n2_s2_010_1x2.err
timeout 900 $TR bench.py --gpus 2 --steps 5 --warmup 3 > $O/bench_n2_s2.log 2> $O/bench_n2_s2.err; echo "== N=2 S2 full rc=$?"; brief $O/bench_n2_s2.log; tail -3 $O/bench_n2_s2.err
timeout 900 $TR bench.py --gpus 2 --steps 5 --warmup 3 --grid 1x2 --no-verify > $O/bench_n2_s2_1x2.log 2> $O/bench_n2_s2_1x2.err; echo "== N=2 S2 full 1x2 rc=$?"; brief $O/bench_n2_s2_1x2.log; tail -3 $O/bench_n2_s2_1x2.err
timeout 600 $TR bench.py --gpus 2 --config s3 --scale 0.02 --steps 2 --warmup 3 > $O/bench_n2_s3_002.log 2> $O/bench_n2_s3_002.err; echo "== N=2 S3/50 rc=$?"; brief $O/bench_n2_s3_002.log; tail -3 $O/bench_n2_s3_002.err
timeout 900 python -m pytest tests -m gpu -q -k "nccl" > $O/pytest_nccl.log 2>&1; echo "pytest nccl rc=$?"; tail -3 $O/pytest_nccl.log
timeout 400 python bench.py --config s4 --steps 3 --warmup 3 --no-cpu-baseline > $O/bench_s4.log 2> $O/bench_s4.err; echo "== S4 rc=$?"; brief $O/bench_s4.log; tail -3 $O/bench_s4.err
timeout 400 python bench.py --config s4 --max-matches 1000 --steps 3 --warmup 3 --no-cpu-baseline > $O/bench_s4_mm1000.log 2> $O/bench_s4_mm1000.err; echo "== S4 MaxMatches 1000 rc=$?"; brief $O/bench_s4_mm1000.log; tail -3 $O/bench_s4_mm1000.err
