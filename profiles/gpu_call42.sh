#!/bin/bash
# round 2, call 42 (8 GPUs): end-of-round kernels -- S2 strong scaling at N=8 (grid chosen by the cost model, verified
# against the unsharded run; 8x1 for comparison) and configs[3] (S3) on 8 GPUs
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
O=gpurun_out
brief() { python - "$1" <<'PY'
import json,sys
try:
    l=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print("   value %.3e %s | %.2f ms/step | e2e %.3e (%.2f ms) | grid %s numa %s | verify %s (%s matches) | launches %s | hbm %.1f GB" % (l["value"], l["unit"], l["ms_per_step"], l["e2e"]["value"], l["e2e"]["ms_per_step"], l["setup"]["grid"], l["setup"].get("numa_node_rank0"), l.get("verify_sharded_equals_unsharded"), l.get("verify_matches"), l["gpu_launches"], l["setup"]["hbm_used_bytes_max_rank"]/1e9))
    print("   stage", {k[3:]: round(v,2) for k,v in l["stage_ms_per_step"].items()}, "roofline frac %.3f" % l["roofline"]["frac"]); print("   ", l["counts"])
except Exception as e:
    print("   no JSON line:", e)
PY
}
tr() { echo "python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port $2"; }
timeout 600 $(tr 8 29531) bench.py --gpus 8 --steps 5 --warmup 3 > $O/bench_n8_s2_end.log 2> $O/bench_n8_s2_end.err; echo "== N=8 S2 rc=$?"; brief $O/bench_n8_s2_end.log; tail -2 $O/bench_n8_s2_end.err
timeout 400 $(tr 8 29532) bench.py --gpus 8 --steps 5 --warmup 3 --grid 8x1 --no-verify > $O/bench_n8_s2_8x1_end.log 2> $O/bench_n8_s2_8x1_end.err; echo "== N=8 S2 8x1 rc=$?"; brief $O/bench_n8_s2_8x1_end.log; tail -2 $O/bench_n8_s2_8x1_end.err
timeout 600 $(tr 8 29533) bench.py --gpus 8 --config s3 --steps 2 --warmup 3 --no-verify > $O/bench_n8_s3_end.log 2> $O/bench_n8_s3_end.err; echo "== N=8 S3 rc=$?"; brief $O/bench_n8_s3_end.log; tail -2 $O/bench_n8_s3_end.err
