#!/bin/bash
# round 2, call 21 (8 GPUs): strong scaling of the default bench (S2) at N=8 and N=4 with the sharded == unsharded
# check, and BASELINE configs[3] (S3: 3e8 x 150 bp reads vs 5e9 bases) on 8 GPUs
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
O=gpurun_out
nvidia-smi --query-gpu=index,name,memory.total --format=csv > $O/box8.txt; free -g >> $O/box8.txt; nproc >> $O/box8.txt
brief() { python - "$1" <<'PY'
import json,sys
try:
    l=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print("   value %.3e %s | %.2f ms/step | e2e %.3e (%.2f ms) | grid %s | verify %s (%s matches) | launches %s | hbm %.1f GB" % (l["value"], l["unit"], l["ms_per_step"], l["e2e"]["value"], l["e2e"]["ms_per_step"], l["setup"]["grid"], l.get("verify_sharded_equals_unsharded"), l.get("verify_matches"), l["gpu_launches"], l["setup"]["hbm_used_bytes_max_rank"]/1e9))
    print("   stage", {k[3:]: round(v,2) for k,v in l["stage_ms_per_step"].items()}, "roofline frac %.3f" % l["roofline"]["frac"]); print("   ", l["counts"])
except Exception as e:
    print("   no JSON line:", e)
PY
}
tr() { echo "python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port $2"; }
timeout 700 $(tr 8 29521) bench.py --gpus 8 --steps 5 --warmup 3 > $O/bench_n8_s2.log 2> $O/bench_n8_s2.err; echo "== N=8 S2 rc=$?"; brief $O/bench_n8_s2.log; tail -2 $O/bench_n8_s2.err
timeout 1200 $(tr 8 29522) bench.py --gpus 8 --config s3 --steps 2 --warmup 3 > $O/bench_n8_s3.log 2> $O/bench_n8_s3.err; echo "== N=8 S3 rc=$?"; brief $O/bench_n8_s3.log; tail -2 $O/bench_n8_s3.err
timeout 500 $(tr 4 29523) bench.py --gpus 4 --steps 5 --warmup 3 --no-verify > $O/bench_n4_s2.log 2> $O/bench_n4_s2.err; echo "== N=4 S2 rc=$?"; brief $O/bench_n4_s2.log; tail -2 $O/bench_n4_s2.err
