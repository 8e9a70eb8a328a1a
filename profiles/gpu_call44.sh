#!/bin/bash
# round 2, call 44 (4 GPUs): S2 strong scaling at N=4 and N=2 with the end-of-round kernels
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
O=gpurun_out
brief() { python - "$1" <<'PY'
import json,sys
try:
    l=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print("   value %.3e %s | %.2f ms/step | e2e %.3e (%.2f ms) | grid %s | verify %s (%s matches) | launches %s | hbm %.1f GB" % (l["value"], l["unit"], l["ms_per_step"], l["e2e"]["value"], l["e2e"]["ms_per_step"], l["setup"]["grid"], l.get("verify_sharded_equals_unsharded"), l.get("verify_matches"), l["gpu_launches"], l["setup"]["hbm_used_bytes_max_rank"]/1e9))
    print("   stage", {k[3:]: round(v,2) for k,v in l["stage_ms_per_step"].items()}, "roofline frac %.3f" % l["roofline"]["frac"])
except Exception as e:
    print("   no JSON line:", e)
PY
}
tr() { echo "python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port $2"; }
timeout 500 $(tr 4 29541) bench.py --gpus 4 --steps 5 --warmup 3 > $O/bench_n4_s2_end.log 2> $O/bench_n4_s2_end.err; echo "== N=4 S2 rc=$?"; brief $O/bench_n4_s2_end.log; tail -2 $O/bench_n4_s2_end.err
timeout 500 $(tr 2 29542) bench.py --gpus 2 --steps 5 --warmup 3 --no-verify > $O/bench_n2_s2_end.log 2> $O/bench_n2_s2_end.err; echo "== N=2 S2 rc=$?"; brief $O/bench_n2_s2_end.log; tail -2 $O/bench_n2_s2_end.err
