#!/bin/bash
# round 2, call 40: read parts on one GPU (upload of part i+1 under the compute of part i): parity + e2e with 1/2/3 parts
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
O=gpurun_out
timeout 1200 python -m pytest tests/test_gpu_readparts.py -m gpu -x -q > $O/pytest_parts.log 2>&1; echo "pytest rc=$?"
tail -5 $O/pytest_parts.log
brief() { python - "$1" <<'PY'
import json,sys
try:
    l=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print("   value %.3e %s | %.2f ms/step | e2e %.3e (%.2f ms, %s parts, h2d %.2f GB d2h %.2f GB) | launches %s" % (l["value"], l["unit"], l["ms_per_step"], l["e2e"]["value"], l["e2e"].get("ms_per_step", 0), l["e2e"].get("read_parts"), l["e2e"]["h2d_bytes_per_step"]/1e9, l["e2e"]["d2h_bytes_per_step"]/1e9, l.get("gpu_launches")))
except Exception as e:
    print("   no JSON line:", e)
PY
}
for p in 2 1 3; do
timeout 900 python bench.py --steps 5 --e2e-parts $p --no-cpu-baseline > $O/bench_parts$p.log 2> $O/bench_parts$p.err; echo "== bench e2e parts $p rc=$?"; brief $O/bench_parts$p.log; tail -3 $O/bench_parts$p.err
done
