#!/bin/bash
# round 2, call 20: full parity suite with the exact front + chunked read upload; default bench line (S2)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
O=gpurun_out
timeout 2400 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -3 $O/pytest_gpu.log
brief() { python - "$1" <<'PY'
import json,sys
try:
    l=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print("   value %.3e %s | %.2f ms/step | e2e %.3e (%.2f ms) | launches %s" % (l["value"], l["unit"], l["ms_per_step"], l["e2e"]["value"], l["e2e"]["ms_per_step"], l["gpu_launches"]))
    print("   stage", {k[3:]: round(v,2) for k,v in l["stage_ms_per_step"].items()}, "roofline frac %.3f" % l["roofline"]["frac"]); print("   ", l["counts"])
except Exception as e:
    print("   no JSON line:", e)
PY
}
timeout 900 python bench.py > $O/bench_default.log 2> $O/bench_default.err; echo "== default bench rc=$?"; brief $O/bench_default.log; tail -3 $O/bench_default.err
timeout 600 python bench.py --config s1 --steps 50 --no-cpu-baseline > $O/bench_s1.log 2> $O/bench_s1.err; echo "== S1 bench rc=$?"; brief $O/bench_s1.log; tail -3 $O/bench_s1.err
