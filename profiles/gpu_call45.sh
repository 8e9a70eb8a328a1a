#!/bin/bash
# round 2, call 45: the other configurations with the final kernels -- configs[4] (S4, default MaxMatches and 1000),
# configs[1] (S1m), S2 at W = 20
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
O=gpurun_out
brief() { python - "$1" <<'PY'
import json,sys
try:
    l=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print("   value %.3e %s | %.2f ms/step | e2e %.3e (%.2f ms, %s parts) | launches %s" % (l["value"], l["unit"], l["ms_per_step"], l["e2e"]["value"], l["e2e"].get("ms_per_step", 0), l["e2e"].get("read_parts"), l.get("gpu_launches")))
    print("   stage", {k[3:]: round(v,2) for k,v in l["stage_ms_per_step"].items()}, "roofline frac %.3f" % l["roofline"]["frac"]); print("   ", l["counts"])
except Exception as e:
    print("   no JSON line:", e)
PY
}
timeout 300 python bench.py --config s4 --steps 3 --warmup 3 --no-cpu-baseline > $O/bench_s4_end.log 2> $O/bench_s4_end.err; echo "== S4 rc=$?"; brief $O/bench_s4_end.log; tail -2 $O/bench_s4_end.err
timeout 300 python bench.py --config s4 --scale 0.2 --max-matches 1000 --steps 2 --warmup 3 --no-cpu-baseline > $O/bench_s4_mm1000_end.log 2> $O/bench_s4_mm1000_end.err; echo "== S4/5 MaxMatches 1000 rc=$?"; brief $O/bench_s4_mm1000_end.log; tail -2 $O/bench_s4_mm1000_end.err
timeout 300 python bench.py --config s1 --steps 50 --no-cpu-baseline > $O/bench_s1_end.log 2> $O/bench_s1_end.err; echo "== S1 rc=$?"; brief $O/bench_s1_end.log; tail -2 $O/bench_s1_end.err
timeout 400 python bench.py --window-width 20 --steps 5 --no-cpu-baseline > $O/bench_s2_w20_end.log 2> $O/bench_s2_w20_end.err; echo "== S2 W=20 rc=$?"; brief $O/bench_s2_w20_end.log; tail -2 $O/bench_s2_w20_end.err
