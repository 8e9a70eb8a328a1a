#!/bin/bash
# round 2, call 14: S4 bench mode (units of 5..8 bases); an L2-resident Bloom front at fewer bits per key
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
O=gpurun_out
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
show() { python - "$1" <<'PY'
import json,sys
l=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
s=l["stage_ms"]; print("   step %.2f ms | " % l["ms_per_step"] + " ".join(f"{k[3:]}={v:.2f}" for k,v in s.items())); print("   ", l["counts"])
PY
}
timeout 600 python bench.py --config s4 --steps 3 --warmup 3 --no-cpu-baseline > $O/bench_s4.log 2> $O/bench_s4.err; echo "== S4 rc=$?"; brief $O/bench_s4.log; tail -3 $O/bench_s4.err
timeout 600 python bench.py --config s4 --scale 0.1 --max-matches 1000 --steps 1 --warmup 3 --no-cpu-baseline > $O/bench_s4_mm1000.log 2> $O/bench_s4_mm1000.err; echo "== S4/10 MaxMatches 1000 rc=$?"; brief $O/bench_s4_mm1000.log; tail -3 $O/bench_s4_mm1000.err
for bits in 0 12 7 5; do
  timeout 400 python profiles/scale_step.py --scale 0.25 --steps 2 --bloom-bits $bits > $O/step_0.25_bits$bits.log 2>&1; echo "== scale 0.25 bloom bits $bits rc=$?"; show $O/step_0.25_bits$bits.log
done
for bits in 0 3; do
  timeout 400 python profiles/scale_step.py --scale 0.5 --steps 2 --bloom-bits $bits > $O/step_0.5_bits$bits.log 2>&1; echo "== scale 0.5 bloom bits $bits rc=$?"; show $O/step_0.5_bits$bits.log
done
