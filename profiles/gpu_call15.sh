#!/bin/bash
# round 2, call 15: parity suite; S4 bench mode; the default bench line (S2, as the driver runs it); launch list with
# DRAM bytes -> profiles/ncu_traffic.json; ncu --set full of the other large kernels at S2
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
timeout 300 python bench.py --config s4 --steps 3 --warmup 3 --no-cpu-baseline > $O/bench_s4.log 2> $O/bench_s4.err; echo "== S4 rc=$?"; brief $O/bench_s4.log; tail -3 $O/bench_s4.err
timeout 300 python bench.py --config s4 --scale 0.2 --max-matches 1000 --steps 2 --warmup 3 --no-cpu-baseline > $O/bench_s4_mm1000.log 2> $O/bench_s4_mm1000.err; echo "== S4/5 MaxMatches 1000 rc=$?"; brief $O/bench_s4_mm1000.log; tail -3 $O/bench_s4_mm1000.err
timeout 900 python bench.py > $O/bench_default.log 2> $O/bench_default.err; echo "== default bench rc=$?"; brief $O/bench_default.log; tail -3 $O/bench_default.err
timeout 1200 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 120 --csv --log-file $O/launches_r02_s2_final.csv \
    python profiles/scale_step.py --scale 1.0 --steps 1 > $O/ncu_lf.log 2>&1
echo "launchlist rc=$?"
timeout 1200 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 120 --csv --log-file $O/launches_r02_s2_w20_final.csv \
    python profiles/scale_step.py --scale 1.0 --steps 1 --window-width 20 > $O/ncu_lfw.log 2>&1
python profiles/make_traffic.py $O/ncu_traffic.json s2:1.0:$O/launches_r02_s2_final.csv s2_w20:1.0:$O/launches_r02_s2_w20_final.csv; echo "traffic rc=$?"
timeout 1800 ncu --set full --clock-control none --import-source on -k regex:"build_insert_kernel|build_windows_kernel|build_scatter_kernel|confirm_pairs_kernel|combine_scatter_kernel|scan_resident_kernel" -s 6 -c 7 \
    -o $O/prof_r02_s2_others python profiles/scale_step.py --scale 1.0 --steps 1 > $O/ncu_fo.log 2>&1
echo "ncufull rc=$?"
