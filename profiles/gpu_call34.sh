#!/bin/bash
# round 2, call 34: validation of the state at the end of the round -- parity suite, default bench line (S2), S1 and
# S2-W20 lines, the CPU arm, launch lists with DRAM bytes (-> profiles/ncu_traffic.json), ncu --set full extracts
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
O=gpurun_out
timeout 2400 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -3 $O/pytest_gpu.log
brief() { python - "$1" <<'PY'
import json,sys
try:
    l=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print("   value %.3e %s | %.2f ms/step | e2e %.3e (%.2f ms) | launches %s" % (l["value"], l["unit"], l["ms_per_step"], l["e2e"]["value"], l["e2e"].get("ms_per_step", 0), l.get("gpu_launches")))
    if "stage_ms_per_step" in l:
        print("   stage", {k[3:]: round(v,2) for k,v in l["stage_ms_per_step"].items()}, "roofline frac %.3f" % l["roofline"]["frac"], "traffic", l["roofline"]["traffic"], l["roofline"].get("random_access"))
        print("   cpu", l.get("cpu_baseline"))
except Exception as e:
    print("   no JSON line:", e)
PY
}
timeout 900 python bench.py > $O/bench_default.log 2> $O/bench_default.err; echo "== default bench rc=$?"; brief $O/bench_default.log; tail -3 $O/bench_default.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_reference.log 2> $O/bench_reference.err; echo "== reference arm rc=$?"; brief $O/bench_reference.log; tail -3 $O/bench_reference.err
timeout 600 python bench.py --config s1 --steps 50 --no-cpu-baseline > $O/bench_s1.log 2> $O/bench_s1.err; echo "== S1 bench rc=$?"; brief $O/bench_s1.log; tail -3 $O/bench_s1.err
timeout 600 python bench.py --window-width 20 --steps 5 --no-cpu-baseline > $O/bench_s2_w20.log 2> $O/bench_s2_w20.err; echo "== S2 W=20 bench rc=$?"; brief $O/bench_s2_w20.log; tail -3 $O/bench_s2_w20.err
timeout 1200 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 200 --csv --log-file $O/launches_r02_s2_end.csv \
    python profiles/scale_step.py --scale 1.0 --steps 1 > $O/ncu_le.log 2>&1
echo "launchlist rc=$?"
timeout 1200 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 200 --csv --log-file $O/launches_r02_s2_w20_end.csv \
    python profiles/scale_step.py --scale 1.0 --steps 1 --window-width 20 > $O/ncu_lew.log 2>&1
python profiles/make_traffic.py $O/ncu_traffic.json s2:1.0:$O/launches_r02_s2_end.csv s2_w20:1.0:$O/launches_r02_s2_w20_end.csv; echo "traffic rc=$?"
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:"scan_direct_kernel|build_insert_kernel|confirm_pairs_kernel|build_scatter_kernel" -s 6 -c 4 \
    -o $O/prof_r02_s2_end python profiles/scale_step.py --scale 1.0 --steps 1 > $O/ncu_fe.log 2>&1
echo "ncufull rc=$?"
