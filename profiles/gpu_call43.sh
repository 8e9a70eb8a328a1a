#!/bin/bash
# round 2, call 43: validation of the final state -- smoke(), parity suite, default bench line (S2), the CPU arm, launch
# lists with DRAM bytes (-> profiles/ncu_traffic.json), ncu --set full extracts of the scan and confirm kernels
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
O=gpurun_out
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 $O/smoke.log
timeout 2400 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -3 $O/pytest_gpu.log
brief() { python - "$1" <<'PY'
import json,sys
try:
    l=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print("   value %.3e %s | %.2f ms/step | e2e %.3e (%.2f ms) | launches %s" % (l["value"], l["unit"], l["ms_per_step"], l["e2e"]["value"], l["e2e"].get("ms_per_step", 0), l.get("gpu_launches")))
    if "stage_ms_per_step" in l:
        print("   stage", {k[3:]: round(v,2) for k,v in l["stage_ms_per_step"].items()}, "roofline frac %.3f" % l["roofline"]["frac"], "traffic", l["roofline"]["traffic"], l["roofline"]["traffic_source"])
        print("   cpu", l.get("cpu_baseline"))
except Exception as e:
    print("   no JSON line:", e)
PY
}
timeout 900 python bench.py > $O/bench_default.log 2> $O/bench_default.err; echo "== default bench rc=$?"; brief $O/bench_default.log; tail -3 $O/bench_default.err
timeout 600 python bench.py --impl reference --steps 1 --warmup 1 > $O/bench_reference.log 2> $O/bench_reference.err; echo "== reference arm rc=$?"; brief $O/bench_reference.log; tail -3 $O/bench_reference.err
timeout 1200 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 200 --csv --log-file $O/launches_r02_s2_final2.csv \
    python profiles/scale_step.py --scale 1.0 --steps 1 > $O/ncu_lf2.log 2>&1
echo "launchlist rc=$?"
timeout 1200 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 200 --csv --log-file $O/launches_r02_s2_w20_final2.csv \
    python profiles/scale_step.py --scale 1.0 --steps 1 --window-width 20 > $O/ncu_lfw2.log 2>&1
python profiles/make_traffic.py $O/ncu_traffic.json s2:1.0:$O/launches_r02_s2_final2.csv s2_w20:1.0:$O/launches_r02_s2_w20_final2.csv; echo "traffic rc=$?"
# the resident step of the run: launches after the sizing step (scan_direct: 4 per step; confirm: 1 per step)
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:"scan_direct_kernel|confirm_pairs_kernel" -s 5 -c 5 \
    -o $O/prof_r02_s2_final2 python profiles/scale_step.py --scale 1.0 --steps 1 > $O/ncu_ff2.log 2>&1
echo "ncufull rc=$?"
