#!/bin/bash
# round 2, call 33: partition counters one per 128-byte line
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
O=gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "random_cases or golden or gendat_medium or maxmatches" > $O/pytest_front.log 2>&1; echo "pytest rc=$?"
tail -3 $O/pytest_front.log
show() { python - "$1" <<'PY'
import json,sys
try:
    l=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    s=l["stage_ms"]; print("   step %.2f ms | " % l["ms_per_step"] + " ".join(f"{k[3:]}={v:.2f}" for k,v in s.items()))
except Exception as e:
    print("   failed:", e)
PY
}
run() { # name scale env...
  name=$1; sc=$2; shift; shift
  env "$@" timeout 400 python profiles/scale_step.py --scale $sc --steps 2 $EXTRA > $O/step_${sc}_$name.log 2>&1; echo "== scale $sc $name rc=$?"; show $O/step_${sc}_$name.log
}
run b2 1.0 MSC_X=1
run b2 0.25 MSC_X=1
EXTRA="--window-width 20" run b2_w20 1.0 MSC_X=1
MSC_TRACE=1 timeout 400 python profiles/scale_step.py --scale 1.0 --steps 1 > $O/trace_s2_b2.log 2>&1; python profiles/trace_names.py $O/trace_s2_b2.log 1 | grep -E "build_|table_clear|pack_" 
timeout 1200 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"build_|table_clear" -c 40 --csv --log-file $O/launches_r02_s2_build_b2.csv \
    python profiles/scale_step.py --scale 1.0 --steps 1 > $O/ncu_lb2.log 2>&1
echo "launchlist rc=$?"
python - <<'PY'
import csv
rows=list(csv.reader(l for l in open('gpurun_out/launches_r02_s2_build_b2.csv') if l.startswith('"')))
ix={h:i for i,h in enumerate(rows[0])}
d={}
for r in rows[1:]:
    d.setdefault(int(r[ix["ID"]]),{"n":r[ix["Kernel Name"]].split("(")[0]})[r[ix["Metric Name"]]]=float(r[ix["Metric Value"]].replace(",",""))
for i,k in list(d.items())[-14:]:
    print(i,k["n"],k.get("gpu__time_duration.sum"),"ns rd",k.get("dram__bytes_read.sum"),"wr",k.get("dram__bytes_write.sum"))
PY
