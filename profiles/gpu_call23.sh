#!/bin/bash
# round 2, call 23: front choice at smaller key counts; launch list with DRAM bytes at S2 (-> profiles/ncu_traffic.json);
# ncu --set full of the exact-front scan kernel and the build kernels at S2
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
O=gpurun_out
show() { python - "$1" <<'PY'
import json,sys
try:
    l=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    s=l["stage_ms"]; print("   step %.2f ms | " % l["ms_per_step"] + " ".join(f"{k[3:]}={v:.2f}" for k,v in s.items()))
    print("   ", {k: l["counts"][k] for k in ("n_keys", "target_bases", "bloom_bytes", "n_candidates", "bloom_pass", "n_matches")})
except Exception as e:
    print("   failed:", e)
PY
}
run() { # name scale env...
  name=$1; sc=$2; shift; shift
  env "$@" timeout 400 python profiles/scale_step.py --scale $sc --steps 2 > $O/step_${sc}_$name.log 2>&1; echo "== scale $sc $name rc=$?"; show $O/step_${sc}_$name.log
}
for sc in 0.125 0.05 0.02; do
  run bloom $sc MSC_FRONT_DIRECT=0
  for mb in 32 64 128; do run direct$mb $sc MSC_FRONT_DIRECT=1 MSC_FRONT_PASS_MB=$mb; done
done
timeout 1200 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 160 --csv --log-file $O/launches_r02_s2_v8.csv \
    python profiles/scale_step.py --scale 1.0 --steps 1 > $O/ncu_l8.log 2>&1
echo "launchlist rc=$?"
timeout 1200 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 160 --csv --log-file $O/launches_r02_s2_w20_v8.csv \
    python profiles/scale_step.py --scale 1.0 --steps 1 --window-width 20 > $O/ncu_l8w.log 2>&1
python profiles/make_traffic.py $O/ncu_traffic.json s2:1.0:$O/launches_r02_s2_v8.csv s2_w20:1.0:$O/launches_r02_s2_w20_v8.csv; echo "traffic rc=$?"
# second step of the run = launches after the sizing step: skip the first step's kernels by name count (scan: 4 launches per step)
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:"scan_direct_kernel" -s 5 -c 2 \
    -o $O/prof_r02_s2_scan_direct python profiles/scale_step.py --scale 1.0 --steps 1 > $O/ncu_fd.log 2>&1
echo "ncufull scan rc=$?"
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:"build_insert_kernel|build_windows_kernel|build_scatter_kernel" -s 3 -c 3 \
    -o $O/prof_r02_s2_build python profiles/scale_step.py --scale 1.0 --steps 1 > $O/ncu_fb.log 2>&1
echo "ncufull build rc=$?"
