#!/bin/bash
# round 2, call 37: scatter kernel with its first phase unrolled (more loads in flight per thread)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
O=gpurun_out
cp muscato_b200/libmuscato_b200.so /tmp/lib_base.so
for v in base su2 su5; do
  if [ $v = base ]; then cp /tmp/lib_base.so muscato_b200/libmuscato_b200.so; else cp muscato_b200/libmuscato_b200_$v.so muscato_b200/libmuscato_b200.so; fi
  MSC_TRACE=1 timeout 400 python profiles/scale_step.py --scale 1.0 --steps 1 > $O/trace_s2_$v.log 2>&1; echo "== $v rc=$?"; python profiles/trace_names.py $O/trace_s2_$v.log 1 | grep -E "build_scatter|build_windows|build_insert"
done
cp /tmp/lib_base.so muscato_b200/libmuscato_b200.so
