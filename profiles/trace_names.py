"""Annotate an MSC_TRACE=1 log (per-launch device times keyed by api.cu line) with kernel names.
  python profiles/trace_names.py gpurun_out/trace_s2.log [segments_from_end]"""
import re
import sys
import os
src = open(os.path.join(os.path.dirname(__file__), "..", "muscato_b200", "csrc", "api.cu")).read().splitlines()


def name(line):
    for l in range(line - 1, max(0, line - 8), -1):
        m = re.search(r"launch_k\([^,]+,\s*([A-Za-z_0-9<>]+)", src[l])
        if m:
            return m.group(1)
        m = re.search(r"(\w+_kernel)", src[l])
        if m:
            return m.group(1)
    return src[line - 1].strip()[:60]


segs = [[]]
for l in open(sys.argv[1]):
    if "---- sync" in l:
        segs.append([])
        continue
    m = re.match(r"\[msc trace\]\s+([\d.]+) us\s+.*:(\d+)", l)
    if m:
        segs[-1].append((float(m.group(1)), int(m.group(2))))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 2
for s in [s for s in segs if s][-n:]:
    print("--- segment %.2f ms" % (sum(t for t, _ in s) / 1000))
    for t, l in s:
        print(f"{t:10.1f} us  :{l}  {name(l)}")
