"""Summarise an ncu launch-list CSV (gpu__time_duration + dram bytes): one line per kernel of the LAST step.
  python profiles/launch_summary.py gpurun_out/launches_r02_s2.csv [first_kernel_of_step]"""
import csv
import sys
from collections import OrderedDict

rows = list(csv.reader(l for l in open(sys.argv[1]) if l.startswith('"')))
ix = {h: i for i, h in enumerate(rows[0])}
d = OrderedDict()
for r in rows[1:]:
    d.setdefault(int(r[ix["ID"]]), {"name": r[ix["Kernel Name"]].split("(")[0].replace("void ", "")})[r[ix["Metric Name"]]] = float(r[ix["Metric Value"]].replace(",", ""))
ids = list(d)
first = sys.argv[2] if len(sys.argv) > 2 else "pack_reads_kernel"
starts = [i for i in ids if d[i]["name"].startswith(first)]
last = [i for i in ids if i >= starts[-1]] if starts else ids
tot = sum(d[i]["gpu__time_duration.sum"] for i in last)
print(f"last step: {len(last)} launches, {tot / 1e6:.3f} ms of kernel time")
for i in last:
    k = d[i]
    t = k["gpu__time_duration.sum"]
    rd, wr = k.get("dram__bytes_read.sum", 0), k.get("dram__bytes_write.sum", 0)
    print(f"{k['name'][:44]:44s} {t / 1e3:10.1f} us {100 * t / tot:5.1f}%  dram rd {rd / 1e9:8.3f} GB wr {wr / 1e9:8.3f} GB  {(rd + wr) / t:7.1f} GB/s")
