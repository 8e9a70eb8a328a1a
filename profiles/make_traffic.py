"""profiles/ncu_traffic.json from ncu launch lists (gpu__time_duration + dram__bytes_read/write per launch):
per kernel of the LAST step, the DRAM bytes per launch -- keyed by the hash of the CUDA sources the capture was
taken from (bench.csrc_sha), so that bench.py can tell when the figures are stale.

  python profiles/make_traffic.py OUT.json CONFIG:SCALE:launches.csv [CONFIG:SCALE:launches.csv ...]
"""
import csv
import json
import os
import sys
from collections import OrderedDict

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402


def last_step(path):
    rows = list(csv.reader(l for l in open(path) if l.startswith('"')))
    ix = {h: i for i, h in enumerate(rows[0])}
    d = OrderedDict()
    for r in rows[1:]:
        d.setdefault(int(r[ix["ID"]]), {"name": r[ix["Kernel Name"]].split("(")[0].replace("void ", "")})[r[ix["Metric Name"]]] = float(r[ix["Metric Value"]].replace(",", ""))
    ids = list(d)
    starts = [i for i in ids if d[i]["name"].startswith("pack_reads_kernel")]
    last = [i for i in ids if i >= starts[-1]] if starts else ids
    out = OrderedDict()
    for i in last:
        k = d[i]
        nm = k["name"].split("<")[0]
        e = out.setdefault(nm, {"launches": 0, "dram_read_bytes": 0.0, "dram_write_bytes": 0.0, "ns": 0.0})
        e["launches"] += 1
        e["dram_read_bytes"] += k.get("dram__bytes_read.sum", 0.0)
        e["dram_write_bytes"] += k.get("dram__bytes_write.sum", 0.0)
        e["ns"] += k["gpu__time_duration.sum"]
    for e in out.values():
        for f in ("dram_read_bytes", "dram_write_bytes", "ns"):
            e[f] /= e["launches"]
    return out


def main():
    out_path = sys.argv[1]
    caps = []
    for spec in sys.argv[2:]:
        cfg, scale, path = spec.split(":", 2)
        caps.append({"config": cfg, "scale": float(scale), "csrc_sha": bench.csrc_sha(), "source": "profiles/r02/" + os.path.basename(path),
                     "how": "ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none, "
                            "profiles/scale_step.py (one resident step), kernels of the last step, per launch",
                     "kernels": last_step(path)})
    json.dump({"captures": caps}, open(out_path, "w"), indent=1)


if __name__ == "__main__":
    main()
