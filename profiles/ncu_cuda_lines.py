"""Per-CUDA-line instruction / stall-sample shares from
`ncu -i rep --page source --csv --print-source cuda,sass --kernel-name regex:K > file.csv`."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 25
cur_file = ""
out = []
hdr = None
for r in rows:
    if len(r) >= 2 and r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
        continue
    if len(r) > 8 and r[0] == "Line No":
        hdr = {k: i for i, k in enumerate(r)}
        continue
    if hdr and len(r) > 8 and r[0].isdigit():
        try:
            out.append((cur_file, int(r[0]), r[1], float(r[hdr["# Samples"]] or 0), float(r[hdr["Instructions Executed"]] or 0)))
        except ValueError:
            pass
ts = sum(o[3] for o in out) or 1
ti = sum(o[4] for o in out) or 1
print(f"total warp instructions {ti:.0f}, samples {ts:.0f}")
for o in sorted(out, key=lambda o: -o[3])[:n]:
    print(f"{100 * o[3] / ts:5.1f}% samples {100 * o[4] / ti:5.1f}% inst  {o[0]}:{o[1]:<4d} {o[2].strip()[:95]}")
