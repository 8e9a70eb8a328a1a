"""Top stall sites from `ncu --page source --csv` output: python ncu_source_top.py file.csv [N]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 15
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
body = [r for r in rows[2:] if len(r) > ix['# Samples']]
tot = sum(float(r[ix['# Samples']] or 0) for r in body)
body.sort(key=lambda r: -float(r[ix['# Samples']] or 0))
for r in body[:n]:
    s = float(r[ix['# Samples']] or 0)
    stalls = {k[6:]: float(r[ix[k]] or 0) for k in hdr if k.startswith('stall_') and '(Not' not in k}
    top = sorted(stalls.items(), key=lambda kv: -kv[1])[:2]
    print(f"{100 * s / tot:5.1f}%  {r[ix['Source']][:90]:90s} exec={r[ix['Instructions Executed']]:>9s} {top}")
