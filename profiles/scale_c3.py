"""BASELINE.json configs[2] at full size on ONE GPU (not the bench workload): 1e8 x 100 bp reads against
1e6 targets x 1 kb plus their reverse complements (2e6 target lines, 2e9 bases), three windows
(S2 of SURVEY.md 8d: 50 % of the reads sampled from the targets with 2 % substitutions).
Checks two size-independent properties on the result and prints one JSON line:
  precision  every match of a random sample has nx == Hamming(read, target[pos:pos+L]) <= nmiss;
  recall     every sampled planted read whose planted site has <= nmiss mismatches and at least one
             mutation-free window (of sufficient dinucleotide count, and not subject to the
             position-0 rule Q1) is reported at that site unless MMTol removed it (a better site
             exists), which is checked too.
Reads are not uniqified / sorted here (duplicates are legal input for the key table).

  python profiles/scale_c3.py [n_reads] [n_genes] [gene_len] [W,W,...]
"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from muscato_b200.config import Config  # noqa: E402
from muscato_b200.engine import HotPath  # noqa: E402

n_reads = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000_000
n_genes = int(sys.argv[2]) if len(sys.argv) > 2 else 1_000_000
GL = int(sys.argv[3]) if len(sys.argv) > 3 else 1000
L = 100
WWS = [int(x) for x in sys.argv[4].split(",")] if len(sys.argv) > 4 else [15]


def mem_available_gb():
    for ln in open("/proc/meminfo"):
        if ln.startswith("MemAvailable"):
            return int(ln.split()[1]) / 1e6
    return 0.0


need_gb = (n_reads * L * 2.5 + 2 * n_genes * GL * 3.0) / 1e9 + 8
avail = mem_available_gb()
if avail < need_gb:
    scale = max(0.05, (avail - 8) / need_gb)
    n_reads = int(n_reads * scale)
    n_genes = int(n_genes * scale)
    print(f"# only {avail:.0f} GB of host memory available: scaled to {n_reads} reads x {n_genes} genes", file=sys.stderr)

rng = np.random.default_rng(3)
bases = np.frombuffer(b"ACGT", dtype=np.uint8)
comp = np.zeros(256, dtype=np.uint8)
comp[list(b"ACGT")] = list(b"TGCA")
t0 = time.time()
fw = bases[rng.integers(0, 4, size=(n_genes, GL), dtype=np.uint8)]
tg = np.empty((2 * n_genes, GL), dtype=np.uint8)      # prep_targets -rev: every target is followed by its reverse complement
tg[0::2] = fw
tg[1::2] = comp[fw[:, ::-1]]
del fw
n_tg = 2 * n_genes
reads = np.empty((n_reads, L), dtype=np.uint8)
half = n_reads // 2
plant_g = np.empty(half, dtype=np.int64)
plant_p = np.empty(half, dtype=np.int32)
CH = 1_000_000
ar = np.arange(L)[None, :]
for lo in range(0, n_reads, CH):
    hi = min(n_reads, lo + CH)
    if lo < half:                                   # sampled from the targets with 2 % substitutions
        hi = min(hi, half)
        g = rng.integers(0, n_tg, size=hi - lo)
        p = rng.integers(0, GL - L + 1, size=hi - lo)
        s = tg[g[:, None], p[:, None] + ar]
        mut = rng.random(s.shape, dtype=np.float32) < 0.02
        s[mut] = bases[rng.integers(0, 4, size=int(mut.sum()), dtype=np.uint8)]
        reads[lo:hi] = s
        plant_g[lo:hi] = g
        plant_p[lo:hi] = p
    else:
        reads[lo:hi] = bases[rng.integers(0, 4, size=(hi - lo, L), dtype=np.uint8)]
gen_s = time.time() - t0
for WW in WWS:
    cfg = Config(Windows=[0, 20, 40], WindowWidth=WW, MaxReadLength=100, PMatch=0.97, MinDinuc=5, MMTol=1).apply_defaults()
    nmiss = int((1 - cfg.PMatch) * float(L))
    read_offs = np.arange(n_reads + 1, dtype=np.uint64) * np.uint64(L)
    tg_offs = np.arange(n_tg + 1, dtype=np.uint64) * np.uint64(GL)
    with HotPath(cfg, device=0, keep_ascii=True) as hp:
        t1 = time.time()
        hp.set_reads((reads.ravel(), read_offs))
        hp.set_targets((tg.ravel(), tg_offs))
        hp.run()
        first_s = time.time() - t1
        hp.rebuild_and_run(3)
        hp.reset_stats()
        K = 3
        t2 = time.time()
        for _ in range(K):
            hp.rebuild_and_run(3)
        step_ms = (time.time() - t2) / K * 1e3
        st = hp.stats()
        t3 = time.time()
        m = hp.fetch()
        fetch_s = time.time() - t3

    # ---- size-independent checks -------------------------------------------------------------------
    t4 = time.time()
    order_ok = True
    if len(m) > 1:
        k1 = m["read_id"].astype(np.int64)
        k2 = m["gene_id"].astype(np.int64) * (GL + 1) + m["pos"].astype(np.int64)
        order_ok = bool(np.all((k1[1:] > k1[:-1]) | ((k1[1:] == k1[:-1]) & (k2[1:] > k2[:-1]))))   # sorted, no duplicates
    samp = rng.integers(0, len(m), size=min(len(m), 200_000))
    ms = m[samp]
    seg = tg[ms["gene_id"].astype(np.int64)[:, None], ms["pos"].astype(np.int64)[:, None] + ar]
    ham = (seg != reads[ms["read_id"].astype(np.int64)]).sum(axis=1)
    precision_ok = bool(np.all(ham == ms["nx"]) and np.all(ham <= nmiss) and np.all(ms["pos"].astype(np.int64) + L <= GL))
    # recall on planted reads
    ps = rng.integers(0, half, size=min(half, 200_000))
    site = tg[plant_g[ps][:, None], plant_p[ps].astype(np.int64)[:, None] + ar]
    diff = site != reads[ps]
    nx_site = diff.sum(axis=1)
    win_ok = np.zeros(len(ps), dtype=bool)
    for q1 in cfg.Windows:
        w = reads[ps][:, q1:q1 + cfg.WindowWidth]
        pairs = w[:, :-1].astype(np.int32) * 256 + w[:, 1:]
        clean = ~diff[:, q1:q1 + cfg.WindowWidth].any(axis=1)
        sp = np.sort(pairs, axis=1)                      # distinct adjacent pairs (utils/entropy.go:5-40)
        ndin = 1 + (sp[:, 1:] != sp[:, :-1]).sum(axis=1)
        # Q1 (cmd/muscato_screen/main.go:305): a window at target position 0 only carries reads of
        # at most min(100 - W, len) bases, so a 100 bp read planted at position 0 is not found through window 0
        q1_ok = ~((plant_p[ps].astype(np.int64) + q1 == 0) & (L > min(100 - cfg.WindowWidth, GL)))
        win_ok |= clean & (ndin >= cfg.MinDinuc) & q1_ok
    must = win_ok & (nx_site <= nmiss)
    first = np.searchsorted(m["read_id"], ps, side="left")
    last = np.searchsorted(m["read_id"], ps, side="right")
    missing = 0
    for i in np.nonzero(must)[0]:
        rows = m[first[i]:last[i]]
        hit = rows[(rows["gene_id"] == plant_g[ps[i]]) & (rows["pos"] == plant_p[ps[i]])]
        if len(hit) == 1 and int(hit[0]["nx"]) == int(nx_site[i]):
            continue
        # legitimately removed by MMTol: the read has a site with nx < nx_site - MMTol
        if len(rows) and int(rows["nx"].min()) + cfg.MMTol < int(nx_site[i]):
            continue
        missing += 1
    check_s = time.time() - t4

    T = n_tg * GL
    scan_ms = st["ms_scan"] / K
    peak = 6547.2
    try:
        peak = float(json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        pass
    alg = T / 4 + 32.0 * T + 16.0 * st["n_candidates"]
    print(json.dumps({
        "workload": "BASELINE.json configs[2] (S2): reads x targets with -rev, Windows=0,20,40 W=%d PMatch=0.97 MinDinuc=5 MMTol=1, one GPU" % WW,
        "reads": n_reads, "keys": st["n_keys"], "targets": n_tg, "target_bases": T, "bloom_bytes": st["bloom_bytes"],
        "table_slots": st["table_slots"], "candidates": st["n_candidates"], "pairs": st["n_pairs"],
        "passing_pairs": st["n_pass"], "matches": st["n_matches"],
        "gen_s": round(gen_s, 1), "first_call_s": round(first_s, 2), "fetch_s": round(fetch_s, 2), "check_s": round(check_s, 1),
        "ms_per_step": step_ms, "bases_per_s": T / (step_ms * 1e-3), "pairs_per_s": st["n_pairs"] / K / (step_ms * 1e-3),
        "stage_ms": {k: st[k] / K for k in st if k.startswith("ms_") and k != "ms_scan_kernel"},
        "scan": {"ms": scan_ms, "positions_per_s": T / (scan_ms * 1e-3), "algorithmic_bytes": alg,
                 "achieved_gbs": alg / (scan_ms * 1e-3) / 1e9, "peak_gbs": peak, "frac_of_hbm": alg / (scan_ms * 1e-3) / 1e9 / peak,
                 "note": "filter larger than L2: T/4 + 32 B per probed position + 16 B per candidate (SURVEY 8d)"},
        "checks": {"sorted_unique": order_ok, "precision_sample": int(len(samp)), "precision_ok": precision_ok,
                   "recall_sample": int(must.sum()), "recall_missing": int(missing)}}))
