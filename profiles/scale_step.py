"""One configs[2]-class step for profiling (ncu launch lists / `ncu --set full`): the block workload
of bench.py at `--scale f`, inputs uploaded once, then `--steps` resident steps (device pack + key
table build + scan + expansion + confirm + combine).  Prints one JSON line with the stage times.

  python profiles/scale_step.py [--config s2] [--scale 0.25] [--steps 2] [--window-width 15]
"""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from muscato_b200 import gendat  # noqa: E402
from muscato_b200.config import Config  # noqa: E402
from muscato_b200.engine import HotPath  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--config", default="s2")
ap.add_argument("--scale", type=float, default=0.25)
ap.add_argument("--steps", type=int, default=2)
ap.add_argument("--window-width", type=int, default=0)
ap.add_argument("--max-matches", type=int, default=0)
ap.add_argument("--trace", action="store_true")
ap.add_argument("--bloom-bits", type=int, default=0, help="Bloom bits per key (0 = the library's own sizing)")
args = ap.parse_args()
spec, cfgd, w = bench.pick_workload(args)
t0 = time.time()
reads, targets, _ = gendat.generate_blocks(spec)
gen_s = time.time() - t0
ro = np.arange(spec.n_reads + 1, dtype=np.uint64) * np.uint64(spec.read_len)
to = np.arange(spec.n_targets + 1, dtype=np.uint64) * np.uint64(spec.gene_len)
cfg = Config(**cfgd).apply_defaults()
with HotPath(cfg, device=0, keep_ascii=True, bloom_bits_per_key=args.bloom_bits) as hp:
    hp.set_reads((reads, ro))
    hp.set_targets((targets, to))
    hp.run()                       # sizes the bounded buffers
    hp.rebuild_and_run(3)
    hp.reset_stats()
    t1 = time.time()
    for _ in range(args.steps):
        hp.rebuild_and_run(3)
    step_ms = (time.time() - t1) / max(1, args.steps) * 1e3
    st = hp.stats()
K = max(1, args.steps)
print(json.dumps({"workload": bench.workload_name(spec, cfgd, w), "gen_s": round(gen_s, 2), "ms_per_step": step_ms,
                  "stage_ms": {k: st[k] / K for k in st if k.startswith("ms_")},
                  "counts": {k: st[k] for k in ("n_reads", "n_keys", "table_slots", "bloom_bytes", "target_bases", "n_candidates",
                                                "n_pairs", "n_pass", "n_matches_pre", "n_matches", "bloom_pass")}}))
