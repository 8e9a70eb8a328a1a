"""CPU stand-in used only by `MSC_SCALE_DRY=1 python profiles/scale_c3.py ...` to exercise the
generator and the property checks of that script without a GPU.  It reports the planted sites
only; it is not part of the product or of any test."""
import numpy as np

MATCH_DTYPE = np.dtype([("read_id", "<u4"), ("gene_id", "<u4"), ("pos", "<u4"), ("nx", "<u4")])


class HotPath:
    def __init__(self, cfg, **kw):
        self.cfg = cfg

    def __enter__(self):
        return self

    def __exit__(self, *a):
        pass

    def set_reads(self, x):
        self.reads = x

    def set_targets(self, x):
        self.tg = x

    def run(self):
        pass

    def rebuild_and_run(self, w):
        pass

    def reset_stats(self):
        pass

    def stats(self):
        z = dict.fromkeys(["n_keys", "bloom_bytes", "table_slots", "n_candidates", "n_pairs", "n_pass", "n_matches"], 0)
        z.update({k: 1.0 for k in ["ms_scan", "ms_build", "ms_scan_kernel"]})
        return z

    def fetch(self):
        import __main__ as M
        L, GL = M.L, M.GL
        rid = np.arange(M.half)
        site = M.tg[M.plant_g[:, None], M.plant_p.astype(np.int64)[:, None] + np.arange(L)[None, :]]
        nx = (site != M.reads[rid]).sum(axis=1)
        keep = nx <= 3
        m = np.zeros(int(keep.sum()), dtype=MATCH_DTYPE)
        m["read_id"], m["gene_id"], m["pos"], m["nx"] = rid[keep], M.plant_g[keep], M.plant_p[keep], nx[keep]
        return m
