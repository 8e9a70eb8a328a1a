"""Wall-clock breakdown of one e2e step of the bench workload (host call by host call)."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from muscato_b200 import gendat
from muscato_b200.config import Config
from muscato_b200.engine import HotPath
syn = gendat.generate(**bench.WORK)
pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
rd_a, rd_o = pin(syn.read_ascii), pin(syn.read_offs.view(np.int64))
tg_a, tg_o = pin(syn.target_ascii), pin(syn.target_offs.view(np.int64))
res = torch.empty(16 << 22, dtype=torch.uint8).pin_memory()
hp = HotPath(Config(**bench.CFG).apply_defaults(), device=0, keep_ascii=True)
acc = {}
for it in range(13):
    t = [time.perf_counter()]
    hp.set_reads_ptr(rd_a.data_ptr(), rd_o.data_ptr(), syn.n_reads); t.append(time.perf_counter())
    hp.set_targets_ptr(tg_a.data_ptr(), tg_o.data_ptr(), syn.n_targets); t.append(time.perf_counter())
    hp.run(); t.append(time.perf_counter())
    n = hp.fetch_into(res.data_ptr(), 1 << 22); t.append(time.perf_counter())
    if it >= 3:
        for k, a, b in zip(("set_reads", "set_targets", "run", "fetch_into"), t, t[1:]):
            acc[k] = acc.get(k, 0) + (b - a) * 1e3 / 10
print({k: round(v, 3) for k, v in acc.items()}, "total", round(sum(acc.values()), 3), "matches", n)
# raw pinned H2D bandwidth for reference
d = torch.empty_like(rd_a, device="cuda")
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(10): d.copy_(rd_a, non_blocking=True)
torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 10
print("pinned H2D 100 MB: %.3f ms = %.1f GB/s" % (dt * 1e3, rd_a.numel() / dt / 1e9))
