import sys, os, numpy as np
sys.path.insert(0, os.getcwd())
import bench
from muscato_b200 import gendat
from muscato_b200.config import Config
from muscato_b200.engine import HotPath
syn = gendat.generate(**bench.WORK)
hp = HotPath(Config(**bench.CFG).apply_defaults(), device=0)
hp.set_reads((syn.read_ascii, syn.read_offs)); hp.set_targets((syn.target_ascii, syn.target_offs)); hp.run()
m = hp.fetch()
c = np.bincount(m["read_id"])
c = c[c > 0]
print("reads with matches", len(c), "max", c.max())
for lo, hi in [(1,1),(2,8),(9,32),(33,256),(257,2048),(2049,10**9)]:
    s = (c >= lo) & (c <= hi)
    print(lo, hi, int(s.sum()), int(c[s].sum()))
