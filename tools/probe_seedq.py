"""quality of the seeding passes: maxima left by a seed-only scan against the final ones.  usage: probe_seedq.py NAME"""
import os, sys
import numpy as np
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
name = sys.argv[1]
sys.path.insert(0, os.path.join(root, "exp", "pkg_" + name) if name != "product" else root)
sys.path.insert(1, root)
import repeatresolver_b200 as rr
import bench
g, msa = bench.make_msa(rr, "Tree_1perc_30000")
pk = rr.Packed(msa, 0)
st = pk.scan(mincov=30, variant="umma_mxf4", flags=rr.FLAG_SEED_ONLY)
Ms, _ = pk.fetch()
st2 = pk.scan(mincov=30, variant="umma_mxf4")
M, _ = pk.fetch()
pos = M > 0
print(name, "seed ms", round(st["kernel_ms"], 2), "groups with a maximum: seeded", int((Ms > 0).sum()), "final", int(pos.sum()),
      "mean seeded/final", float((Ms[pos] / M[pos]).mean()), "seeded == final", int((Ms[pos] == M[pos]).sum()),
      "full ms", round(st2["kernel_ms"], 2), "exact", st2["exact_evals"], flush=True)
