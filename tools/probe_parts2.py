"""what a part of an N-way split costs with the thresholds it gets (max of all parts' seeding passes) against the best it
could get (the final maxima of a whole scan).  usage: probe_parts2.py [parts]"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import repeatresolver_b200 as rr
import bench
parts = int(sys.argv[1]) if len(sys.argv) > 1 else 8
g, msa = bench.make_msa(rr, "Tree_1perc_30000")
pk = rr.Packed(msa, 0)
st = pk.scan(mincov=30)
st = pk.scan(mincov=30)
Mfull, _ = pk.fetch()
print("whole scan", round(st["kernel_ms"], 2), st["exact_evals"], flush=True)
seeds = []
for q in range(parts):
    pk.scan(mincov=30, part_index=q, part_count=parts, flags=rr.FLAG_SEED_ONLY)
    seeds.append(pk.fetch()[0])
thr = np.maximum.reduce(seeds)
for p in range(parts):
    out = []
    for name, t in (("seeded", thr), ("final", Mfull), ("seeded", thr)):
        pk.scan(mincov=30, part_index=p, part_count=parts, flags=rr.FLAG_SEED_ONLY)
        pk.set_thresholds(t)
        s = pk.scan(mincov=30, part_index=p, part_count=parts, flags=rr.FLAG_SKIP_SEED)
        out.append((name, round(s["kernel_ms"], 2), s["exact_evals"]))
    print("part", p, out, flush=True)
