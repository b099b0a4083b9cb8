"""time the scan of a timing-experiment build (tools/exp_build.sh): usage: python tools/exp_probe.py NAME [WORKLOAD] [reps]"""
import os
import sys

root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
name = sys.argv[1]
sys.path.insert(0, os.path.join(root, "exp", "pkg_" + name) if name != "product" else root)
sys.path.insert(1, root)
import repeatresolver_b200 as rr  # noqa: E402
import bench  # noqa: E402

wl = sys.argv[2] if len(sys.argv) > 2 else "Tree_1perc_30000"
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
g, msa = bench.make_msa(rr, wl)
pk = rr.Packed(msa, 0)
ms = []
for _ in range(reps):
    st = pk.scan(mincov=30, variant="umma_mxf4")
    ms.append(round(st["kernel_ms"], 2))
print(name, os.path.dirname(rr.__file__), ms, "exact", st["exact_evals"], "tier2", st["bound_evals"], flush=True)
