"""time the scan of a timing-experiment build (tools/exp_build.sh): usage: python tools/exp_probe.py NAME [WORKLOAD] [reps] [variant] [flags]"""
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
variant = sys.argv[4] if len(sys.argv) > 4 else "umma_mxf4"
flags = int(sys.argv[5], 0) if len(sys.argv) > 5 else 0
g, msa = bench.make_msa(rr, wl)
pk = rr.Packed(msa, 0)
ms = []
for _ in range(reps):
    try:
        st = pk.scan(mincov=30, variant=variant, flags=flags)
    except rr.RRError as e:
        if flags & 0x400 and "mismatch" in str(e):   # MMA-only: no pair tests counted
            st = {"kernel_ms": float("nan"), "exact_evals": 0, "bound_evals": 0}
        else:
            raise
    ms.append(round(st["kernel_ms"], 2))
print(name, variant, hex(flags), os.path.dirname(rr.__file__), ms, "exact", st["exact_evals"], "tier2", st["bound_evals"], "units", st.get("work_units"), flush=True)
