import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import repeatresolver_b200 as rr
import bench
wl = sys.argv[1] if len(sys.argv) > 1 else "Tree_1perc_10000_25"
g, msa = bench.make_msa(rr, wl)
pk = rr.Packed(msa, 0)
ref = None
for variant, flags in [("umma_f4", 0), ("umma_f4", 0), ("umma_f4", 0x1000), ("umma", 0), ("bitset", 0), ("umma_f4", rr.FLAG_NO_PRUNE if wl != "Tree_1perc_30000" else 0)]:
    st = pk.scan(mincov=30, variant=variant, flags=flags)
    M, A = pk.fetch()
    if ref is None:
        ref = (M.copy(), A.copy())
    dm = np.nonzero(M != ref[0])[0]
    da = np.nonzero(A != ref[1])[0]
    print(variant, hex(flags), "pairs", st["pair_tests"], "M diffs", len(dm), "A diffs", len(da), flush=True)
    for gidx in dm[:5]:
        print("   M", gidx, repr(M[gidx]), repr(ref[0][gidx]), A[gidx], ref[1][gidx])
    for gidx in da[:5]:
        if M[gidx] == ref[0][gidx]:
            print("   A-only", gidx, repr(M[gidx]), A[gidx], ref[1][gidx])
