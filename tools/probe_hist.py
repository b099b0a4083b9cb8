import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import repeatresolver_b200 as rr
import bench
wl = sys.argv[1] if len(sys.argv) > 1 else "Tree_1perc_30000"
g, msa = bench.make_msa(rr, wl)
pk = rr.Packed(msa, 0)
st = pk.scan(mincov=30, variant="umma")
M, A = pk.fetch()
gs, cv = pk.sizes()
R = g.rows
q = 30 // 4
colok = (gs > q) & (gs < R)
print("groups", len(M), "colok", colok.sum(), "M>0", (M > 0).sum())
edges = [0, 1e-9, 0.01, 0.1, 0.3, 0.5, 1, 2, 3, 4, 5, 6, 8, 10, 20, 50, 98, 100]
h, _ = np.histogram(M[colok], bins=edges)
for lo, hi, n in zip(edges[:-1], edges[1:], h):
    print(f"M in [{lo},{hi}): {n}")
# group size relative to coverage for low-M groups
covg = np.repeat(cv, 5)
frac = gs / np.maximum(covg, 1)
low = colok & (M < 0.5)
print("low-M groups: size/coverage quantiles", np.quantile(frac[low], [0.01, 0.1, 0.5, 0.9, 0.99]) if low.any() else None)
print("low-M groups by base index", np.bincount(np.nonzero(low)[0] % 5, minlength=5))
mid = colok & (M >= 0.5) & (M < 3)
print("mid-M groups: size quantiles", np.quantile(gs[mid], [0.01, 0.1, 0.5, 0.9, 0.99]) if mid.any() else None, np.bincount(np.nonzero(mid)[0] % 5, minlength=5))
print({k: st[k] for k in ("kernel_ms", "exact_evals", "bound_evals", "pair_tests")})
