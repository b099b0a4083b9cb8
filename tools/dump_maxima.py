"""Scan a bench workload on cuda:0 and save the maxima / partners (numpy .npy) for offline analysis of pruning
thresholds.  usage: dump_maxima.py WORKLOAD OUT_PREFIX"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import repeatresolver_b200 as rr  # noqa: E402
import bench  # noqa: E402

wl, out = sys.argv[1], sys.argv[2]
g, msa = bench.make_msa(rr, wl)
pk = rr.Packed(msa, 0)
st = pk.scan(mincov=30, variant="umma_mxf4")
M, A = pk.fetch()
np.save(out + "_M.npy", M)
np.save(out + "_A.npy", A.astype(np.int32))
print(wl, g.rows, g.cols, {k: st[k] for k in ("kernel_ms", "pair_tests", "exact_evals", "bound_evals")})
