"""BASELINE.json configs[4] shape: ~40k reads, 100 kbp repeat (rows x columns ~ 4e4 x 4.4e5)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import repeatresolver_b200 as rr
copies = int(sys.argv[1]) if len(sys.argv) > 1 else 92
mem_gb = os.sysconf("SC_PAGE_SIZE") * os.sysconf("SC_PHYS_PAGES") / 1e9
print("host memory GB", round(mem_gb), "cores", os.cpu_count(), flush=True)
t0 = time.time()
g = rr.MsaGen(type="Tree", copies=copies, coverage=40, repeat_len=100000, diff=0.01, seed=1005, threads=16)
print("rows", g.rows, "cols", g.cols, "cells GB", g.rows * g.cols / 1e9, "gen plan s", round(time.time() - t0, 1), flush=True)
if g.rows * g.cols / 1e9 > 0.45 * mem_gb:
    raise SystemExit("not enough host memory for the cell matrix")
msa = rr.MSA.alloc(g.rows, g.cols, codes=True)
g.codes(out=msa.cells())
print("generated s", round(time.time() - t0, 1), flush=True)
t1 = time.time()
pk = rr.Packed(msa, 0)
print("pack s", round(time.time() - t1, 2), flush=True)
for rep in range(2):
    t2 = time.time()
    st = pk.scan(mincov=30, variant="auto")
    print("scan", {k: st[k] for k in ("kernel_ms", "prepare_ms", "pair_tests", "exact_evals", "executed_ops", "row_sites", "variant")},
          "wall s", round(time.time() - t2, 2), flush=True)
M, A = pk.fetch()
print("pair tests/s %.3e" % (st["pair_tests"] / (st["kernel_ms"] * 1e-3)), "algorithmic int8-rate TOP/s %.0f" % (2.0 * g.rows * st["pair_tests"] / (st["kernel_ms"] * 1e-3) / 1e12))
print("M>0", int((M > 0).sum()), "of", len(M), "saturated", int((M > 98).sum()))
# the oracle spot check of this shape lives in tests/test_gpu_scaleup.py (opt-in: RR_RUN_SCALEUP=1)
