"""Timing probe for rr_cliquer_batch: a generated MSA, the queries a Group_Refinement pass
would issue (minor groups of a plausible size), device time of the two kernels and wall time of the whole call.
Usage: python tools/probe_cliquer.py [copies] [repeat_len] [n_queries] [out.json] [reps]
(`ncu -k regex:rr_k_cliquer_counts -c 1` with reps = 1 captures the count kernel.)
The figures it prints: candidate pairs per second (one pair = one Schnitt of RepeatResolver.c:1213) and the bytes of
bitsets one launch has to stream at least once (6 bitsets per candidate site), i.e. the HBM floor of the count kernel."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import repeatresolver_b200 as rr  # noqa: E402


def main():
    copies = int(sys.argv[1]) if len(sys.argv) > 1 else 100
    repeat_len = int(sys.argv[2]) if len(sys.argv) > 2 else 6000
    nq = int(sys.argv[3]) if len(sys.argv) > 3 else 1024
    out = sys.argv[4] if len(sys.argv) > 4 else None
    reps = int(sys.argv[5]) if len(sys.argv) > 5 else 3
    t = time.time()
    g = rr.MsaGen(type="Tree", copies=copies, coverage=40, repeat_len=repeat_len, diff=0.01, seed=1002, flank=1000)
    codes = g.codes()
    t_gen = time.time() - t
    pk = rr.Packed(rr.MSA.from_cells(codes, codes=True), 0)
    gs, _ = pk.sizes()
    cand = np.flatnonzero((gs > 30) & (gs < codes.shape[0] // 3))
    queries = cand[::max(1, len(cand) // nq)][:nq].astype(np.int32)
    res = {"rows": int(codes.shape[0]), "cols": int(codes.shape[1]), "queries": int(len(queries)), "gen_s": round(t_gen, 2)}
    w32 = 4 * ((codes.shape[0] + 127) // 128)
    res["bitset_bytes_one_pass"] = int(6 * codes.shape[1] * w32 * 4)
    for kernel in ("",):
        runs = []
        for rep in range(reps):
            t = time.time()
            members, scores, n, st = pk.cliquer_batch(queries, 30, 30, 3.0)
            runs.append(dict(st, wall_ms=round((time.time() - t) * 1e3, 2), mean_clique=float(n.mean())))
        st = runs[-1]
        res["kernel"] = {"last_run": st, "kernel_ms_all": [round(r["kernel_ms"], 3) for r in runs],
                                  "pairs_per_s_kernel": st["pairs"] / (st["kernel_ms"] * 1e-3),
                                  "pairs_per_s_call": st["pairs"] / (st["wall_ms"] * 1e-3),
                                  "us_per_query_kernel": st["kernel_ms"] * 1e3 / len(queries),
                                  "checksum": int(members.astype(np.int64).sum()), "score_sum": float(scores.sum())}
    line = json.dumps(res)
    print(line)
    if out:
        with open(out, "w") as f:
            f.write(line + "\n")
    pk.close()


if __name__ == "__main__":
    main()
