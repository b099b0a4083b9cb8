"""Timing probe for the rows after Cliquer on a generated MSA: Relative_Vars with the part's rows
packed as an MSA of their own (rr_relative_vars) and as a mask over the packed whole MSA (rr_relative_vars_packed), which must
agree, and Kmeans (csrc/rr_kmeans.cu) on the groups it selects.  The read partition is the reads' symbol at the most
significant site (MaxCorrs from a scan on the same GPU).  Wall times of the C-ABI calls.
Usage: python tools/probe_rows.py [copies] [repeat_len] [out.json]"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import repeatresolver_b200 as rr  # noqa: E402


def main():
    copies = int(sys.argv[1]) if len(sys.argv) > 1 else 40
    repeat_len = int(sys.argv[2]) if len(sys.argv) > 2 else 4000
    out = sys.argv[3] if len(sys.argv) > 3 else None
    g = rr.MsaGen(type="Tree", copies=copies, coverage=40, repeat_len=repeat_len, diff=0.01, seed=1006, flank=800)
    codes = g.codes()
    msa = rr.MSA.from_cells(codes, codes=True)
    M, A, st = rr.Parallel_AllMaxCorrsRechner(msa, 30, 1)
    site = int(np.argmax(M)) // 5
    ut = codes[:, site].astype(np.int32)
    pk = rr.Packed(msa, 0)
    res = {"rows": int(codes.shape[0]), "cols": int(codes.shape[1]), "scan_kernel_ms": st["kernel_ms"], "partition_site": site, "parts": {}}
    for u_no in sorted(set(int(x) for x in ut)):
        size = int((ut == u_no).sum())
        if size < 40:
            continue
        part = {"reads": size}
        t = time.time()
        v0 = rr.Relative_Vars(msa, ut, u_no, M, 3.0, 8)
        part["relvars_packed_per_part_ms"] = round((time.time() - t) * 1e3, 2)
        v1 = v0
        t = time.time()
        v2 = pk.relative_vars(ut, u_no, M, 3.0, 8)
        part["relvars_masked_ms"] = round((time.time() - t) * 1e3, 2)
        part["vars"] = int(len(v0))
        part["paths_agree"] = bool(np.array_equal(v0, v1) and np.array_equal(v0, v2))
        if len(v0):
            t = time.time()
            n, after = rr.Kmeans(msa, ut, u_no, v0, 8)
            part["kmeans_ms"] = round((time.time() - t) * 1e3, 2)
            part["clusters"] = int(n)
        res["parts"][str(u_no)] = part
    line = json.dumps(res)
    print(line)
    if out:
        with open(out, "w") as f:
            f.write(line + "\n")


if __name__ == "__main__":
    main()
