#!/usr/bin/env python
"""TEST INFRASTRUCTURE.  Generates tests/golden/relvars.json: what the UNMODIFIED /root/reference/RepeatResolver.c
(oracle/_ref/ref_relvars_driver: its Einlesen 293-429 + Relative_Vars 2424-2493 + Relative_Group_Significance 506-523,
linked against oracle/gsl_shim.c) returns on windows of the committed golden MSAs.  SURVEY.md section 8f row 3 ("next"):
the fixture pins oracle/maxcorr_oracle.c:rr_oracle_relative_vars before a GPU path for it exists.
Run in the build container only (`make -C oracle && python oracle/gen_golden_relvars.py`)."""
import gzip
import json
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
DRV = os.path.join(ROOT, "oracle", "_ref", "ref_relvars_driver")
GOLD = os.path.join(ROOT, "tests", "golden")


def run_driver(text, von, bis, maxcorrs, unterteilung, u_no, cutoff, mingroup, pairs=()):
    with tempfile.TemporaryDirectory() as d:
        p = os.path.join(d, "M")
        with open(p, "wb") as f:
            f.write(text)
        with open(os.path.join(d, "mc"), "w") as f:
            f.write("\n".join("%.17g" % x for x in maxcorrs) + "\n")
        with open(os.path.join(d, "ut"), "w") as f:
            f.write("\n".join(str(int(x)) for x in unterteilung) + "\n")
        out = subprocess.run([DRV, p, str(von), str(bis), os.path.join(d, "mc"), os.path.join(d, "ut"), str(u_no), repr(cutoff),
                              str(mingroup)] + [f"{i}:{j}" for i, j in pairs], capture_output=True, text=True)
        assert out.returncode == 0, out.stderr + out.stdout
    lines = out.stdout.splitlines()
    shape = next(l for l in lines if l and l[0].isdigit())
    R, N = (int(x) for x in shape.split())
    v = next(l for l in lines if l.startswith("VARS ")).split()
    vars_ = [int(x) for x in v[2:]]
    assert len(vars_) == int(v[1])
    scores = {}
    for l in lines:
        if l.startswith("PAIR "):
            f = l.split()
            scores[f"{f[1]}:{f[2]}"] = float(f[3]).hex()
    return R, N, vars_, scores


def main():
    import oracle_lib as O
    from test_oracle_relvars import partition_by_site, window_codes
    cases = {}
    for name, mincov, cutoff, mingroup, frac in (("tree_small", 10, 2.0, 3, (0.1, 0.9)), ("distributed_small", 12, 2.5, 4, (0.2, 0.8)),
                                                 ("saturated", 30, 5.0, 10, (0.0, 1.0))):
        with gzip.open(os.path.join(GOLD, name + ".msa.gz"), "rb") as f:
            text = f.read()
        width = len(text.split(b"\n")[0])
        von, bis = int(frac[0] * (width - 1)), int(frac[1] * (width - 1))
        codes = window_codes(text, von, bis)
        o = O.Oracle.from_codes(codes)
        M, A, P = o.scan(mincov)
        ut, site = partition_by_site(codes, M)
        parts = {}
        for u_no in sorted(set(int(x) for x in ut)):
            if (ut == u_no).sum() < 2 * mingroup:
                continue
            mine = o.relative_vars(ut, u_no, M, cutoff, mingroup)
            pairs = [(int(mine[a]), int(mine[b])) for a in range(0, min(len(mine), 12), 3) for b in range(a + 1, len(mine), 7)
                     if mine[b] >= mine[a] + 100][:12]
            R, N, vars_, scores = run_driver(text, von, bis, M, ut, u_no, cutoff, mingroup, pairs)
            assert (R, N) == codes.shape, ((R, N), codes.shape)
            parts[str(u_no)] = {"vars": vars_, "pair_scores": scores, "part_size": int((ut == u_no).sum())}
            print(name, codes.shape, "site", site, "u_no", u_no, "size", int((ut == u_no).sum()), "vars", len(vars_),
                  "restatement agrees:", list(mine) == vars_)
        cases[name] = {"von": von, "bis": bis, "mincov": mincov, "cutoff": cutoff, "mingroup": mingroup, "rows": int(codes.shape[0]),
                       "cols": int(codes.shape[1]), "partition_site": int(site), "parts": parts}
    with open(os.path.join(GOLD, "relvars.json"), "w") as f:
        json.dump(cases, f, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
