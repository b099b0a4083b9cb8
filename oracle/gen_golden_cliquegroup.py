#!/usr/bin/env python
"""TEST INFRASTRUCTURE.  Generates tests/golden/cliquegroup.json: what the UNMODIFIED /root/reference/RepeatResolver.c
(oracle/_ref/ref_cliquegroup_driver: its Einlesen 293-429, Cliquer 1179-1240, CliqueGroup 976-1008 and CliqueCoverage
1064-1096) returns for a few query groups and cutoffs of the committed golden MSAs - the second half of SURVEY.md section
8f row 2 (Group_Refinement, 1662-1664).  Run in the build container only
(`make -C oracle && python oracle/gen_golden_cliquegroup.py`)."""
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
sys.path.insert(0, os.path.join(ROOT, "oracle"))
DRV = os.path.join(ROOT, "oracle", "_ref", "ref_cliquegroup_driver")
GOLD = os.path.join(ROOT, "tests", "golden")
from gen_golden_cliquer import window_codes  # noqa: E402


def run_driver(text, von, bis, mincov, maxclique, greedy, cutoffs, queries):
    with tempfile.TemporaryDirectory() as d:
        p = os.path.join(d, "M")
        with open(p, "wb") as f:
            f.write(text)
        out = subprocess.run([DRV, p, str(von), str(bis), str(mincov), str(maxclique), repr(greedy), ",".join(str(c) for c in cutoffs)]
                             + [str(q) for q in queries], capture_output=True, text=True)
        assert out.returncode == 0, out.stderr + out.stdout
    lines = [l for l in out.stdout.splitlines() if l and (l[0].isdigit() or l[0] == "-")]
    R, N, sc = (int(x) for x in lines[0].split())
    res = {}
    for l in lines[1:]:
        head, g, v = l.split("|")
        f = head.split()
        a, c, n = int(f[0]), int(f[1]), int(f[2])
        res.setdefault(a, {"clique": [int(x) for x in f[3:3 + n]], "cutoffs": {}})
        assert res[a]["clique"] == [int(x) for x in f[3:3 + n]]
        gw, vw = g.split(), v.split()
        assert len(gw) == sc and len(vw) == sc
        res[a]["cutoffs"][str(c)] = {"group": gw, "coverage": vw}
    return R, N, sc, res


def main():
    import oracle_lib as O
    cases = {}
    for name, mincov, maxclique, greedy, frac in (("tree_small", 10, 12, 2.0, (0.1, 0.9)), ("distributed_small", 12, 30, 3.0, (0.2, 0.8)),
                                                  ("saturated", 30, 8, 5.0, (0.0, 1.0))):
        with gzip.open(os.path.join(GOLD, name + ".msa.gz"), "rb") as f:
            text = f.read()
        width = len(text.split(b"\n")[0])
        von, bis = int(frac[0] * (width - 1)), int(frac[1] * (width - 1))
        codes = window_codes(text, von, bis)
        o = O.Oracle.from_codes(codes)
        M, A, P = o.scan(mincov)
        order = np.argsort(-M, kind="stable")
        queries = [int(q) for q in order[:4]] + [int(q) for q in order[len(order) // 3: len(order) // 3 + 2]]
        cutoffs = [-1, 0, 1, 3, maxclique // 2, maxclique, 100]
        R, N, sc, res = run_driver(text, von, bis, mincov, maxclique, greedy, cutoffs, queries)
        assert (R, N) == codes.shape and sc == R // 64 + 1, ((R, N, sc), codes.shape)
        cases[name] = {"von": von, "bis": bis, "mincov": mincov, "maxclique": maxclique, "greedy": greedy, "rows": R, "cols": N, "sc": sc,
                       "queries": {str(q): res[q] for q in queries}}
        print(name, codes.shape, "queries", queries, "clique sizes", [len(res[q]["clique"]) for q in queries])
    with open(os.path.join(GOLD, "cliquegroup.json"), "w") as f:
        json.dump(cases, f, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
