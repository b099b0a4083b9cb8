#!/usr/bin/env python
"""TEST INFRASTRUCTURE.  Generates tests/golden/cliquer.json: what the UNMODIFIED /root/reference/RepeatResolver.c
(oracle/_ref/ref_cliquer_driver: its Einlesen 293-429 + Cliquer 1179-1240 + Group_PositiveSignificance 472-488, linked
against oracle/gsl_shim.c) returns for a few query groups of the committed golden MSAs.  SURVEY.md section 8f row 2
("next"): the fixture pins oracle/maxcorr_oracle.c:rr_oracle_cliquer before a GPU path for it exists.
Run in the build container only (`make -C oracle && python oracle/gen_golden_cliquer.py`)."""
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
DRV = os.path.join(ROOT, "oracle", "_ref", "ref_cliquer_driver")
GOLD = os.path.join(ROOT, "tests", "golden")


def run_driver(text, von, bis, mincov, maxclique, greedy, queries):
    with tempfile.TemporaryDirectory() as d:
        p = os.path.join(d, "M")
        with open(p, "wb") as f:
            f.write(text)
        out = subprocess.run([DRV, p, str(von), str(bis), str(mincov), str(maxclique), repr(greedy)] + [str(q) for q in queries],
                             capture_output=True, text=True)
        assert out.returncode == 0, out.stderr + out.stdout
    lines = [l for l in out.stdout.splitlines() if l and (l[0].isdigit() or l[0] == "-")]
    R, N = (int(x) for x in lines[0].split())
    res = {}
    for l in lines[1:]:
        f = l.split()
        res[int(f[0])] = {"n": int(f[1]), "members": [int(x.split(":")[0]) for x in f[2:]],
                          "scores": [float(x.split(":")[1]).hex() for x in f[2:]]}
    return R, N, res


def window_codes(text, von, bis):
    """the reader's row rule (RepeatResolver.c:330): a symbol (not a blank) at both ends of the window"""
    import oracle_lib as O  # noqa: F401  (only for the shared code table below)
    lines = text.split(b"\n")
    if lines and lines[-1] == b"":
        lines.pop()
    rows = [l for l in lines if l[von:von + 1] != b" " and l[bis:bis + 1] != b" "]
    table = np.full(256, 5, dtype=np.uint8)
    for ch, k in ((b"aA", 0), (b"cC", 1), (b"gG", 2), (b"tT", 3), (b"-_", 4)):
        for c in ch:
            table[c] = k
    return np.stack([table[np.frombuffer(l[von:bis + 1], dtype=np.uint8)] for l in rows])


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
        queries = [int(q) for q in order[:5]] + [int(q) for q in order[len(order) // 3: len(order) // 3 + 2]]
        R, N, res = run_driver(text, von, bis, mincov, maxclique, greedy, queries)
        assert (R, N) == codes.shape, ((R, N), codes.shape)
        cases[name] = {"von": von, "bis": bis, "mincov": mincov, "maxclique": maxclique, "greedy": greedy, "rows": R, "cols": N,
                       "queries": {str(q): res[q] for q in queries}}
        print(name, codes.shape, "queries", queries, "sizes", [res[q]["n"] for q in queries])
    with open(os.path.join(GOLD, "cliquer.json"), "w") as f:
        json.dump(cases, f, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
