#!/usr/bin/env python
"""TEST INFRASTRUCTURE.  Generates tests/golden/subdivision.json: the read partition the UNMODIFIED
/root/reference/RepeatResolver.c (oracle/_ref/ref_subdivision_driver: its Einlesen 293-429, MaxCorrsEinlesen 609-646 and
Kmeans_Subdivision 3382-3404 = Unterteilungskomprimierung, then Relative_Vars + Kmeans for every large part, then
Unterteilungskomprimierung again) makes of the partitions of tests/golden/relvars.json.  Run in the build container only
(`make -C oracle && python oracle/gen_golden_subdivision.py`)."""
import gzip
import json
import os
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
DRV = os.path.join(ROOT, "oracle", "_ref", "ref_subdivision_driver")
GOLD = os.path.join(ROOT, "tests", "golden")


def run_driver(text, von, bis, maxcorr_text, unterteilung, cutoff, mingroup):
    with tempfile.TemporaryDirectory() as d:
        p, m, u = os.path.join(d, "M"), os.path.join(d, "MaxCorrsOf_M"), os.path.join(d, "ut")
        with open(p, "wb") as f:
            f.write(text)
        with open(m, "wb") as f:
            f.write(maxcorr_text)
        with open(u, "w") as f:
            f.write("\n".join(str(int(x)) for x in unterteilung) + "\n")
        out = subprocess.run([DRV, p, str(von), str(bis), m, u, repr(cutoff), str(mingroup)], capture_output=True, text=True)
        assert out.returncode == 0, out.stderr + out.stdout
    lines = out.stdout.splitlines()
    k = next(i for i, l in enumerate(lines) if l.startswith("PARTS"))
    R, N = (int(x) for x in lines[k - 1].split())
    parts = [int(x) for x in lines[k].split()[1:]]
    assert len(parts) == R
    return R, N, parts


def main():
    import numpy as np
    import oracle_lib as O
    from test_oracle_relvars import partition_by_site, relvars_cases, window_codes
    cases = {}
    for name, rel in sorted(relvars_cases().items()):
        with gzip.open(os.path.join(GOLD, name + ".msa.gz"), "rb") as f:
            text = f.read()
        codes = window_codes(text, rel["von"], rel["bis"])
        o = O.Oracle.from_codes(codes)
        M, _, _ = o.scan(rel["mincov"])
        mtext = O.fmt_lines(M)
        ut, _ = partition_by_site(codes, M)
        ut = ut.copy()
        ut[::11] = -1                                                         # reads left out of the partition stay out
        runs = {}
        for mingroup in (rel["mingroup"], 2 * rel["mingroup"] + 1):
            R, N, after = run_driver(text, rel["von"], rel["bis"], mtext, ut, rel["cutoff"], mingroup)
            assert (R, N) == codes.shape
            runs[str(mingroup)] = after
            print(name, "mingroup", mingroup, "parts before", len(set(int(x) for x in ut if x >= 0)), "after", len(set(x for x in after if x >= 0)))
        cases[name] = {"von": rel["von"], "bis": rel["bis"], "mincov": rel["mincov"], "cutoff": rel["cutoff"], "rows": R, "cols": N,
                       "before": [int(x) for x in ut], "after": runs}
    with open(os.path.join(GOLD, "subdivision.json"), "w") as f:
        json.dump(cases, f, sort_keys=True, separators=(",", ":"))


if __name__ == "__main__":
    main()
