#!/usr/bin/env python
"""TEST INFRASTRUCTURE.  Generates tests/golden/kmeans.json: what the UNMODIFIED /root/reference/RepeatResolver.c
(oracle/_ref/ref_kmeans_driver: its Einlesen 293-429 + Kmeans 2604-2821) makes of the parts and group selections of
tests/golden/relvars.json.  SURVEY.md section 8f row 4 ("next"): the fixture pins oracle/maxcorr_oracle.c:rr_oracle_kmeans
before a GPU path for it exists.  Run in the build container only (`make -C oracle && python oracle/gen_golden_kmeans.py`)."""
import gzip
import json
import os
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
DRV = os.path.join(ROOT, "oracle", "_ref", "ref_kmeans_driver")
GOLD = os.path.join(ROOT, "tests", "golden")


def run_driver(text, von, bis, unterteilung, u_no, mingroup, vars_):
    with tempfile.TemporaryDirectory() as d:
        p = os.path.join(d, "M")
        with open(p, "wb") as f:
            f.write(text)
        with open(os.path.join(d, "ut"), "w") as f:
            f.write("\n".join(str(int(x)) for x in unterteilung) + "\n")
        out = subprocess.run([DRV, p, str(von), str(bis), os.path.join(d, "ut"), str(u_no), str(mingroup)] + [str(int(v)) for v in vars_],
                             capture_output=True, text=True)
        assert out.returncode == 0, out.stderr + out.stdout
    lines = out.stdout.splitlines()
    shape = next(l for l in lines if l and l[0].isdigit())
    R, N = (int(x) for x in shape.split())
    split = int(next(l for l in lines if l.startswith("SPLIT ")).split()[1])
    parts = [int(x) for x in next(l for l in lines if l.startswith("PARTS")).split()[1:]]
    assert len(parts) == R
    return R, N, split, parts


def main():
    import oracle_lib as O
    from test_oracle_relvars import partition_by_site, relvars_cases, window_codes
    cases = {}
    for name, rel in sorted(relvars_cases().items()):
        with gzip.open(os.path.join(GOLD, name + ".msa.gz"), "rb") as f:
            text = f.read()
        codes = window_codes(text, rel["von"], rel["bis"])
        o = O.Oracle.from_codes(codes)
        M, _, _ = o.scan(rel["mincov"])
        ut, _ = partition_by_site(codes, M)
        parts = {}
        for u_no, want in sorted(rel["parts"].items()):
            if not want["vars"]:
                continue
            for mingroup in (rel["mingroup"], 2 * rel["mingroup"] + 1):
                R, N, split, after = run_driver(text, rel["von"], rel["bis"], ut, int(u_no), mingroup, want["vars"])
                assert (R, N) == codes.shape
                n0, u0 = o.kmeans(ut, int(u_no), want["vars"], mingroup)
                parts[f"{u_no}/{mingroup}"] = {"split": split, "after": after}
                print(name, "u_no", u_no, "mingroup", mingroup, "vars", len(want["vars"]), "clusters", split,
                      "restatement agrees:", n0 == split and list(u0) == after)
        cases[name] = parts
    with open(os.path.join(GOLD, "kmeans.json"), "w") as f:
        json.dump(cases, f, sort_keys=True, separators=(",", ":"))


if __name__ == "__main__":
    main()
