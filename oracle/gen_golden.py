#!/usr/bin/env python
"""TEST INFRASTRUCTURE.  Generates tests/golden/*: small MSAs and the MaxCorrsOf_* files the
UNMODIFIED reference program (oracle/_ref/MaxCorrelation_ref = /root/reference/MaxCorrelation.c
linked against oracle/gsl_shim.c) writes for them.  Run in the build container only
(`make -C oracle && python oracle/gen_golden.py`); the fixtures are committed because
/root/reference does not exist on the GPU box.

The reference ships no fixtures of its own, so apart from the
formula-defined case of SURVEY.md Appendix G and one MSA made by the reference's own pipeline
(real_pipeline_msareal, see main()) these are generated ones.
"""
import gzip
import json
import os
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REF = os.path.join(ROOT, "oracle", "_ref", "MaxCorrelation_ref")
GOLD = os.path.join(ROOT, "tests", "golden")


def kat_msa():
    """SURVEY.md Appendix G"""
    rows = []
    for r in range(48):
        s = ""
        for c in range(32):
            if r >= 42 and c < 6: ch = " "
            elif r < 5 and c >= 27: ch = " "
            elif c in (2, 25): ch = "C" if r < 16 else "A"
            elif c in (4, 28): ch = "G" if r % 3 == 0 else "T"
            elif c in (7, 30): ch = "-" if 10 <= r < 22 else "C"
            elif (7 * r + 3 * c) % 13 == 0: ch = "-"
            elif (r + 2 * c) % 17 == 0: ch = "T"
            else: ch = "A"
            s += ch
        rows.append(s)
    return ("\n".join(rows) + "\n").encode()


def run_ref(text, cov, threads=3):
    with tempfile.TemporaryDirectory() as d:
        with open(os.path.join(d, "M"), "wb") as f:
            f.write(text)
        out = subprocess.run([REF, "M", "-c", str(cov), "-p", str(threads)], cwd=d, capture_output=True, text=True)
        assert out.returncode == 0, out.stderr
        with open(os.path.join(d, "MaxCorrsOf_M"), "rb") as f:
            return f.read()


def main():
    from repeatresolver_b200 import MsaGen
    cases = {}
    cases["kat_appendix_g"] = (kat_msa(), [30, 20, 44])
    g = MsaGen(type="Tree", copies=3, coverage=12, repeat_len=700, diff=0.02, seed=11, flank=300, min_overlap=60)
    cases["tree_small"] = (g.text(), [10, 30])
    g = MsaGen(type="Distributed", copies=6, coverage=10, repeat_len=900, diff=0.02, seed=12, flank=400, min_overlap=80)
    cases["distributed_small"] = (g.text(), [12])
    g = MsaGen(type="EquiDistant", copies=4, coverage=150, repeat_len=120, diff=0.06, seed=13, flank=60, min_overlap=100)
    cases["saturated"] = (g.text(), [30])
    # InitialAligner-style _MSA: lower case, no spaces (every row covers every column), '_' as gap
    g = MsaGen(type="Tree", copies=2, coverage=30, repeat_len=150, diff=0.04, seed=14, flank=50, min_overlap=150)
    t = g.text().decode().lower().replace(" ", "-")
    cases["initial_aligner_style"] = (t.encode(), [30])
    # ragged input: a short line, a long line, '_' gaps, junk characters, last line without '\n'
    g = MsaGen(type="Tree", copies=2, coverage=25, repeat_len=200, diff=0.03, seed=15, flank=200, min_overlap=40)
    lines = g.text().decode().split("\n")[:-1]
    lines.insert(3, lines[3][:-5])
    lines.insert(9, lines[9] + "ACGT")
    lines[5] = lines[5].replace("-", "_")
    lines[7] = lines[7].replace("A", "N", 3).replace("C", "x", 2)
    lines.insert(12, "")
    ragged = "\n".join(lines)  # no trailing newline: the last row is dropped by the reference
    cases["ragged"] = (ragged.encode(), [20])

    # The reference's OWN pipeline output (BASELINE.json configs[0]): DataSimulator.py -c 40 -n 10 -d 1 -l 5000 -t Tree ->
    # ReadCutter -> InitialAligner -p 8 -> PW_ReAligner (a snapshot after its sixth realignment round), all built from the
    # unmodified sources under /root/reference in a scratch directory (DataSimulator.py run under python3 through a
    # scratch copy with print() calls).  DataSimulator.py takes no seed, so the MSA cannot be regenerated bit for bit:
    # the file itself is the fixture (312 reads x 18 557 columns, 53 % blanks, 28 % gaps; 7.7e7 pair tests at -c 30).
    real = os.environ.get("RR_REAL_MSAREAL", "/tmp/pipe/Tree_1perc_5000kb_MSAreal")
    committed = os.path.join(GOLD, "real_pipeline_msareal.msa.gz")
    if os.path.exists(committed):
        with gzip.open(committed, "rb") as f:
            cases["real_pipeline_msareal"] = (f.read(), [30])
    elif os.path.exists(real):
        with open(real, "rb") as f:
            cases["real_pipeline_msareal"] = (f.read(), [30])

    index = {}
    for name, (text, covs) in cases.items():
        with gzip.GzipFile(os.path.join(GOLD, name + ".msa.gz"), "wb", mtime=0) as f:
            f.write(text)
        index[name] = {}
        for cov in covs:
            out = run_ref(text, cov)
            fn = f"{name}.c{cov}.maxcorrs.gz"
            with gzip.GzipFile(os.path.join(GOLD, fn), "wb", mtime=0) as f:
                f.write(out)
            vals = [float(x) for x in out.split()]
            index[name][str(cov)] = {"file": fn, "lines": len(vals), "nonzero": sum(v > 0 for v in vals),
                                     "max": max(vals) if vals else 0.0}
            print(name, cov, index[name][str(cov)])
    with open(os.path.join(GOLD, "index.json"), "w") as f:
        json.dump(index, f, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
