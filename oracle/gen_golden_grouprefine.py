#!/usr/bin/env python
"""TEST INFRASTRUCTURE.  Generates tests/golden/grouprefine.json: what the UNMODIFIED /root/reference/RepeatResolver.c
(oracle/_ref/ref_grouprefine_driver: its Einlesen 293-429, MaxCorrsEinlesen 609-646 and Group_Refinement 1634-1693 - Cliquer,
Dropoff_Cutoff, CliqueGroup, CliqueCoverage, GroupPrecision) leaves in Sizes / Cutoffs / Drop_Off / Cliques / C_Groups /
C_Coverage / MaxCorrs for windows of the committed golden MSAs - SURVEY.md section 8f row 2 as a whole.  The same call through
Parallel_Group_Refinement (1770-1821), which keeps cutoff and greedy in an int array (1793-1794), is checked here to equal the
serial form at the truncated values.  Run in the build container only
(`make -C oracle && python oracle/gen_golden_grouprefine.py`)."""
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
DRV = os.path.join(ROOT, "oracle", "_ref", "ref_grouprefine_driver")
GOLD = os.path.join(ROOT, "tests", "golden")
from gen_golden_cliquer import window_codes  # noqa: E402


def run_driver(text, von, bis, maxcorr_text, cutoff, mincov, maxclique, greedy, threads=0):
    """-> (R, N, sc, {group: record}, [(maj, min), ...] in the order the reference printed them)"""
    with tempfile.TemporaryDirectory() as d:
        p, m = os.path.join(d, "M"), os.path.join(d, "MaxCorrsOf_M")
        with open(p, "wb") as f:
            f.write(text)
        with open(m, "wb") as f:
            f.write(maxcorr_text)
        out = subprocess.run([DRV, p, str(von), str(bis), m, repr(cutoff), str(mincov), str(maxclique), repr(greedy), str(threads)],
                             capture_output=True, text=True)
        assert out.returncode == 0, out.stderr + out.stdout
    prec = [tuple(int(x) for x in l.split()[2::2]) for l in out.stdout.splitlines() if l.startswith("Group Precision")]
    lines = out.stdout.splitlines()
    first = next(k for k, l in enumerate(lines) if l.startswith("GR ")) - 1 if any(l.startswith("GR ") for l in lines) else len(lines) - 1
    R, N, sc = (int(x) for x in lines[first].split())
    res = {}
    for l in lines[first + 1:]:
        head, mem, g, v = l.split("|")
        f = head.split()
        assert f[0] == "GR"
        res[int(f[1])] = {"size": int(f[2]), "cutoff": int(f[3]), "drop_off": float.fromhex(f[4]).hex(), "maxcorr": float.fromhex(f[5]).hex(),
                          "clique": [int(x) for x in mem.split()], "group": g.split(), "coverage": v.split()}
    return R, N, sc, res, prec


def main():
    import oracle_lib as O
    cases = {}
    for name, mincov, maxclique, greedy, cutoff, frac in (("tree_small", 10, 12, 2.5, 4.5, (0.1, 0.9)),
                                                          ("distributed_small", 12, 30, 3.0, 3.0, (0.2, 0.8)),
                                                          ("saturated", 30, 8, 5.5, 6.25, (0.0, 1.0))):
        with gzip.open(os.path.join(GOLD, name + ".msa.gz"), "rb") as f:
            text = f.read()
        width = len(text.split(b"\n")[0])
        von, bis = int(frac[0] * (width - 1)), int(frac[1] * (width - 1))
        codes = window_codes(text, von, bis)
        o = O.Oracle.from_codes(codes)
        M, A, P = o.scan(mincov)
        mtext = O.fmt_lines(M)
        R, N, sc, res, prec = run_driver(text, von, bis, mtext, cutoff, mincov, maxclique, greedy)
        assert (R, N) == codes.shape and sc == R // 64 + 1, ((R, N, sc), codes.shape)
        refined = [i for i in sorted(res) if res[i]["size"] > 5]
        assert len(prec) == 2 * len(refined), (len(prec), len(refined))
        for k, i in enumerate(refined):                                  # 1678-1679: Groups[i], then C_Groups[i]
            res[i]["precision"] = [list(prec[2 * k]), list(prec[2 * k + 1])]
        # the parallel form computes the same at the truncated cutoff and greedy (1793-1794)
        par = run_driver(text, von, bis, mtext, cutoff, mincov, maxclique, greedy, threads=3)
        ser = run_driver(text, von, bis, mtext, float(int(cutoff)), mincov, maxclique, float(int(greedy)))
        assert par[3] == ser[3] and sorted(par[4]) == sorted(ser[4]), name
        cases[name] = {"von": von, "bis": bis, "mincov": mincov, "maxclique": maxclique, "greedy": greedy, "cutoff": cutoff,
                       "rows": R, "cols": N, "sc": sc, "maxcorrs_text_sha": __import__("hashlib").sha256(mtext).hexdigest(),
                       "groups": {str(i): res[i] for i in sorted(res)}}
        print(name, codes.shape, "groups above the cutoff", len(res), "refined", len(refined),
              "cutoffs", sorted(set(res[i]["cutoff"] for i in refined)), "parallel form:", len(par[3]), "groups")
    with open(os.path.join(GOLD, "grouprefine.json"), "w") as f:
        json.dump(cases, f, indent=None, sort_keys=True, separators=(",", ":"))


DEEP = dict(gen=dict(type="Tree", copies=80, coverage=40, repeat_len=1500, diff=0.03, seed=5, flank=500, min_overlap=100),
            frac=(0.3, 0.7), mincov=30, maxclique=30, greedy=3.0, cutoff=7.5, every=50)


def deep_inputs():
    """the deep case: a window of a generated MSA that > 3 000 reads span (several 32-word chunks per bitset on the device),
    MaxCorrs = 50 for every `every`-th group of a plausible size and for a few others, 0 elsewhere -> (text, von, bis, codes, M)"""
    import numpy as np
    import repeatresolver_b200 as rr
    text = rr.MsaGen(**DEEP["gen"]).text()
    width = len(text.split(b"\n")[0])
    von, bis = int(DEEP["frac"][0] * (width - 1)), int(DEEP["frac"][1] * (width - 1))
    codes = window_codes(text, von, bis)
    R, N = codes.shape
    sizes = np.stack([(codes == k).sum(axis=0) for k in range(5)], axis=1).ravel()          # group 5 * site + k
    M = np.zeros(5 * N)
    cand = np.flatnonzero((sizes > 30) & (sizes < R // 3))
    M[cand[::DEEP["every"]]] = 50.0
    # some groups without partners as well - but not group 0: for a query without partners the reference's trimming loop reads
    # Clique[-1] (1231), harmless unless the query is 0, which then matches the zero word in front of the array and sends the
    # loop below the array (a crash in this build)
    M[[7, 5 * N - 1]] = 50.0
    M[np.random.default_rng(8).integers(1, 5 * N, 12)] = 50.0
    return text, von, bis, codes, M


def main_deep():
    import oracle_lib as O
    text, von, bis, codes, M = deep_inputs()
    R, N, sc, res, prec = run_driver(text, von, bis, O.fmt_lines(M), DEEP["cutoff"], DEEP["mincov"], DEEP["maxclique"], DEEP["greedy"])
    assert (R, N) == codes.shape and sc == R // 64 + 1
    refined = [i for i in sorted(res) if res[i]["size"] > 5]
    assert len(prec) == 2 * len(refined)
    for k, i in enumerate(refined):
        res[i]["precision"] = [list(prec[2 * k]), list(prec[2 * k + 1])]
    case = {"rows": R, "cols": N, "sc": sc, "von": von, "bis": bis, "groups": {str(i): res[i] for i in sorted(res)}}
    print("deep", codes.shape, "groups above the cutoff", len(res), "refined", len(refined), "cutoffs", sorted(set(res[i]["cutoff"] for i in refined)))
    with open(os.path.join(GOLD, "grouprefine_deep.json"), "w") as f:
        json.dump(case, f, indent=None, sort_keys=True, separators=(",", ":"))


if __name__ == "__main__":
    main()
    main_deep()
