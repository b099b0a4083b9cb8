"""SURVEY.md section 8f row 2, second half: CliqueGroup / CliqueCoverage (/root/reference/RepeatResolver.c:976-1008,
1064-1096; called by Group_Refinement 1662-1664).
  * the numpy restatement (tests/oracle_lib.py: clique_group, clique_coverage) against the committed output of the
    UNMODIFIED RepeatResolver.c (tests/golden/cliquegroup.json, made by oracle/gen_golden_cliquegroup.py);
  * against the reference binary itself on a fresh input, where oracle/_ref/ref_cliquegroup_driver exists."""
import json
import os
import subprocess

import numpy as np
import pytest

import oracle_lib as O
from conftest import GOLD, ROOT, golden_msa
from test_oracle_cliquer import window_codes

DRV = os.path.join(ROOT, "oracle", "_ref", "ref_cliquegroup_driver")


def cliquegroup_cases():
    with open(os.path.join(GOLD, "cliquegroup.json")) as f:
        return json.load(f)


def words(hexes):
    return np.array([int(h, 16) for h in hexes], dtype=np.uint64)


@pytest.mark.parametrize("name", sorted(cliquegroup_cases()))
def test_clique_groups_match_the_unmodified_reference(name):
    case = cliquegroup_cases()[name]
    codes = window_codes(golden_msa(name), case["von"], case["bis"])
    assert codes.shape == (case["rows"], case["cols"])
    checked = nonempty = 0
    for q, rec in case["queries"].items():
        clique = rec["clique"]
        assert clique[0] == int(q)
        for c, want in rec["cutoffs"].items():
            g = O.bitset_words(O.clique_group(codes, clique, int(c)))
            v = O.bitset_words(O.clique_coverage(codes, clique, int(c)))
            assert np.array_equal(g, words(want["group"])), (name, q, c)
            assert np.array_equal(v, words(want["coverage"])), (name, q, c)
            checked += 1
            nonempty += int(g.any())
            if int(c) < 0:
                assert int(sum(bin(int(w)).count("1") for w in g)) == case["rows"]      # every read
            if int(c) >= len(clique):
                assert not g.any() and not v.any()
    assert checked >= 40 and nonempty >= 10


@pytest.mark.skipif(not os.path.exists(DRV), reason="oracle/_ref is built in the build container only")
def test_clique_groups_against_the_reference_binary_on_a_fresh_input(tmp_path):
    import repeatresolver_b200 as rr
    g = rr.MsaGen(type="Tree", copies=5, coverage=20, repeat_len=900, diff=0.03, seed=77, flank=400, min_overlap=80)
    text = g.text()
    p = tmp_path / "M"
    p.write_bytes(text)
    width = len(text.split(b"\n")[0])
    von, bis = width // 10, width - 1 - width // 10
    codes = window_codes(text, von, bis)
    o = O.Oracle.from_codes(codes)
    M, _, _ = o.scan(12)
    queries = [int(q) for q in np.argsort(-M, kind="stable")[:6]]
    cutoffs = [0, 2, 5, 9]
    out = subprocess.run([DRV, str(p), str(von), str(bis), "12", "20", "2.5", ",".join(map(str, cutoffs))] + [str(q) for q in queries],
                         capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    lines = [l for l in out.stdout.splitlines() if l and l[0].isdigit()]
    R, N, sc = (int(x) for x in lines[0].split())
    assert (R, N) == codes.shape
    n = 0
    for l in lines[1:]:
        head, gw, vw = l.split("|")
        f = head.split()
        c, size = int(f[1]), int(f[2])
        clique = [int(x) for x in f[3:3 + size]]
        assert np.array_equal(O.bitset_words(O.clique_group(codes, clique, c)), words(gw.split()))
        assert np.array_equal(O.bitset_words(O.clique_coverage(codes, clique, c)), words(vw.split()))
        n += 1
    assert n == len(queries) * len(cutoffs)
