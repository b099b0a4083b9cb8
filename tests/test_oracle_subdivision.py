"""Kmeans_Subdivision (/root/reference/RepeatResolver.c:3382-3404, called by main at 4065), the caller of the scope rows 8f-3
and 8f-4: Unterteilungskomprimierung (1823-1843), Relative_Vars + Kmeans for every part of more than 2 * mingroup reads, in
order, and Unterteilungskomprimierung again.
  * the host composition of the product (repeatresolver_b200.Kmeans_Subdivision) with the oracle's Relative_Vars and Kmeans in
    place of the device calls - same signatures - against the committed output of the UNMODIFIED RepeatResolver.c
    (tests/golden/subdivision.json, oracle/gen_golden_subdivision.py); the device calls themselves are pinned one by one
    (tests/test_zz_gpu_relvars.py, tests/test_zz_gpu_kmeans.py) and composed on the GPU in tests/test_zzz_gpu_subdivision.py;
  * against the reference binary on a fresh input where oracle/_ref/ref_subdivision_driver exists."""
import json
import os
import sys

import numpy as np
import pytest

import oracle_lib as O
import repeatresolver_b200 as rr
from conftest import GOLD, ROOT, golden_msa
from test_oracle_relvars import window_codes

DRV = os.path.join(ROOT, "oracle", "_ref", "ref_subdivision_driver")


def subdivision_cases():
    with open(os.path.join(GOLD, "subdivision.json")) as f:
        return json.load(f)


def oracle_calls(o):
    """the oracle's Relative_Vars and Kmeans with the signatures of the device calls"""
    return (lambda msa, u, k, M, cutoff, mingroup: o.relative_vars(u, k, M, cutoff, mingroup),
            lambda msa, u, k, Vars, mingroup: o.kmeans(u, k, Vars, mingroup))


def case_inputs(name, case):
    codes = window_codes(golden_msa(name), case["von"], case["bis"])
    assert codes.shape == (case["rows"], case["cols"])
    o = O.Oracle.from_codes(codes)
    M, _, _ = o.scan(case["mincov"])
    M = np.array([float(l) for l in O.fmt_lines(M).split()], dtype=np.float64)   # as MaxCorrsEinlesen reads them
    return codes, o, M


def test_unterteilungskomprimierung():
    n, u = rr.Unterteilungskomprimierung([5, 5, -1, 2, 9, 2, 5, 0])
    assert n == 4 and list(u) == [0, 0, -1, 1, 2, 1, 0, 3]
    n, u = rr.Unterteilungskomprimierung([-1, -1])
    assert n == 0 and list(u) == [-1, -1]
    n, u = rr.Unterteilungskomprimierung(np.zeros(0, dtype=np.int32))
    assert n == 0 and len(u) == 0
    n, u = rr.Unterteilungskomprimierung([3, 3, 3])
    assert n == 1 and list(u) == [0, 0, 0]


@pytest.mark.parametrize("name", sorted(subdivision_cases()))
def test_kmeans_subdivision_matches_the_unmodified_reference(name):
    case = subdivision_cases()[name]
    codes, o, M = case_inputs(name, case)
    rv, km = oracle_calls(o)
    split = 0
    for mingroup, want in case["after"].items():
        n, u = rr.Kmeans_Subdivision(None, case["before"], M, case["cutoff"], int(mingroup), relative_vars=rv, kmeans=km)
        assert list(u) == want, (name, mingroup)
        assert n == len(set(x for x in want if x >= 0))
        assert [x < 0 for x in want] == [x < 0 for x in case["before"]]
        split += n > len(set(x for x in case["before"] if x >= 0))
    assert split >= 1


@pytest.mark.skipif(not os.path.exists(DRV), reason="oracle/_ref is built in the build container only")
def test_kmeans_subdivision_against_the_reference_binary_on_a_fresh_input():
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    from gen_golden_subdivision import run_driver
    g = rr.MsaGen(type="Tree", copies=5, coverage=20, repeat_len=900, diff=0.03, seed=79, flank=400, min_overlap=80)
    text = g.text()
    width = len(text.split(b"\n")[0])
    von, bis = width // 10, width - 1 - width // 10
    codes = window_codes(text, von, bis)
    o = O.Oracle.from_codes(codes)
    M, _, _ = o.scan(12)
    mtext = O.fmt_lines(M)
    M = np.array([float(l) for l in mtext.split()], dtype=np.float64)
    site = int(np.argmax(M)) // 5
    before = codes[:, site].astype(np.int32) + 3                             # part numbers with gaps: renamed first
    before[5::17] = -1
    rv, km = oracle_calls(o)
    for mingroup in (4, 9):
        R, N, want = run_driver(text, von, bis, mtext, before, 2.5, mingroup)
        assert (R, N) == codes.shape
        n, u = rr.Kmeans_Subdivision(None, before, M, 2.5, mingroup, relative_vars=rv, kmeans=km)
        assert list(u) == want, mingroup
