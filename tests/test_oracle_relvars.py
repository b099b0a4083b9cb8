"""SURVEY.md section 8f row 3 ("next"): the oracle for Relative_Vars (/root/reference/RepeatResolver.c:2424-2493), the
selection of the groups that vary inside one part of the read partition - an all-pairs two-sided hypergeometric test
among the selected groups restricted to the part's reads (Triple_Schnitt 150-161, Relative_Group_Significance 506-523,
CumHypGeo_Log 490-504).  Pinned before a GPU path for it exists:
  * the C restatement (oracle/maxcorr_oracle.c: rr_oracle_relative_vars, rr_oracle_relative_score) against the committed
    output of the UNMODIFIED RepeatResolver.c (tests/golden/relvars.json, made by oracle/gen_golden_relvars.py);
  * against the reference binary itself on a fresh input, where oracle/_ref/ref_relvars_driver exists;
  * the two-sided score against scipy's hypergeometric tails."""
import json
import os

import numpy as np
import pytest

import oracle_lib as O
from conftest import GOLD, ROOT, golden_msa
from test_oracle_cliquer import window_codes

DRV = os.path.join(ROOT, "oracle", "_ref", "ref_relvars_driver")


def partition_by_site(codes, M):
    """a read partition with structure: the reads by their symbol at the site of the most significant group
    (0..4, 5 = not covered there)"""
    site = int(np.argmax(M)) // 5
    return codes[:, site].astype(np.int32), site


def relvars_cases():
    path = os.path.join(GOLD, "relvars.json")
    if not os.path.exists(path):                       # only while oracle/gen_golden_relvars.py (which imports this module) writes it
        return {}
    with open(path) as f:
        return json.load(f)


def test_fixture_is_committed():
    assert sorted(relvars_cases()) == ["distributed_small", "saturated", "tree_small"]


@pytest.mark.parametrize("name", sorted(relvars_cases()))
def test_relative_vars_matches_the_unmodified_reference(name):
    case = relvars_cases()[name]
    codes = window_codes(golden_msa(name), case["von"], case["bis"])
    assert codes.shape == (case["rows"], case["cols"])
    o = O.Oracle.from_codes(codes)
    M, _, _ = o.scan(case["mincov"])
    ut, site = partition_by_site(codes, M)
    assert site == case["partition_site"]
    gs = o.gsize()
    nonempty = 0
    for u_no, want in case["parts"].items():
        u_no = int(u_no)
        assert int((ut == u_no).sum()) == want["part_size"]
        got = o.relative_vars(ut, u_no, M, case["cutoff"], case["mingroup"])
        assert list(got) == want["vars"], u_no
        nonempty += len(got) > 0
        # the scores of named pairs, bit for bit: counts by numpy, score by the restatement
        U = ut == u_no
        for key, hexz in want["pair_scores"].items():
            i, j = (int(x) for x in key.split(":"))
            gi, gj = codes[:, i // 5] == i % 5, codes[:, j // 5] == j % 5
            z = O.relative_score(int((gi & gj & U).sum()), int((gj & U).sum()), int((gi & U).sum()), int(U.sum()))
            assert float(z).hex() == hexz, key
        assert (gs[got] >= case["mingroup"]).all() and (M[got] > case["cutoff"]).all()
    assert nonempty > 0


@pytest.mark.skipif(not os.path.exists(DRV), reason="oracle/_ref/ref_relvars_driver not built (no reference sources here)")
def test_relative_vars_matches_the_reference_binary_on_fresh_input():
    import sys
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import repeatresolver_b200 as rr
    from gen_golden_relvars import run_driver
    g = rr.MsaGen(type="Tree", copies=8, coverage=40, repeat_len=1500, diff=0.01, seed=23, flank=300)
    text = g.text()
    N = g.cols
    von, bis = N // 4, 3 * N // 4
    codes = window_codes(text, von, bis)
    o = O.Oracle.from_codes(codes)
    M, _, _ = o.scan(30)
    ut, _ = partition_by_site(codes, M)
    checked = 0
    for u_no in sorted(set(int(x) for x in ut)):
        if (ut == u_no).sum() < 16:
            continue
        mine = o.relative_vars(ut, u_no, M, 3.0, 8)
        R, Nw, vars_, _ = run_driver(text, von, bis, M, ut, u_no, 3.0, 8)
        assert (R, Nw) == codes.shape and list(mine) == vars_
        checked += len(vars_) > 0
    assert checked >= 1


def test_relative_score_is_the_smaller_tail():
    from scipy.stats import hypergeom
    rng = np.random.default_rng(17)
    lower = upper = 0
    for _ in range(1500):
        cov = int(rng.integers(2, 2500))
        gr1, gr2 = int(rng.integers(1, cov + 1)), int(rng.integers(1, cov + 1))
        lo, hi = max(0, gr1 + gr2 - cov), min(gr1, gr2)
        s = int(rng.integers(lo, hi + 1))
        z = O.relative_score(s, gr1, gr2, cov)
        P = hypergeom.cdf(s, cov, gr2, gr1)           # P[X <= s]
        Q = hypergeom.sf(s - 1, cov, gr2, gr1)        # P[X >= s]
        with np.errstate(divide="ignore"):
            want = -np.log10(P if (P < Q or s == 0) else Q)
        want = 99.0 if (np.isinf(want) or want > 99) else want
        assert abs(z - want) <= 1e-7 * max(1.0, abs(want)) + 1e-9, (s, gr1, gr2, cov, z, want)
        lower += P < Q
        upper += Q <= P
    assert lower > 100 and upper > 100
    assert O.relative_score(0, 0, 5, 10) == 0.0 and O.relative_score(0, 5, 0, 10) == 0.0     # 517
    # anti-correlation counts: two groups that never share a read inside the part are significant
    assert O.relative_score(0, 40, 40, 100) > 10.0
    assert O.relative_score(40, 40, 40, 100) > 20.0
    # schnitt - 1 wraps for schnitt = 0 (493): the upper tail is then 0 and the lower tail is used
    assert O.hyper_Q(0xFFFFFFFF, 40, 60, 40) == 0.0


# ---- the product's host half (rr_relative_vars_from_counts, rr_relative_score_host) against the oracle ------------------
def test_relative_score_host_bitwise_equal_to_oracle():
    import repeatresolver_b200 as rr
    rng = np.random.default_rng(19)
    for _ in range(4000):
        cov = int(rng.integers(1, 4000))
        gr1, gr2 = int(rng.integers(0, cov + 1)), int(rng.integers(0, cov + 1))
        lo, hi = max(0, gr1 + gr2 - cov), min(gr1, gr2)
        s = int(rng.integers(lo, hi + 1))
        assert rr.relative_score_host(s, gr1, gr2, cov) == O.relative_score(s, gr1, gr2, cov), (s, gr1, gr2, cov)


def group_matrix(codes):
    """X[r][5*site+k] = read r has symbol k at the site"""
    R, N = codes.shape
    X = np.zeros((R, 5 * N), dtype=np.int32)
    for k in range(5):
        X[:, k::5] = codes == k
    return X


@pytest.mark.parametrize("name", sorted(relvars_cases()))
def test_product_host_half_matches_the_reference_on_given_counts(name):
    import repeatresolver_b200 as rr
    case = relvars_cases()[name]
    codes = window_codes(golden_msa(name), case["von"], case["bis"])
    o = O.Oracle.from_codes(codes)
    M, _, _ = o.scan(case["mincov"])
    ut, _ = partition_by_site(codes, M)
    X = group_matrix(codes)
    for u_no, want in case["parts"].items():
        XU = X[ut == int(u_no)]
        gsize_u = XU.sum(0)
        sel = rr.relative_vars_from_counts(M, gsize_u, len(XU), case["cutoff"], case["mingroup"])
        assert (M[sel] > case["cutoff"]).all() and (gsize_u[sel] >= case["mingroup"]).all() and (np.diff(sel) > 0).all()
        got = rr.relative_vars_from_counts(M, gsize_u, len(XU), case["cutoff"], case["mingroup"],
                                           lambda s: XU[:, s].T @ XU[:, s])
        assert list(got) == want["vars"], u_no


def test_product_host_half_argument_errors_and_empty_part():
    import repeatresolver_b200 as rr
    M = np.full(20, 5.0)
    gu = np.full(20, 3, dtype=np.int32)
    with pytest.raises(rr.RRError):
        rr.relative_vars_from_counts(M, gu, 10, 3.0, 0)                  # mingroup >= 1
    with pytest.raises(rr.RRError):
        rr.relative_vars_from_counts(M, gu, 10, -1.0, 2)                 # cutoff >= 0
    assert len(rr.relative_vars_from_counts(M, gu, 10, 3.0, 4)) == 0     # no group holds mingroup reads of the part
    assert len(rr.relative_vars_from_counts(M, gu, 10, 5.0, 2)) == 0     # MaxCorrs > cutoff is strict (2432)
    assert list(rr.relative_vars_from_counts(M, gu, 10, 3.0, 2)) == list(range(20))
