"""SURVEY.md section 8f row 2 ("next"): the oracle for Group_Refinement's inner step, Cliquer
(/root/reference/RepeatResolver.c:1179-1240), pinned before a GPU path for it exists.
  * the C restatement (oracle/maxcorr_oracle.c: rr_oracle_cliquer, rr_oracle_group_score) against the committed
    output of the UNMODIFIED RepeatResolver.c (tests/golden/cliquer.json, made by oracle/gen_golden_cliquer.py);
  * against the reference binary itself on a fresh input, where oracle/_ref/ref_cliquer_driver exists;
  * the score variant's relation to MaxCorrelation's (472-488 vs 421-434)."""
import gzip
import json
import os
import subprocess

import numpy as np
import pytest

import oracle_lib as O
from conftest import GOLD, ROOT, golden_msa

DRV = os.path.join(ROOT, "oracle", "_ref", "ref_cliquer_driver")
CODES = np.full(256, 5, dtype=np.uint8)
for _ch, _k in ((b"aA", 0), (b"cC", 1), (b"gG", 2), (b"tT", 3), (b"-_", 4)):
    for _c in _ch:
        CODES[_c] = _k


def window_codes(text, von, bis):
    """RepeatResolver.c's reader (293-429): rows with a symbol at both ends of the window, its columns only"""
    lines = text.split(b"\n")
    if lines and lines[-1] == b"":
        lines.pop()
    rows = [l for l in lines if l[von:von + 1] != b" " and l[bis:bis + 1] != b" "]
    return np.stack([CODES[np.frombuffer(l[von:bis + 1], dtype=np.uint8)] for l in rows])


def cliquer_cases():
    with open(os.path.join(GOLD, "cliquer.json")) as f:
        return json.load(f)


@pytest.mark.parametrize("name", sorted(cliquer_cases()))
def test_cliquer_matches_the_unmodified_reference(name):
    case = cliquer_cases()[name]
    codes = window_codes(golden_msa(name), case["von"], case["bis"])
    assert codes.shape == (case["rows"], case["cols"])
    o = O.Oracle.from_codes(codes)
    sizes = set()
    for q, want in case["queries"].items():
        members, best = o.cliquer(int(q), case["mincov"], case["maxclique"], case["greedy"])
        assert members[0] == int(q) and best[0] == 100.0
        assert list(members[1:]) == want["members"], q
        assert [float(b).hex() for b in best[1:]] == want["scores"], q      # bit for bit
        assert len(members) == want["n"] <= case["maxclique"]
        assert (np.diff(best[1:]) <= 0).all() and (best[1:] > case["greedy"]).all()
        sizes.add(len(members))
    assert 1 in sizes and max(sizes) > 1                                      # empty and non-empty cliques both pinned
    if name == "saturated":                                                   # 97.90 + F1 (486) is exercised
        top = [float.fromhex(s) for w in case["queries"].values() for s in w["scores"]]
        assert any(97.9 < z <= 98.9 for z in top)


@pytest.mark.skipif(not os.path.exists(DRV), reason="oracle/_ref/ref_cliquer_driver not built (no reference sources here)")
def test_cliquer_matches_the_reference_binary_on_fresh_input(tmp_path):
    import repeatresolver_b200 as rr
    g = rr.MsaGen(type="Tree", copies=8, coverage=40, repeat_len=1500, diff=0.01, seed=21, flank=300)
    text = g.text()
    p = tmp_path / "M"
    p.write_bytes(text)
    N = g.cols
    von, bis = N // 4, 3 * N // 4
    codes = window_codes(text, von, bis)
    o = O.Oracle.from_codes(codes)
    M, A, P = o.scan(30)
    queries = [int(q) for q in np.argsort(-M, kind="stable")[:8]] + [7, 5 * (codes.shape[1] // 2) + 4]
    out = subprocess.run([DRV, str(p), str(von), str(bis), "30", "30", "3.0"] + [str(q) for q in queries],
                         capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    lines = [l for l in out.stdout.splitlines() if l and l[0].isdigit()]
    assert tuple(int(x) for x in lines[0].split()) == codes.shape
    assert len(lines) == 1 + len(queries)
    nonempty = 0
    for l in lines[1:]:
        f = l.split()
        a, n = int(f[0]), int(f[1])
        members, best = o.cliquer(a, 30, 30, 3.0)
        assert len(members) == n
        assert list(members[1:]) == [int(x.split(":")[0]) for x in f[2:]]
        assert np.array_equal(best[1:], np.array([float(x.split(":")[1]) for x in f[2:]]))
        nonempty += n > 1
    assert nonempty >= 8


def test_group_score_is_the_scan_score_below_saturation_and_97_90_plus_f1_above():
    rng = np.random.default_rng(5)
    for _ in range(2000):
        cov = int(rng.integers(2, 3000))
        gr1 = int(rng.integers(1, cov + 1))
        gr2 = int(rng.integers(1, cov + 1))
        s = int(rng.integers(max(1, gr1 + gr2 - cov), min(gr1, gr2) + 1)) if max(1, gr1 + gr2 - cov) <= min(gr1, gr2) else 0
        if s < 1:
            continue
        zi, zj = gr1 + int(rng.integers(0, 50)), gr2 + int(rng.integers(0, 50))
        a, b = O.score(s, gr1, gr2, cov, zi, zj), O.group_score(s, gr1, gr2, cov, zi, zj)
        if a <= 98.0:
            assert a == b                                                     # 472-488 == 421-434 below saturation
        else:
            assert b == 97.90 + (a - 98.0) or abs(b - (a - 0.1)) < 1e-12      # same F1, other offset (486 vs 432)
    assert O.group_score(0, 0, 5, 10, 3, 5) == 0.0                            # 482
    # deep saturation: F1 = 2s / (|Gi| + |Gj|)
    assert O.group_score(400, 400, 400, 2000, 400, 400) == 97.90 + 1.0


# ---- the product's host half of Cliquer (rr_cliquer_from_counts) against the oracle, on the oracle's own counts ------
@pytest.mark.parametrize("name", sorted(cliquer_cases()))
def test_product_host_half_matches_the_oracle_on_given_counts(name):
    import repeatresolver_b200 as rr
    case = cliquer_cases()[name]
    codes = window_codes(golden_msa(name), case["von"], case["bis"])
    o = O.Oracle.from_codes(codes)
    gs = o.gsize()
    G = 5 * codes.shape[1]
    groups = np.arange(G, dtype=np.int32)
    for q, want in case["queries"].items():
        q = int(q)
        counts = np.array([o.counts(i, q) for i in range(G)], dtype=np.int32)
        members, scores = rr.cliquer_from_counts(q, groups, counts, gs, gs[q], case["mincov"], case["maxclique"], case["greedy"])
        assert list(members[1:]) == want["members"] and members[0] == q
        assert [float(z).hex() for z in scores[1:]] == want["scores"] and scores[0] == 100.0


def test_group_score_host_bitwise_equal_to_oracle():
    import repeatresolver_b200 as rr
    rng = np.random.default_rng(9)
    for _ in range(3000):
        cov = int(rng.integers(2, 4000))
        gr1, gr2 = int(rng.integers(1, cov + 1)), int(rng.integers(1, cov + 1))
        lo, hi = max(1, gr1 + gr2 - cov), min(gr1, gr2)
        if lo > hi:
            continue
        s = int(rng.integers(lo, hi + 1))
        zi, zj = gr1 + int(rng.integers(0, 40)), gr2 + int(rng.integers(0, 40))
        assert rr.group_score_host(s, gr1, gr2, cov, zi, zj) == O.group_score(s, gr1, gr2, cov, zi, zj)


# ---- the host half of the device batch path (rr_cliquer_from_hits): the device stage is emulated with the oracle -----
def emulated_hits(o, queries, mincov, greedy, rng, noise=2e-13):
    """what rr_k_cliquer_counts + rr_k_cliquer_score leave behind: every pair above mincov/4 whose (device) score
    exceeds greedy less a margin, in arbitrary order; the device score is the oracle's, off by a few ulp"""
    import repeatresolver_b200 as rr
    gs = o.gsize()
    hits = []
    thr = min(greedy - 1e-9 * max(1.0, abs(greedy)), 97.89)
    for slot, q in enumerate(queries):
        for i in range(len(gs)):
            if i == q:
                continue
            c = o.counts(i, q)
            if c[0] <= mincov // 4:
                continue
            z = O.group_score(c[0], c[1], c[2], c[3], gs[i], gs[q]) * (1.0 + noise * float(rng.uniform(-1, 1)))
            if z > thr:
                hits.append((slot, i, c[0], c[1], c[2], c[3], z))
    hits = np.array(hits, dtype=rr.HIT_DTYPE)
    rng.shuffle(hits)
    return hits


@pytest.mark.parametrize("name", sorted(cliquer_cases()))
def test_batch_host_half_on_emulated_device_hits(name):
    import repeatresolver_b200 as rr
    case = cliquer_cases()[name]
    codes = window_codes(golden_msa(name), case["von"], case["bis"])
    o = O.Oracle.from_codes(codes)
    queries = [int(q) for q in case["queries"]]
    hits = emulated_hits(o, queries, case["mincov"], case["greedy"], np.random.default_rng(3))
    members, scores, n = rr.cliquer_from_hits(queries, hits, o.gsize(), case["mincov"], case["maxclique"], case["greedy"])
    for k, q in enumerate(queries):
        want = case["queries"][str(q)]
        assert n[k] == want["n"] and members[k, 0] == q and scores[k, 0] == 100.0
        assert list(members[k, 1:n[k]]) == want["members"] and (members[k, n[k]:] == -1).all()
        assert [float(z).hex() for z in scores[k, 1:n[k]]] == want["scores"] and (scores[k, n[k]:] == 0).all()
    # a smaller clique than the number of hits: the cut at the weakest of the top maxclique-1 device scores
    for maxclique in (2, 3, 5):
        members, scores, n = rr.cliquer_from_hits(queries, hits, o.gsize(), case["mincov"], maxclique, case["greedy"])
        for k, q in enumerate(queries):
            m0, b0 = o.cliquer(q, case["mincov"], maxclique, case["greedy"])
            assert list(members[k, :n[k]]) == list(m0) and np.array_equal(scores[k, :n[k]], b0)


def test_batch_host_half_keeps_group_order_among_equal_scores():
    """duplicated columns give exactly equal scores: TheBestUpdater keeps the earlier group (1156-1176), also where the
    tie straddles the end of the clique and the device scores of the tied pairs differ in the last bits"""
    import repeatresolver_b200 as rr
    rng = np.random.default_rng(11)
    R, base = 120, 12
    cols = rng.integers(0, 2, size=(R, base)).astype(np.uint8)
    cols[:, 1] = cols[:, 0]                       # strongly correlated sites ...
    flip = rng.random(R) < 0.1
    cols[flip, 1] ^= 1
    codes = np.concatenate([cols, cols, cols[:, :4], cols], axis=1)      # ... several times over
    o = O.Oracle.from_codes(codes)
    G = 5 * codes.shape[1]
    queries = [0, 1, 5, 5 * base + 1, G - 5]
    hits = emulated_hits(o, queries, 8, 1.0, rng, noise=4e-13)
    for maxclique in (2, 3, 4, 6, 30):
        members, scores, n = rr.cliquer_from_hits(queries, hits, o.gsize(), 8, maxclique, 1.0)
        ties = 0
        for k, q in enumerate(queries):
            m0, b0 = o.cliquer(q, 8, maxclique, 1.0)
            assert list(members[k, :n[k]]) == list(m0) and np.array_equal(scores[k, :n[k]], b0), (maxclique, q)
            ties += int((np.diff(b0[1:]) == 0).sum())
        assert ties > 0 or maxclique == 2


def test_batch_host_half_rejects_bad_hits():
    import repeatresolver_b200 as rr
    bad = np.array([(3, 1, 5, 5, 5, 9, 4.0)], dtype=rr.HIT_DTYPE)       # slot 3 of 1 query
    with pytest.raises(rr.RRError):
        rr.cliquer_from_hits([0], bad, np.full(10, 5, dtype=np.int32))
    with pytest.raises(rr.RRError):
        rr.cliquer_from_hits([10], bad[:0], np.full(10, 5, dtype=np.int32))


def test_batch_host_half_is_immune_to_the_saturation_switch_on_the_device():
    """a device score next to 98 may come from the other formula than the host's (486: raw > 98 -> 97.90 + F1): every hit
    at or above 97.89 is re-evaluated, so whatever the device reports for those must not change the result"""
    import repeatresolver_b200 as rr
    case = cliquer_cases()["saturated"]
    codes = window_codes(golden_msa("saturated"), case["von"], case["bis"])
    o = O.Oracle.from_codes(codes)
    queries = [int(q) for q in case["queries"]]
    rng = np.random.default_rng(8)
    hits = emulated_hits(o, queries, case["mincov"], case["greedy"], rng)
    high = hits["z"] >= 97.89
    assert high.sum() >= 5
    for maxclique in (3, case["maxclique"], 40):
        want = [o.cliquer(q, case["mincov"], maxclique, case["greedy"]) for q in queries]
        for _ in range(5):
            h = hits.copy()
            h["z"][high] = rng.uniform(97.89, 98.9, size=int(high.sum()))     # arbitrary order among the saturated ones
            members, scores, n = rr.cliquer_from_hits(queries, h, o.gsize(), case["mincov"], maxclique, case["greedy"])
            for k, (m0, z0) in enumerate(want):
                assert list(members[k, :n[k]]) == list(m0) and np.array_equal(scores[k, :n[k]], z0), (maxclique, k)


@pytest.mark.parametrize("greedy", [97.0, 97.95, 98.3, 98.89])
def test_batch_host_half_with_a_threshold_inside_the_saturation_band(greedy):
    """greedy next to or inside (97.90, 98.90]: the device threshold is clamped to 97.89 so that both formulas' values reach
    the host, which applies Z > greedy itself"""
    import repeatresolver_b200 as rr
    case = cliquer_cases()["saturated"]
    codes = window_codes(golden_msa("saturated"), case["von"], case["bis"])
    o = O.Oracle.from_codes(codes)
    queries = [int(q) for q in case["queries"]]
    hits = emulated_hits(o, queries, case["mincov"], greedy, np.random.default_rng(10))
    members, scores, n = rr.cliquer_from_hits(queries, hits, o.gsize(), case["mincov"], 6, greedy)
    some = 0
    for k, q in enumerate(queries):
        m0, z0 = o.cliquer(q, case["mincov"], 6, greedy)
        assert list(members[k, :n[k]]) == list(m0) and np.array_equal(scores[k, :n[k]], z0), (greedy, q)
        some += len(m0) > 1
    assert some > 0 or greedy > 98.5
