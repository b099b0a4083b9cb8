"""SURVEY.md section 8f row 4 ("next"): the oracle for Kmeans (/root/reference/RepeatResolver.c:2604-2821), the split of
one part of the read partition by the reads' signatures over the groups Relative_Vars selected (read x read GrMatch
similarity 163-175, majority-of-five centroids, best-centroid assignment, dissolution of small clusters).  Pinned before
a GPU path for it exists:
  * the C restatement (oracle/maxcorr_oracle.c: rr_oracle_kmeans) against the committed output of the UNMODIFIED
    RepeatResolver.c (tests/golden/kmeans.json, made by oracle/gen_golden_kmeans.py);
  * against the reference binary itself on a fresh input, where oracle/_ref/ref_kmeans_driver exists;
  * a pure-numpy restatement of the similarity sweeps on a small case."""
import json
import os
import sys

import numpy as np
import pytest

import oracle_lib as O
from conftest import GOLD, ROOT, golden_msa
from test_oracle_relvars import partition_by_site, relvars_cases, window_codes

DRV = os.path.join(ROOT, "oracle", "_ref", "ref_kmeans_driver")


def kmeans_cases():
    path = os.path.join(GOLD, "kmeans.json")
    if not os.path.exists(path):                       # only while oracle/gen_golden_kmeans.py writes it
        return {}
    with open(path) as f:
        return json.load(f)


def test_fixture_is_committed():
    assert sorted(kmeans_cases()) == ["distributed_small", "saturated", "tree_small"]


@pytest.mark.parametrize("name", sorted(kmeans_cases()))
def test_kmeans_matches_the_unmodified_reference(name):
    rel = relvars_cases()[name]
    codes = window_codes(golden_msa(name), rel["von"], rel["bis"])
    o = O.Oracle.from_codes(codes)
    M, _, _ = o.scan(rel["mincov"])
    ut, _ = partition_by_site(codes, M)
    split_seen = set()
    for key, want in kmeans_cases()[name].items():
        u_no, mingroup = (int(x) for x in key.split("/"))
        n, after = o.kmeans(ut, u_no, rel["parts"][str(u_no)]["vars"], mingroup)
        assert n == want["split"] and list(after) == want["after"], key
        # only the reads of the part move, all of them to numbers above the old maximum (2814-2815)
        moved = after != ut
        assert (moved == (ut == u_no)).all() and (after[moved] > ut.max()).all()
        assert len(set(after[moved])) == n
        split_seen.add(n)
    assert max(split_seen) > 1


@pytest.mark.skipif(not os.path.exists(DRV), reason="oracle/_ref/ref_kmeans_driver not built (no reference sources here)")
def test_kmeans_matches_the_reference_binary_on_fresh_input():
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import repeatresolver_b200 as rr
    from gen_golden_kmeans import run_driver
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
        vars_ = o.relative_vars(ut, u_no, M, 3.0, 8)
        if len(vars_) == 0:
            continue
        vars_ = vars_[:400]                                              # the driver takes them on its command line
        for mingroup in (4, 8, 20):
            n, after = o.kmeans(ut, u_no, vars_, mingroup)
            R, Nw, split, want = run_driver(text, von, bis, ut, u_no, mingroup, vars_)
            assert (R, Nw) == codes.shape and n == split and list(after) == want, (u_no, mingroup)
            checked += 1
    assert checked >= 3


def test_kmeans_first_assignment_against_numpy():
    """mingroup <= 2 skips the dissolution loop: the result is the plain best-centroid assignment, restated with numpy"""
    rng = np.random.default_rng(4)
    R, N = 40, 30
    codes = rng.integers(0, 2, size=(R, N)).astype(np.uint8)
    codes[:20, :10] = 0
    codes[20:, :10] = 1                                                  # two families of reads
    o = O.Oracle.from_codes(codes)
    ut = np.zeros(R, dtype=np.int32)
    ut[::7] = 3                                                          # a few reads belong to another part
    vars_ = np.array([5 * c + 1 for c in range(N)], dtype=np.int32)     # the "1" group of every column
    n, after = o.kmeans(ut, 0, vars_, 2)
    I = np.flatnonzero(ut == 0)
    S = (codes[I] == 1).astype(np.int64)                                 # signatures over vars_
    scv64 = (len(vars_) // 64 + 1) * 64

    def match(a, b):
        return scv64 - int((a != b).sum())

    cen = np.zeros_like(S)
    for i in range(len(I)):
        bs, bj = [0] * 5, [0] * 5
        for j in range(len(I)):
            sc_ = match(S[j], S[i])
            for k in range(5):
                for l in range(k + 1, 5):
                    if bs[l] < bs[k]:
                        bs[l], bs[k] = bs[k], bs[l]
                        bj[l], bj[k] = bj[k], bj[l]
            if sc_ > bs[0]:
                bs[0], bj[0] = sc_, j
        cen[i] = S[bj].sum(0) > 2
    cluster = []
    for i in range(len(I)):
        best, bjj = 0, 0
        for j in range(len(I)):
            sc_ = match(cen[j], S[i])
            if sc_ > best and i != j:
                best, bjj = sc_, j
        cluster.append(bjj)
    want = ut.copy()
    want[I] = np.array(cluster) + ut.max() + 1
    assert list(after) == list(want) and n == len(set(cluster))
    # reads of the two families do not share a cluster
    fam = (np.arange(R) >= 20)[I]
    for c in set(cluster):
        assert len(set(fam[np.array(cluster) == c])) == 1


# ---- the product's host pieces and the integer rules its kernels share with the host (csrc/rr_kmeans.h) ----------------
def popcount64(a):
    return np.unpackbits(a.view(np.uint8), axis=-1).sum(-1).astype(np.int64)


def test_majority_of_five_is_exact_for_all_patterns():
    import repeatresolver_b200 as rr
    rng = np.random.default_rng(6)
    for _ in range(300):
        w = [int(x) for x in rng.integers(0, 2 ** 63, 5, dtype=np.uint64) * 2 + rng.integers(0, 2, 5, dtype=np.uint64)]
        got = rr.kmeans_majority5_host(*w)
        want = 0
        for b in range(64):
            if sum((x >> b) & 1 for x in w) > 2:
                want |= 1 << b
        assert got == want
    assert rr.kmeans_majority5_host(2 ** 64 - 1, 2 ** 64 - 1, 2 ** 64 - 1, 0, 0) == 2 ** 64 - 1
    assert rr.kmeans_majority5_host(2 ** 64 - 1, 2 ** 64 - 1, 0, 0, 0) == 0


@pytest.mark.parametrize("name", sorted(kmeans_cases()))
def test_product_pieces_chain_to_the_reference_result(name):
    """signatures (host), five-slot rule and majority (shared with the kernels), first-best assignment (numpy here, a kernel
    in the product), dissolution (host): chained, they must give the unmodified reference's partition"""
    import repeatresolver_b200 as rr
    rel = relvars_cases()[name]
    codes = window_codes(golden_msa(name), rel["von"], rel["bis"])
    o = O.Oracle.from_codes(codes)
    M, _, _ = o.scan(rel["mincov"])
    ut, _ = partition_by_site(codes, M)
    msa = rr.MSA.from_cells(codes, codes=True)
    for key, want in kmeans_cases()[name].items():
        u_no, mingroup = (int(x) for x in key.split("/"))
        vars_ = np.array(rel["parts"][str(u_no)]["vars"], dtype=np.int32)
        reads, sig = rr.kmeans_signatures(msa, ut, u_no, vars_)
        assert list(reads) == list(np.flatnonzero(ut == u_no))
        bits = np.zeros((len(reads), sig.shape[1] * 64), dtype=np.uint8)
        bits[:, :len(vars_)] = codes[reads][:, vars_ // 5] == vars_ % 5
        assert np.array_equal(sig, np.packbits(bits, axis=1, bitorder="little").view(np.uint64))
        n, scv = sig.shape
        cen = np.zeros_like(sig)
        for i in range(n):
            bj = rr.kmeans_top5_host(sig, i)
            for z in range(scv):
                cen[i, z] = rr.kmeans_majority5_host(*(int(sig[j, z]) for j in bj))
        cluster = np.zeros(n, dtype=np.int32)
        for i in range(n):
            score = scv * 64 - popcount64(cen ^ sig[i]).reshape(n, -1).sum(1)
            score[i] = -1                                            # i != j (2717)
            best = int(score.max())
            cluster[i] = int(np.argmax(score)) if best > 0 else 0    # first best; nothing above 0 leaves slot 0
        final, split = rr.kmeans_finish(sig, cen, cluster, mingroup)
        after = ut.copy()
        after[reads] = final + ut.max() + 1
        assert split == want["split"] and list(after) == want["after"], key
    msa.close()


def test_kmeans_needs_a_device():
    import repeatresolver_b200 as rr
    if rr.device_count() > 0:
        pytest.skip("a CUDA device is present")
    codes = np.zeros((4, 6), dtype=np.uint8)
    msa = rr.MSA.from_cells(codes, codes=True)
    with pytest.raises(rr.RRError):
        rr.Kmeans(msa, np.zeros(4, dtype=np.int32), 0, [0, 5], 3)
    with pytest.raises(rr.RRError):
        rr.Relative_Vars(msa, np.zeros(4, dtype=np.int32), 0, np.full(30, 9.0), 3.0, 2)
    msa.close()
