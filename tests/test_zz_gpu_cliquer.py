"""GPU parity tests (-m gpu) for SURVEY.md section 8f row 2: Cliquer (/root/reference/RepeatResolver.c:1179-1240),
called through the C ABI (rr_cliquer_batch = device path, rr_cliquer = plain path), against
  * the committed output of the UNMODIFIED RepeatResolver.c (tests/golden/cliquer.json), bit for bit;
  * the oracle's restatement on a flanked MSA deep enough for several 32-word chunks per bitset (the kernel's two
    exact skips), with a query count that is not a multiple of the batch, sub-ranges [anfang, ende), small cliques,
    and the list-overflow retry forced through the debug hook rr_debug_set_cliquer_cap.
Bar: members identical, scores identical as doubles (integer counts on the device, final scores with the host libm).
The file sorts last on purpose: a first failure here must not hide the scan's parity tests under `-x`."""
import numpy as np
import pytest

import repeatresolver_b200 as rr
from conftest import golden_msa
import oracle_lib as O
from test_oracle_cliquer import cliquer_cases, window_codes

pytestmark = pytest.mark.gpu


def check(members, scores, n, k, oracle_members, oracle_scores):
    assert n[k] == len(oracle_members), (k, n[k], len(oracle_members))
    assert list(members[k, :n[k]]) == list(oracle_members), k
    assert (members[k, n[k]:] == -1).all()
    assert [float(z).hex() for z in scores[k, :n[k]]] == [float(z).hex() for z in oracle_scores], k
    assert (scores[k, n[k]:] == 0).all()


@pytest.mark.parametrize("name", sorted(cliquer_cases()))
def test_cliquer_golden(name):
    case = cliquer_cases()[name]
    codes = window_codes(golden_msa(name), case["von"], case["bis"])
    pk = rr.Packed(rr.MSA.from_cells(codes, codes=True), 0)
    queries = [int(q) for q in case["queries"]]
    members, scores, n, st = pk.cliquer_batch(queries, case["mincov"], case["maxclique"], case["greedy"])
    assert st["launches"] >= 2 and st["retries"] == 0 and st["hits"] <= st["candidates"] <= st["pairs"]
    for k, q in enumerate(queries):
        want = case["queries"][str(q)]
        check(members, scores, n, k, [q] + want["members"], [100.0] + [float.fromhex(s) for s in want["scores"]])
        m1, z1 = pk.cliquer(q, case["mincov"], case["maxclique"], case["greedy"])      # the plain path
        assert list(m1) == [q] + want["members"] and [float(z).hex() for z in z1[1:]] == want["scores"]
        assert list(rr.Cliquer(pk, 0, codes.shape[1], case["mincov"], case["maxclique"], case["greedy"], q)[:n[k]]) == list(m1)
    pk.close()


@pytest.fixture(scope="module")
def deep():
    g = rr.MsaGen(type="Tree", copies=80, coverage=40, repeat_len=1500, diff=0.03, seed=5, flank=500, min_overlap=100)
    codes = g.codes()
    assert codes.shape[0] > 3072                                   # > 3 chunks of 32 words per bitset
    o = O.Oracle.from_codes(codes)
    gs = o.gsize()
    cand = np.flatnonzero((gs > 30) & (gs < codes.shape[0] // 3))
    queries = [int(q) for q in cand[::len(cand) // 61][:61]] + [0, 5 * codes.shape[1] - 1]      # 63 queries: a partial last block with 4 and with 6 per block
    pk = rr.Packed(rr.MSA.from_cells(codes, codes=True), 0)
    yield codes, o, pk, queries
    pk.close()


def test_cliquer_batch_deep_msa_against_oracle(deep):
    codes, o, pk, queries = deep
    members, scores, n, st = pk.cliquer_batch(queries, 30, 30, 3.0)
    sizes = set()
    for k, q in enumerate(queries):
        m0, z0 = o.cliquer(q, 30, 30, 3.0)
        check(members, scores, n, k, m0, z0)
        sizes.add(len(m0))
    assert 1 in sizes and 30 in sizes and len(sizes) > 5
    inside = len(queries)                                           # every query group is itself a candidate column
    assert st["pairs"] == len(queries) * 5 * codes.shape[1] - inside
    assert 0 < st["hits"] <= st["candidates"] < st["pairs"] and st["host_evals"] <= st["hits"]
    print(f"cliquer_batch deep: R={codes.shape[0]} N={codes.shape[1]} queries={len(queries)} {st}")
    # the plain path on a few of them
    for q in queries[2:40:9]:
        m0, z0 = o.cliquer(q, 30, 30, 3.0)
        m1, z1 = pk.cliquer(q, 30, 30, 3.0)
        assert list(m1) == list(m0) and np.array_equal(z1, z0)


@pytest.mark.parametrize("anfang,ende,maxclique,greedy,mincov", [(1000, 4000, 6, 3.0, 30), (37, 38, 30, 0.5, 8),
                                                                 (6000, 10 ** 6, 2, 10.0, 30), (300, 300, 5, 3.0, 30),
                                                                 (0, 2500, 1, 3.0, 30)])
def test_cliquer_batch_subranges_and_small_cliques(deep, anfang, ende, maxclique, greedy, mincov):
    codes, o, pk, queries = deep
    qs = queries[:21]
    members, scores, n, st = pk.cliquer_batch(qs, mincov, maxclique, greedy, anfang, ende)
    for k, q in enumerate(qs):
        m0, z0 = o.cliquer(q, mincov, maxclique, greedy, anfang, min(ende, codes.shape[1]))
        check(members, scores, n, k, m0, z0)


def test_cliquer_batch_list_overflow_retries(deep):
    codes, o, pk, queries = deep
    want = pk.cliquer_batch(queries[:13], 30, 30, 3.0)
    rr.debug.set_cliquer_cap(700)
    try:
        got = pk.cliquer_batch(queries[:13], 30, 30, 3.0)
    finally:
        rr.debug.set_cliquer_cap(0)
    assert got[3]["retries"] > 0 and want[3]["retries"] == 0
    for a, b in zip(want[:3], got[:3]):
        assert np.array_equal(a, b)


def test_group_refinement_cliques_mirror(deep):
    codes, o, pk, queries = deep
    M = np.zeros(5 * codes.shape[1])
    M[queries[:9]] = np.linspace(5, 50, 9)
    groups, cliques, sizes, scores, st = rr.Group_Refinement_Cliques(pk, M, 10.0, 0, codes.shape[1], 30, 30, 3.0)
    assert list(groups) == sorted(q for q in queries[:9] if M[q] > 10.0)
    for k, q in enumerate(groups):
        m0, z0 = o.cliquer(int(q), 30, 30, 3.0)
        assert list(cliques[k, :len(m0)]) == list(m0) and cliques[k, len(m0)] == -1
        assert sizes[k] == (len(m0) if q > 0 else 0)               # 1650 counts entries > 0 from Clique[0]


def test_cliquer_argument_errors(deep):
    codes, o, pk, queries = deep
    with pytest.raises(rr.RRError):
        pk.cliquer_batch([5 * codes.shape[1]], 30, 30, 3.0)
    with pytest.raises(rr.RRError):
        pk.cliquer_batch([0], -1, 30, 3.0)
    with pytest.raises(rr.RRError):
        pk.cliquer_batch([0], 30, 0, 3.0)
    members, scores, n, st = pk.cliquer_batch([], 30, 30, 3.0)
    assert members.shape == (0, 31) and st["launches"] == 0


# ---- CliqueGroup / CliqueCoverage (RepeatResolver.c:976-1008, 1064-1096) through rr_clique_groups ------------------------
def _golden_words(hexes):
    return np.array([int(h, 16) for h in hexes], dtype=np.uint64)


@pytest.mark.parametrize("name", ["tree_small", "distributed_small", "saturated"])
def test_clique_groups_golden(name):
    """the device path against the committed output of the UNMODIFIED RepeatResolver.c (tests/golden/cliquegroup.json): the
    clique comes from the product's own Cliquer, group and coverage words must be identical"""
    from test_oracle_cliquegroup import cliquegroup_cases
    case = cliquegroup_cases()[name]
    codes = window_codes(golden_msa(name), case["von"], case["bis"])
    pk = rr.Packed(rr.MSA.from_cells(codes, codes=True), 0)
    for q, rec in case["queries"].items():
        clique = rr.Cliquer(pk, 0, codes.shape[1], case["mincov"], case["maxclique"], case["greedy"], int(q))
        n = int(np.argmax(clique < 0)) if (clique < 0).any() else len(clique)
        assert list(clique[:n]) == rec["clique"]
        cuts = sorted(int(c) for c in rec["cutoffs"])
        G, V = pk.clique_groups(np.tile(clique, (len(cuts), 1)), cuts)
        for k, c in enumerate(cuts):
            assert np.array_equal(G[k], _golden_words(rec["cutoffs"][str(c)]["group"])), (name, q, c)
            assert np.array_equal(V[k], _golden_words(rec["cutoffs"][str(c)]["coverage"])), (name, q, c)
            assert np.array_equal(rr.CliqueGroup(pk, clique, c), G[k]) and np.array_equal(rr.CliqueCoverage(pk, clique, c), V[k])
    pk.close()


def test_clique_groups_deep_against_the_oracle(deep):
    """several 32-read words per bitset, cliques of 0 .. 100 members (random, repeated and the product's own cliques), every
    kind of cutoff; one output at a time as well"""
    codes, o, pk = deep[0], deep[1], deep[2]
    R, N = codes.shape
    rng = np.random.default_rng(3)
    cliques, cuts = [], []
    for n in (0, 1, 2, 7, 30, 31, 64, 99, 100):
        for c in (-1, 0, 1, n // 2, n - 1, n, 127, 1000):
            m = [int(x) for x in rng.integers(0, 5 * N, n)]
            if n > 3:
                m[2] = m[0]
            cliques.append(m)
            cuts.append(c)
    G, V = pk.clique_groups(cliques, cuts)
    for k, (m, c) in enumerate(zip(cliques, cuts)):
        assert np.array_equal(G[k], O.bitset_words(O.clique_group(codes, m, c))), (k, len(m), c)
        assert np.array_equal(V[k], O.bitset_words(O.clique_coverage(codes, m, c))), (k, len(m), c)
        assert list(rr.group_reads(G[k], R)) == list(np.flatnonzero(O.clique_group(codes, m, c)))
    G2, none = pk.clique_groups(cliques, cuts, want_coverage=False)
    assert none is None and np.array_equal(G2, G)
    none, V2 = pk.clique_groups(cliques, cuts, want_groups=False)
    assert none is None and np.array_equal(V2, V)


def test_clique_groups_bad_arguments(deep):
    pk = deep[2]
    with pytest.raises(rr.RRError):
        pk.clique_groups([[5 * deep[0].shape[1]]], [0])                      # group out of range
    with pytest.raises(rr.RRError):
        pk.clique_groups([list(range(101))], [0])                            # more members than the reference's limit (986)
    G, V = pk.clique_groups(np.zeros((0, 4), dtype=np.int32), [])
    assert G.shape[0] == 0 and V.shape[0] == 0


# ---- Group_Refinement as a whole (RepeatResolver.c:1634-1693) through rr_group_refinement ------------------------------------
def _check_refinement(got, records, codes, maxclique):
    """records: {group: Sizes / Cutoffs / Drop_Off / MaxCorrs / Cliques / C_Groups / C_Coverage of the reference or the oracle}"""
    R = codes.shape[0]
    assert [int(g) for g in got["groups"]] == sorted(int(i) for i in records)
    refined = 0
    for q, g in enumerate(got["groups"]):
        want = records[str(int(g))]
        assert list(got["Cliques"][q]) == list(want["clique"]), g
        assert int(got["Sizes"][q]) == want["size"] and int(got["Cutoffs"][q]) == want["cutoff"], g
        assert float(got["Drop_Off"][q]).hex() == want["drop_off"], g
        assert float(got["MaxCorrs"][g]).hex() == want["maxcorr"], g
        if want["size"] > 5:
            refined += 1
            assert np.array_equal(got["C_Groups"][q], _golden_words(want["group"])), g
            assert np.array_equal(got["C_Coverage"][q], _golden_words(want["coverage"])), g
            if "precision" in want:
                assert list(rr.GroupPrecision(O.bitset_words(codes[:, g // 5] == g % 5), R)) == want["precision"][0]
                assert list(rr.GroupPrecision(got["C_Groups"][q], R)) == want["precision"][1]
        else:
            assert not got["C_Groups"][q].any() and not got["C_Coverage"][q].any() and got["MaxCorrs"][g] == 0.0
    return refined


@pytest.mark.parametrize("name", ["tree_small", "distributed_small", "saturated"])
def test_group_refinement_golden(name):
    """the device path against what the UNMODIFIED RepeatResolver.c left in its globals (tests/golden/grouprefine.json)"""
    from test_oracle_grouprefine import grouprefine_cases, case_inputs
    case = grouprefine_cases()[name]
    codes, o, M = case_inputs(name, case)
    pk = rr.Packed(rr.MSA.from_cells(codes, codes=True), 0)
    got = rr.Group_Refinement(pk, M, case["cutoff"], 0, codes.shape[1], case["mincov"], case["maxclique"], case["greedy"])
    assert _check_refinement(got, case["groups"], codes, case["maxclique"]) >= 20
    untouched = np.ones(len(M), dtype=bool)
    untouched[got["groups"]] = False
    assert np.array_equal(got["MaxCorrs"][untouched], M[untouched])
    # the parallel form truncates cutoff and greedy (1793-1794)
    par = rr.Parallel_Group_Refinement(pk, M, case["cutoff"], 0, codes.shape[1], case["mincov"], case["maxclique"], case["greedy"], 4)
    ser = rr.Group_Refinement(pk, M, float(int(case["cutoff"])), 0, codes.shape[1], case["mincov"], case["maxclique"], float(int(case["greedy"])))
    for key in ("MaxCorrs", "groups", "Cliques", "Sizes", "Cutoffs", "Drop_Off", "C_Groups", "C_Coverage"):
        assert np.array_equal(par[key], ser[key]), key
    # only the cutoffs
    few = pk.group_refinement(M, case["cutoff"], case["mincov"], case["maxclique"], case["greedy"], want_groups=False, want_coverage=False)
    assert few["C_Groups"] is None and np.array_equal(few["Cutoffs"], got["Cutoffs"]) and np.array_equal(few["Drop_Off"], got["Drop_Off"])
    pk.close()


def test_group_refinement_deep_against_the_oracle(deep):
    """several 32-read words per bitset, a sub-range of columns, cliques of up to 100 members, queries without partners:
    everything identical to the restatement on the oracle's Cliquer"""
    codes, o, pk = deep[0], deep[1], deep[2]
    R, N = codes.shape
    gs = o.gsize()
    rng = np.random.default_rng(8)
    M = np.zeros(5 * N)
    cand = np.flatnonzero((gs > 30) & (gs < R // 3))
    M[cand[::max(1, len(cand) // 90)]] = 50.0
    M[[0, 5 * N - 1]] = 50.0
    M[rng.integers(0, 5 * N, 20)] = 50.0                                     # some without partners: Sizes <= 5
    for anfang, ende, maxclique, greedy in ((0, N, 30, 3.0), (N // 4, 3 * N // 4, 12, 6.5), (0, N, 100, 1.0)):
        got = pk.group_refinement(M, 7.5, 30, maxclique, greedy, anfang, ende)
        want_M, want = O.group_refinement(o, codes, M, 7.5, 30, maxclique, greedy, anfang, ende)
        assert np.array_equal(got["MaxCorrs"], want_M)
        assert [int(g) for g in got["groups"]] == sorted(want)
        refined = 0
        for q, g in enumerate(got["groups"]):
            w = want[int(g)]
            assert list(got["Cliques"][q]) == list(w["clique"]) and int(got["Sizes"][q]) == w["size"], g
            assert int(got["Cutoffs"][q]) == w["cutoff"] and float(got["Drop_Off"][q]).hex() == float(w["drop_off"]).hex(), g
            if w["size"] > 5:
                refined += 1
                assert np.array_equal(got["C_Groups"][q], O.bitset_words(w["group"])), g
                assert np.array_equal(got["C_Coverage"][q], O.bitset_words(w["coverage"])), g
            else:
                assert not got["C_Groups"][q].any() and not got["C_Coverage"][q].any()
        assert refined >= 20, (anfang, ende, refined)


def test_group_refinement_arguments(deep):
    codes, o, pk = deep[0], deep[1], deep[2]
    N = codes.shape[1]
    M = np.zeros(5 * N)
    got = pk.group_refinement(M, 3.0)                                        # nothing above the cutoff
    assert len(got["groups"]) == 0 and got["C_Groups"].shape == (0, codes.shape[0] // 64 + 1) and np.array_equal(got["MaxCorrs"], M)
    with pytest.raises(rr.RRError):
        pk.group_refinement(M, 3.0, maxclique=101)                           # Dropoff_Cutoff's array of 100 (1463)
    with pytest.raises(rr.RRError):
        pk.group_refinement(M, 3.0, greedy=-1.0)
    with pytest.raises(ValueError):
        pk.group_refinement(M[:-1], 3.0)
    # too little room: the count comes back, nothing is written
    import ctypes as C
    M[[10, 20, 30]] = 9.0
    n = C.c_int64(0)
    buf = np.zeros(64, dtype=np.int64)
    from repeatresolver_b200._lib import lib
    rc = lib.rr_group_refinement(pk._h, M.ctypes.data, 3.0, 0, N, 30, 30, 3.0, 2, buf.ctypes.data, buf.ctypes.data, buf.ctypes.data,
                                    buf.ctypes.data, buf.ctypes.data, None, None, C.byref(n), None)
    assert rc != 0 and n.value == 3 and M[10] == 9.0


def test_group_refinement_with_group_zero_as_a_member():
    """1650 sizes a clique by its first entry <= 0, CliqueGroup by its first negative entry (986-993): the input of
    tests/test_oracle_grouprefine.py, where the unmodified reference confirms the restatement on exactly this case"""
    from test_oracle_grouprefine import group_zero_member_case
    text, von, bis, codes = group_zero_member_case()
    o = O.Oracle.from_codes(codes)
    M, _, _ = o.scan(12)
    pk = rr.Packed(rr.MSA.from_cells(codes, codes=True), 0)
    got = pk.group_refinement(M, 6.0, 12, 16, 3.0)
    want_M, want = O.group_refinement(o, codes, M, 6.0, 12, 16, 3.0)
    assert np.array_equal(got["MaxCorrs"], want_M) and [int(g) for g in got["groups"]] == sorted(want)
    seen = set()
    for q, g in enumerate(got["groups"]):
        w = want[int(g)]
        assert list(got["Cliques"][q]) == list(w["clique"]) and int(got["Sizes"][q]) == w["size"] and int(got["Cutoffs"][q]) == w["cutoff"]
        assert float(got["Drop_Off"][q]).hex() == float(w["drop_off"]).hex()
        if w["size"] > 5:
            assert np.array_equal(got["C_Groups"][q], O.bitset_words(w["group"])) and np.array_equal(got["C_Coverage"][q], O.bitset_words(w["coverage"]))
        if 0 in list(w["clique"][1:]):
            seen.add(w["size"] > 5)
    assert seen == {True, False}
    pk.close()
