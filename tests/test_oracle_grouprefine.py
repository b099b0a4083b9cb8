"""SURVEY.md section 8f row 2 as a whole: Group_Refinement (/root/reference/RepeatResolver.c:1634-1693) = Cliquer for every
group above the cutoff, Sizes (1650), Dropoff_Cutoff (1460-1522; the results of BestCutoff and KorrMaxCutoff are overwritten
at 1662), CliqueGroup, CliqueCoverage, the two GroupPrecision printouts, MaxCorrs zeroed where Sizes <= 5.
  * the restatement (tests/oracle_lib.py: group_refinement on the C oracle's Cliquer, dropoff_cutoff, group_precision) against
    the committed output of the UNMODIFIED RepeatResolver.c (tests/golden/grouprefine.json, oracle/gen_golden_grouprefine.py);
  * against the reference binary itself on a fresh input - serial and Parallel_Group_Refinement, whose int argument array
    truncates cutoff and greedy (1793-1794) - where oracle/_ref/ref_grouprefine_driver exists."""
import json
import os
import sys

import numpy as np
import pytest

import oracle_lib as O
from conftest import GOLD, ROOT, golden_msa
from test_oracle_cliquer import window_codes

DRV = os.path.join(ROOT, "oracle", "_ref", "ref_grouprefine_driver")


def grouprefine_cases():
    with open(os.path.join(GOLD, "grouprefine.json")) as f:
        return json.load(f)


def words(hexes):
    return np.array([int(h, 16) for h in hexes], dtype=np.uint64)


def case_inputs(name, case):
    """codes of the window and the MaxCorrs the reference read: the oracle's scan through the "%f" text"""
    codes = window_codes(golden_msa(name), case["von"], case["bis"])
    assert codes.shape == (case["rows"], case["cols"])
    o = O.Oracle.from_codes(codes)
    M, _, _ = o.scan(case["mincov"])
    M = np.array([float(l) for l in O.fmt_lines(M).split()], dtype=np.float64)
    return codes, o, M


def check_against_records(codes, got_M, got, records, maxclique):
    """records: {group: the reference's Sizes / Cutoffs / Drop_Off / MaxCorrs / Cliques / C_Groups / C_Coverage}"""
    assert sorted(got) == sorted(int(i) for i in records)
    refined = 0
    for key, want in records.items():
        i = int(key)
        r = got[i]
        assert list(r["clique"]) == want["clique"] and len(want["clique"]) == maxclique + 1, i
        assert r["size"] == want["size"], i
        assert r["cutoff"] == want["cutoff"], i
        assert float(r["drop_off"]).hex() == want["drop_off"], i
        assert float(got_M[i]).hex() == want["maxcorr"], i
        if want["size"] > 5:
            refined += 1
            assert np.array_equal(O.bitset_words(r["group"]), words(want["group"])), i
            assert np.array_equal(O.bitset_words(r["coverage"]), words(want["coverage"])), i
            if "precision" in want:
                assert list(O.group_precision(codes[:, i // 5] == i % 5)) == want["precision"][0], i
                assert list(O.group_precision(r["group"])) == want["precision"][1], i
        else:
            assert r["group"] is None and want["group"] == [] and want["coverage"] == [] and got_M[i] == 0.0
    return refined


@pytest.mark.parametrize("name", sorted(grouprefine_cases()))
def test_group_refinement_matches_the_unmodified_reference(name):
    case = grouprefine_cases()[name]
    codes, o, M = case_inputs(name, case)
    got_M, got = O.group_refinement(o, codes, M, case["cutoff"], case["mincov"], case["maxclique"], case["greedy"])
    refined = check_against_records(codes, got_M, got, case["groups"], case["maxclique"])
    assert refined >= 20
    untouched = np.ones(len(M), dtype=bool)
    untouched[[int(i) for i in case["groups"]]] = False
    assert np.array_equal(got_M[untouched], M[untouched])


def deep_case():
    """the deep case of oracle/gen_golden_grouprefine.py (a window > 3 000 generated reads span; MaxCorrs = 50 on chosen groups):
    (golden record, codes of the window, MaxCorrs, parameters)"""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    from gen_golden_grouprefine import DEEP, deep_inputs
    with open(os.path.join(GOLD, "grouprefine_deep.json")) as f:
        case = json.load(f)
    text, von, bis, codes, M = deep_inputs()
    assert (von, bis) == (case["von"], case["bis"]) and codes.shape == (case["rows"], case["cols"])
    return case, codes, M, DEEP


def test_group_refinement_deep_matches_the_unmodified_reference():
    """51 words of 64 reads per group: the restatement against what the unmodified RepeatResolver.c left for the deep case"""
    case, codes, M, P = deep_case()
    o = O.Oracle.from_codes(codes)
    got_M, got = O.group_refinement(o, codes, M, P["cutoff"], P["mincov"], P["maxclique"], P["greedy"])
    assert check_against_records(codes, got_M, got, case["groups"], P["maxclique"]) >= 30


def test_dropoff_cutoff_rules():
    """the corners of 1488-1509 on hand-made member counts: first minimum wins, zero denominators are skipped, nothing
    admissible leaves cutoff max(1, c) and Drop_Off 1e6"""
    # 8 reads, clique of 6 single-site groups: read r is in the first cnt[r] members
    def codes_for(cnt, size):
        c = np.full((len(cnt), size + 1), 1, dtype=np.uint8)              # site 0 unused: group 0 would end Sizes (1650)
        for r, n in enumerate(cnt):
            c[r, 1:1 + n] = 0
        return c, [5 * (k + 1) for k in range(size)]
    codes, clique = codes_for([6, 6, 6, 3, 3, 1, 0, 0], 6)
    # sizes = [6, 5, 5, 3, 3, 3]; drops: k=1: (6-5)/min(3,5)=1/3, k=2: (5-3)/3, k=3: (5-3)/3, k=4: (3-3)/3 = 0
    assert O.dropoff_cutoff(codes, clique, 6) == (4, 0.0)
    codes, clique = codes_for([6, 6, 6, 6], 6)                            # every read in every member: signumber - sizes = 0
    assert O.dropoff_cutoff(codes, clique, 6) == (1, 1000000.0)
    codes, clique = codes_for([2, 2, 5, 5, 0, 0], 6)                      # sizes = [4, 4, 2, 2, 2, 0]: ties at k = 3 and 4?
    # k=1: (4-2)/min(2,4)=1, k=2: (4-2)/2=1, k=3: (2-2)/2=0, k=4: (2-0)/2=1 -> 3
    assert O.dropoff_cutoff(codes, clique, 6) == (3, 0.0)
    assert O.dropoff_cutoff(codes, clique, 6, c=4) == (4, 1.0)


@pytest.mark.skipif(not os.path.exists(DRV), reason="oracle/_ref is built in the build container only")
def test_group_refinement_against_the_reference_binary_on_a_fresh_input():
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    from gen_golden_grouprefine import run_driver
    import repeatresolver_b200 as rr
    g = rr.MsaGen(type="Tree", copies=5, coverage=20, repeat_len=900, diff=0.03, seed=78, flank=400, min_overlap=80)
    text = g.text()
    width = len(text.split(b"\n")[0])
    von, bis = width // 10, width - 1 - width // 10
    codes = window_codes(text, von, bis)
    o = O.Oracle.from_codes(codes)
    M, _, _ = o.scan(12)
    mtext = O.fmt_lines(M)
    M = np.array([float(l) for l in mtext.split()], dtype=np.float64)
    cutoff, greedy, maxclique = 5.75, 2.5, 20
    R, N, sc, res, prec = run_driver(text, von, bis, mtext, cutoff, 12, maxclique, greedy)
    assert (R, N) == codes.shape
    got_M, got = O.group_refinement(o, codes, M, cutoff, 12, maxclique, greedy)
    refined = check_against_records(codes, got_M, got, {str(k): v for k, v in res.items()}, maxclique)
    assert refined >= 10 and len(prec) == 2 * refined
    # Parallel_Group_Refinement: the same with (int)cutoff and (int)greedy
    R, N, sc, res, prec = run_driver(text, von, bis, mtext, cutoff, 12, maxclique, greedy, threads=4)
    got_M, got = O.group_refinement(o, codes, M, float(int(cutoff)), 12, maxclique, float(int(greedy)))
    assert check_against_records(codes, got_M, got, {str(k): v for k, v in res.items()}, maxclique) >= refined


def group_zero_member_case():
    """an MSA window whose first column is a copy of a strongly correlated 'A' column, so that group 0 enters the cliques of
    that column's partners: (text, von, bis, codes of the window)"""
    import repeatresolver_b200 as rr
    g = rr.MsaGen(type="Tree", copies=4, coverage=25, repeat_len=700, diff=0.04, seed=31, flank=300, min_overlap=80)
    text = g.text()
    lines = [bytearray(l) for l in text.split(b"\n") if l]
    width = len(lines[0])
    von, bis = width // 8, width - 1 - width // 8
    codes = window_codes(text, von, bis)
    M, _, _ = O.Oracle.from_codes(codes).scan(12)
    src = next(int(q) for q in np.argsort(-M, kind="stable") if q % 5 == 0 and q >= 5)
    for l in lines:
        if l[von:von + 1] != b" " and l[von + src // 5:von + src // 5 + 1] != b" ":
            l[von] = l[von + src // 5]
    text = b"\n".join(bytes(l) for l in lines) + b"\n"
    return text, von, bis, window_codes(text, von, bis)


@pytest.mark.skipif(not os.path.exists(DRV), reason="oracle/_ref is built in the build container only")
def test_group_zero_as_a_member_ends_sizes_but_not_the_clique():
    """1650 sizes a clique by its first entry <= 0, CliqueGroup by its first negative entry (986-993): with group 0 among the
    members, Dropoff_Cutoff sees the members before it and CliqueGroup all of them."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    from gen_golden_grouprefine import run_driver
    text, von, bis, codes = group_zero_member_case()
    o = O.Oracle.from_codes(codes)
    M, _, _ = o.scan(12)
    mtext = O.fmt_lines(M)
    M = np.array([float(l) for l in mtext.split()], dtype=np.float64)
    cutoff, greedy, maxclique = 6.0, 3.0, 16
    R, N, sc, res, prec = run_driver(text, von, bis, mtext, cutoff, 12, maxclique, greedy)
    assert (R, N) == codes.shape
    inner = [i for i, r in res.items() if 0 in r["clique"][1:]]
    assert any(res[i]["size"] > 5 for i in inner) and any(res[i]["size"] <= 5 for i in inner), [(i, res[i]["size"]) for i in inner]
    got_M, got = O.group_refinement(o, codes, M, cutoff, 12, maxclique, greedy)
    check_against_records(codes, got_M, got, {str(k): v for k, v in res.items()}, maxclique)
