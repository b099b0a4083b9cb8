"""GPU test (-m gpu): rr_group_refinement on the deep case (3 229 reads x 2 641 columns, 51 words of 64 reads per group, 83 groups
above the cutoff, 37 refined) against what the UNMODIFIED /root/reference/RepeatResolver.c left in its globals
(tests/golden/grouprefine_deep.json, oracle/gen_golden_grouprefine.py); the CPU suite holds the restatement against the same
file (tests/test_oracle_grouprefine.py).  Written after the round's GPU minutes were spent: the device path it exercises is
the one tests/test_zz_gpu_cliquer.py validates against the restatement at the same depth.  Sorts last on purpose."""
import numpy as np
import pytest

import oracle_lib as O
import repeatresolver_b200 as rr
from test_oracle_grouprefine import deep_case

def _words(hexes):
    return np.array([int(h, 16) for h in hexes], dtype=np.uint64)


def check(got, case, codes):
    """got: the layout of Packed.group_refinement; case: the reference's record"""
    R = codes.shape[0]
    assert [int(g) for g in got["groups"]] == sorted(int(i) for i in case["groups"])
    refined = 0
    for q, g in enumerate(got["groups"]):
        want = case["groups"][str(int(g))]
        assert list(got["Cliques"][q]) == want["clique"], g
        assert int(got["Sizes"][q]) == want["size"] and int(got["Cutoffs"][q]) == want["cutoff"], g
        assert float(got["Drop_Off"][q]).hex() == want["drop_off"] and float(got["MaxCorrs"][g]).hex() == want["maxcorr"], g
        if want["size"] > 5:
            refined += 1
            assert np.array_equal(got["C_Groups"][q], _words(want["group"])) and np.array_equal(got["C_Coverage"][q], _words(want["coverage"])), g
            assert list(rr.GroupPrecision(O.bitset_words(codes[:, g // 5] == g % 5), R)) == want["precision"][0]
            assert list(rr.GroupPrecision(got["C_Groups"][q], R)) == want["precision"][1]
        else:
            assert not got["C_Groups"][q].any() and not got["C_Coverage"][q].any() and got["MaxCorrs"][g] == 0.0
    return refined


@pytest.mark.gpu
def test_group_refinement_deep_golden():
    case, codes, M, P = deep_case()
    pk = rr.Packed(rr.MSA.from_cells(codes, codes=True), 0)
    got = rr.Group_Refinement(pk, M, P["cutoff"], 0, codes.shape[1], P["mincov"], P["maxclique"], P["greedy"])
    assert check(got, case, codes) >= 30
    pk.close()


def test_the_checker_of_the_deep_case_on_the_restatement():
    """no GPU: the restatement's result in the product's layout through the same checker (so the GPU test above cannot fail on
    its own bookkeeping)"""
    case, codes, M, P = deep_case()
    o = O.Oracle.from_codes(codes)
    want_M, want = O.group_refinement(o, codes, M, P["cutoff"], P["mincov"], P["maxclique"], P["greedy"])
    groups = np.array(sorted(want), dtype=np.int32)
    sc = codes.shape[0] // 64 + 1
    got = {"MaxCorrs": want_M, "groups": groups, "Cliques": np.array([want[g]["clique"] for g in groups], dtype=np.int32),
           "Sizes": np.array([want[g]["size"] for g in groups], dtype=np.int32), "Cutoffs": np.array([want[g]["cutoff"] for g in groups], dtype=np.int32),
           "Drop_Off": np.array([want[g]["drop_off"] for g in groups]), "C_Groups": np.zeros((len(groups), sc), dtype=np.uint64),
           "C_Coverage": np.zeros((len(groups), sc), dtype=np.uint64)}
    for k, g in enumerate(groups):
        if want[g]["size"] > 5:
            got["C_Groups"][k] = O.bitset_words(want[g]["group"])
            got["C_Coverage"][k] = O.bitset_words(want[g]["coverage"])
    assert check(got, case, codes) >= 30
