"""GPU test (-m gpu) of Kmeans_Subdivision (/root/reference/RepeatResolver.c:3382-3404), the caller of Relative_Vars and Kmeans:
the product's host composition over its two device calls (rr_relative_vars, rr_kmeans) against the committed output of the
UNMODIFIED RepeatResolver.c (tests/golden/subdivision.json) and against the same composition over the oracle's calls on a
fresh input.  The composition logic itself is pinned on the CPU (tests/test_oracle_subdivision.py), the two device calls one
by one (tests/test_zz_gpu_relvars.py, tests/test_zz_gpu_kmeans.py).  Sorts last on purpose."""
import numpy as np
import pytest

import oracle_lib as O
import repeatresolver_b200 as rr
from test_oracle_subdivision import case_inputs, oracle_calls, subdivision_cases

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", sorted(subdivision_cases()))
def test_kmeans_subdivision_golden(name):
    case = subdivision_cases()[name]
    codes, o, M = case_inputs(name, case)
    msa = rr.MSA.from_cells(codes, codes=True)
    for mingroup, want in case["after"].items():
        before = np.array(case["before"], dtype=np.int32)
        n, u = rr.Kmeans_Subdivision(msa, before, M, case["cutoff"], int(mingroup))
        assert list(u) == want, (name, mingroup)
        assert n == len(set(x for x in want if x >= 0))
        assert list(before) == case["before"]                                # the caller's partition is not touched
    msa.close()


def test_kmeans_subdivision_fresh_input_against_the_oracle_composition():
    g = rr.MsaGen(type="Tree", copies=6, coverage=30, repeat_len=1200, diff=0.03, seed=80, flank=400, min_overlap=80)
    codes = g.codes()
    o = O.Oracle.from_codes(codes)
    M, _, _ = o.scan(15)
    site = int(np.argmax(M)) // 5
    before = codes[:, site].astype(np.int32) + 2
    before[3::19] = -1
    rv, km = oracle_calls(o)
    msa = rr.MSA.from_cells(codes, codes=True)
    for mingroup in (5, 12):
        want_n, want = rr.Kmeans_Subdivision(None, before, M, 3.0, mingroup, relative_vars=rv, kmeans=km)
        n, u = rr.Kmeans_Subdivision(msa, before, M, 3.0, mingroup)
        assert n == want_n and np.array_equal(u, want), mingroup
    msa.close()
