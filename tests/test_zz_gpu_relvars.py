"""GPU parity tests (-m gpu) for SURVEY.md section 8f row 3, first version: Relative_Vars
(/root/reference/RepeatResolver.c:2424-2493) through the C ABI (rr_relative_vars: the part's rows packed on the device,
triple intersections from rr_pair_counts, two-sided score on the host) against the committed output of the UNMODIFIED
RepeatResolver.c (tests/golden/relvars.json) and the oracle on a fresh input.  Bar: identical group lists.
NOT YET RUN ON A GPU when it was committed (the round's GPU budget was spent): the device steps it uses (rr_pack,
rr_pair_counts) are the ones the scan's tests cover, the host half is pinned on the CPU in tests/test_oracle_relvars.py.
Sorts last so that a failure here cannot hide other tests under `-x`."""
import os

import numpy as np
import pytest

import repeatresolver_b200 as rr
from conftest import golden_msa
import oracle_lib as O
from test_oracle_relvars import partition_by_site, relvars_cases, window_codes

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", sorted(relvars_cases()))
def test_relative_vars_golden(name):
    case = relvars_cases()[name]
    codes = window_codes(golden_msa(name), case["von"], case["bis"])
    o = O.Oracle.from_codes(codes)
    M, _, _ = o.scan(case["mincov"])
    ut, _ = partition_by_site(codes, M)
    msa = rr.MSA.from_cells(codes, codes=True)
    for u_no, want in case["parts"].items():
        got = rr.Relative_Vars(msa, ut, int(u_no), M, case["cutoff"], case["mingroup"])
        assert list(got) == want["vars"], u_no
    assert len(rr.Relative_Vars(msa, ut, 77, M, case["cutoff"], case["mingroup"])) == 0      # an empty part
    msa.close()


def test_relative_vars_fresh_input_against_oracle():
    g = rr.MsaGen(type="Tree", copies=8, coverage=40, repeat_len=1500, diff=0.01, seed=23, flank=300)
    codes = g.codes()
    o = O.Oracle.from_codes(codes)
    M, _, _ = o.scan(30)
    ut, _ = partition_by_site(codes, M)
    msa = rr.MSA.from_cells(codes, codes=True)
    nonempty = 0
    for u_no in sorted(set(int(x) for x in ut)):
        want = o.relative_vars(ut, u_no, M, 3.0, 8)
        got = rr.Relative_Vars(msa, ut, u_no, M, 3.0, 8)
        assert list(got) == list(want), u_no
        nonempty += len(want) > 0
    assert nonempty >= 1
    msa.close()


@pytest.mark.skipif(os.environ.get("RR_TEST_UNVALIDATED") != "1",
                    reason="csrc/rr_relvars.cu has never run on a GPU: opt in with RR_TEST_UNVALIDATED=1 (first thing next round)")
def test_relative_vars_experimental_pair_kernel():
    """the tiled all-pairs kernel (RR_RELVARS_KERNEL=1) must give what the default path and the oracle give"""
    g = rr.MsaGen(type="Tree", copies=8, coverage=40, repeat_len=1500, diff=0.01, seed=23, flank=300)
    codes = g.codes()
    o = O.Oracle.from_codes(codes)
    M, _, _ = o.scan(30)
    ut, _ = partition_by_site(codes, M)
    msa = rr.MSA.from_cells(codes, codes=True)
    os.environ["RR_RELVARS_KERNEL"] = "1"
    try:
        for u_no in sorted(set(int(x) for x in ut)):
            for cutoff, mingroup in ((3.0, 8), (6.0, 20), (1.0, 3)):
                want = o.relative_vars(ut, u_no, M, cutoff, mingroup)
                got = rr.Relative_Vars(msa, ut, u_no, M, cutoff, mingroup)
                assert list(got) == list(want), (u_no, cutoff, mingroup)
    finally:
        del os.environ["RR_RELVARS_KERNEL"]
    # the same on the packed copy of the whole MSA, the part applied as a mask (rr_relative_vars_packed)
    pk = rr.Packed(msa, 0)
    for u_no in sorted(set(int(x) for x in ut)):
        for cutoff, mingroup in ((3.0, 8), (6.0, 20)):
            want = o.relative_vars(ut, u_no, M, cutoff, mingroup)
            assert list(pk.relative_vars(ut, u_no, M, cutoff, mingroup)) == list(want), (u_no, cutoff, mingroup)
    pk.close()
    msa.close()
