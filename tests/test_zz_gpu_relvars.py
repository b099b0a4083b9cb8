"""GPU parity tests (-m gpu) for SURVEY.md section 8f row 3: Relative_Vars (/root/reference/RepeatResolver.c:2424-2493)
through the C ABI - rr_relative_vars (the part's rows packed on the device, all pairs in the tiled kernel of
csrc/rr_relvars.cu) and rr_relative_vars_packed (the part as a mask over the packed whole MSA) - against the committed
output of the UNMODIFIED RepeatResolver.c (tests/golden/relvars.json) and the oracle on a fresh input.
Bar: identical group lists.  Sorts last so that a failure here cannot hide other tests under `-x`."""
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


def test_relative_vars_cutoffs_and_masked_form():
    """several cutoffs / minimum group sizes, and the same parts as masks over the packed whole MSA"""
    g = rr.MsaGen(type="Tree", copies=8, coverage=40, repeat_len=1500, diff=0.01, seed=23, flank=300)
    codes = g.codes()
    o = O.Oracle.from_codes(codes)
    M, _, _ = o.scan(30)
    ut, _ = partition_by_site(codes, M)
    msa = rr.MSA.from_cells(codes, codes=True)
    for u_no in sorted(set(int(x) for x in ut)):
        for cutoff, mingroup in ((3.0, 8), (6.0, 20), (1.0, 3)):
            want = o.relative_vars(ut, u_no, M, cutoff, mingroup)
            got = rr.Relative_Vars(msa, ut, u_no, M, cutoff, mingroup)
            assert list(got) == list(want), (u_no, cutoff, mingroup)
    # the same on the packed copy of the whole MSA, the part applied as a mask (rr_relative_vars_packed)
    pk = rr.Packed(msa, 0)
    for u_no in sorted(set(int(x) for x in ut)):
        for cutoff, mingroup in ((3.0, 8), (6.0, 20)):
            want = o.relative_vars(ut, u_no, M, cutoff, mingroup)
            assert list(pk.relative_vars(ut, u_no, M, cutoff, mingroup)) == list(want), (u_no, cutoff, mingroup)
    pk.close()
    msa.close()
