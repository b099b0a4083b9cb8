"""GPU parity tests (-m gpu) for SURVEY.md section 8f row 4: Kmeans (/root/reference/RepeatResolver.c:2604-2821) through
the C ABI (rr_kmeans: signatures and dissolution on the host, the two read x read sweeps and the centroids in
csrc/rr_kmeans.cu) against the committed output of the UNMODIFIED RepeatResolver.c (tests/golden/kmeans.json) and the
oracle on a fresh input.  Bar: the partition after the call identical, integer for integer.
Everything these kernels share with the host is pinned on the CPU as well (tests/test_oracle_kmeans.py)."""
import numpy as np
import pytest

import repeatresolver_b200 as rr
from conftest import golden_msa
import oracle_lib as O
from test_oracle_kmeans import kmeans_cases
from test_oracle_relvars import partition_by_site, relvars_cases, window_codes

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", sorted(kmeans_cases()))
def test_kmeans_golden(name):
    rel = relvars_cases()[name]
    codes = window_codes(golden_msa(name), rel["von"], rel["bis"])
    o = O.Oracle.from_codes(codes)
    M, _, _ = o.scan(rel["mincov"])
    ut, _ = partition_by_site(codes, M)
    msa = rr.MSA.from_cells(codes, codes=True)
    for key, want in kmeans_cases()[name].items():
        u_no, mingroup = (int(x) for x in key.split("/"))
        n, after = rr.Kmeans(msa, ut, u_no, rel["parts"][str(u_no)]["vars"], mingroup)
        assert n == want["split"] and list(after) == want["after"], key
    msa.close()


def test_kmeans_fresh_input_against_oracle():
    g = rr.MsaGen(type="Tree", copies=12, coverage=40, repeat_len=2000, diff=0.02, seed=29, flank=400)
    codes = g.codes()
    o = O.Oracle.from_codes(codes)
    M, _, _ = o.scan(30)
    ut, _ = partition_by_site(codes, M)
    msa = rr.MSA.from_cells(codes, codes=True)
    checked = 0
    for u_no in sorted(set(int(x) for x in ut)):
        vars_ = o.relative_vars(ut, u_no, M, 3.0, 8)
        for v in (vars_, vars_[:63], vars_[:64], vars_[:1], vars_[:0]):      # word boundaries of the signatures, no groups at all
            for mingroup in (2, 5, 12):
                n0, want = o.kmeans(ut, u_no, v, mingroup)
                n, after = rr.Kmeans(msa, ut, u_no, v, mingroup)
                assert n == n0 and np.array_equal(after, want), (u_no, len(v), mingroup)
                checked += 1
    assert checked > 10
    msa.close()
