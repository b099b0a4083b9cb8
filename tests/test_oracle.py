"""Pins the oracle (oracle/maxcorr_oracle.c + oracle/gsl_shim.c): known answers of SURVEY.md
Appendix B, exact rational sums (mpmath), scipy, the committed outputs of the unmodified
reference program (tests/golden, made by oracle/gen_golden.py), and - where oracle/_ref was
built in this container - the reference binary itself on fresh random inputs."""
import math
import os
import subprocess

import numpy as np
import pytest

from conftest import GOLDEN_CASES, ROOT, golden_maxcorrs, golden_msa
import oracle_lib as O

# (s, gr1, gr2, cov, exact -log10 Q)
KAT = [
    (1, 8, 8, 4000, 1.7985408974079), (3, 10, 12, 300, 2.2973779135563), (12, 40, 60, 312, 1.2559344789595),
    (25, 300, 310, 4000, 0.42012337764016), (60, 350, 400, 4000, 4.9836598330043),
    (200, 1000, 900, 4100, 0.01707749507505), (19, 19, 19, 4000, 51.33545043587),
    (150, 3400, 160, 4000, 3.3673915363254), (40, 45, 50, 171, 23.326216873284),
    (2, 8, 3400, 4000, 5.0128826272675e-6),
]


def exact_score(s, gr1, gr2, cov):
    import mpmath as mp
    mp.mp.dps = 60
    tot = mp.mpf(0)
    den = mp.binomial(cov, gr1)
    for x in range(s, min(gr1, gr2) + 1):
        tot += mp.binomial(gr2, x) * mp.binomial(cov - gr2, gr1 - x) / den
    return float(-mp.log10(tot))


@pytest.mark.parametrize("s,gr1,gr2,cov,exact", KAT)
def test_shim_known_answers(s, gr1, gr2, cov, exact):
    z = O.score(s, gr1, gr2, cov)
    assert z == pytest.approx(exact, rel=2e-10, abs=1e-15)
    assert "%f" % z == "%f" % exact


def test_shim_saturation():
    # Appendix B last two rows: > 98 -> 98 + 2s/(|Gi|+|Gj|)
    assert O.score(300, 300, 320, 640, 300, 320) == 98.0 + 600.0 / 620.0
    assert O.score(900, 2000, 1800, 13700, 2000, 1800) == 98.0 + 1800.0 / 3800.0  # pdf underflow -> inf -> 99
    assert O.score(0, 5, 5, 100) == 0.0 and O.score(3, 0, 5, 100) == 0.0 and O.score(3, 5, 0, 100) == 0.0


def test_shim_vs_mpmath_random():
    rng = np.random.default_rng(5)
    worst = 0.0
    for _ in range(150):
        cov = int(rng.integers(20, 3000))
        gr1 = int(rng.integers(1, cov + 1))
        gr2 = int(rng.integers(1, cov + 1))
        lo, hi = max(1, gr1 + gr2 - cov), min(gr1, gr2)
        if lo > hi:
            continue
        s = int(rng.integers(lo, hi + 1))
        e = exact_score(s, gr1, gr2, cov)
        if e > 97.5 or e < 1e-3:
            continue
        z = O.score(s, gr1, gr2, cov, gr1, gr2)
        worst = max(worst, abs(z - e) / e)
    assert worst < 5e-10, worst


def test_shim_vs_scipy():
    from scipy.stats import hypergeom
    rng = np.random.default_rng(6)
    worst = 0.0
    for _ in range(300):
        cov = int(rng.integers(20, 14000))
        gr1 = int(rng.integers(1, cov + 1))
        gr2 = int(rng.integers(1, cov + 1))
        mean = gr1 * gr2 / cov
        s = int(min(min(gr1, gr2), max(1, max(gr1 + gr2 - cov, mean + rng.integers(0, 40)))))
        q = hypergeom.sf(s - 1, cov, gr2, gr1)
        if not (1e-90 < q < 0.99):
            continue
        z = O.score(s, gr1, gr2, cov, gr1, gr2)
        worst = max(worst, abs(z - (-math.log10(q))) / (-math.log10(q)))
    assert worst < 1e-9, worst


def test_lnfact_values():
    assert O.lnfact(0) == 0.0 and O.lnfact(1) == 0.0
    assert O.lnfact(170) == math.log(float(math.factorial(170)))
    for n in (171, 200, 1000, 13700, 400000):
        assert O.lnfact(n) == pytest.approx(math.lgamma(n + 1.0), rel=1e-13)


@pytest.mark.parametrize("name,cov", GOLDEN_CASES)
def test_oracle_reproduces_reference_outputs(name, cov, tmp_path):
    """byte-for-byte: restatement vs the committed output of the unmodified reference"""
    o = O.Oracle.from_text(golden_msa(name), tmp_path)
    M, A, P = o.scan(cov, threads=4)
    assert O.fmt_lines(M) == golden_maxcorrs(name, cov)
    # argmax instrumentation is self-consistent: the recorded partner reproduces the maximum
    gs = o.gsize()
    for g in np.nonzero(M > 0)[0][:200]:
        i, j = min(g, A[g]), max(g, A[g])
        c = o.counts(i, j)
        assert O.score(c[0], c[1], c[2], c[3], gs[i], gs[j]) == M[g]
    assert (A[M == 0] == -1).all()


def test_kat_appendix_g_pair_counts(tmp_path):
    o = O.Oracle.from_text(golden_msa("kat_appendix_g"), tmp_path)
    assert (o.R, o.N) == (48, 32)
    assert o.scan(30)[2] == 135 and o.scan(44)[2] == 1


def test_oracle_thread_count_independent(tmp_path):
    o = O.Oracle.from_text(golden_msa("tree_small"), tmp_path)
    a = o.scan(10, threads=1)
    b = o.scan(10, threads=7)
    assert (a[0] == b[0]).all() and (a[1] == b[1]).all() and a[2] == b[2]


REF_BIN = os.path.join(ROOT, "oracle", "_ref", "MaxCorrelation_ref")


@pytest.mark.skipif(not os.path.exists(REF_BIN), reason="oracle/_ref not built (no /root/reference here)")
@pytest.mark.parametrize("seed", [101, 102])
def test_oracle_vs_reference_binary_fresh_input(seed, tmp_path):
    from repeatresolver_b200 import MsaGen
    g = MsaGen(type=["Tree", "EquiDistant"][seed % 2], copies=5, coverage=14, repeat_len=500, diff=0.03, seed=seed,
               flank=300, min_overlap=50)
    p = tmp_path / "M"
    p.write_bytes(g.text())
    subprocess.run([REF_BIN, "M", "-c", "15", "-p", "2"], cwd=tmp_path, check=True, capture_output=True)
    ref = (tmp_path / "MaxCorrsOf_M").read_bytes()
    o = O.Oracle.load(str(p))
    assert O.fmt_lines(o.scan(15)[0]) == ref
