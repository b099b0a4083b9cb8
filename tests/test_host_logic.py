"""CPU-only tests of the product's host side: the C ABI loads and exports every declared
symbol, the parser keeps the reference's row rules, ln(n!) / score / bounds (host build of
csrc/rr_score.h) agree with the oracle, first-break sweep, CLI argument handling.  No scan
is run here: without a GPU the ABI must refuse (there is no CPU path)."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

import repeatresolver_b200 as rr
from repeatresolver_b200 import _lib
from conftest import GOLDEN_CASES, ROOT, golden_msa
import oracle_lib as O


def test_abi_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "rr_maxcorr.h")).read()
    declared = set(re.findall(r"\b(rr_[a-z0-9_]+)\s*\(", hdr))
    declared -= {"rr_last_error"} - set(_lib.ABI_SYMBOLS)
    assert declared == set(_lib.ABI_SYMBOLS), declared ^ set(_lib.ABI_SYMBOLS)
    for s in _lib.ABI_SYMBOLS:
        assert hasattr(_lib.lib, s), s
    hdr = open(os.path.join(ROOT, "include", "rr_msagen.h")).read()
    declared = set(re.findall(r"\b(rr_msagen_[a-z0-9_]+)\s*\(", hdr))
    assert declared == set(_lib.MSAGEN_SYMBOLS)
    for s in _lib.MSAGEN_SYMBOLS:
        assert hasattr(_lib.gen, s), s
    assert b"sm_100a" in _lib.lib.rr_version()


def test_no_cpu_fallback_without_device():
    if rr.device_count() > 0:
        pytest.skip("a GPU is present")
    m = rr.MSA.from_text(b"ACGT\nACGT\n")
    with pytest.raises(rr.RRError) as e:
        rr.Packed(m)
    assert e.value.code == -5
    with pytest.raises(rr.RRError) as e:
        rr.Parallel_AllMaxCorrsRechner(m, 30, 1)
    assert e.value.code == -5


@pytest.mark.parametrize("name", sorted({n for n, _ in GOLDEN_CASES}))
def test_parser_row_rules_match_oracle(name, tmp_path):
    text = golden_msa(name)
    m = rr.MSA.from_text(text)
    o = O.Oracle.from_text(text, tmp_path)
    assert (m.rows, m.cols) == (o.R, o.N)
    p = tmp_path / "f.msa"
    p.write_bytes(text)
    m2 = rr.MSA.read(str(p))
    assert (m2.cells() == m.cells()).all()


def test_parser_edge_cases(tmp_path):
    # first line fixes the width; other widths are skipped; last line without '\n' is dropped
    m = rr.MSA.from_text(b"ACGT\nAC\nACGTA\nA-GT\n\nTTTT")
    assert (m.rows, m.cols) == (2, 4)
    assert bytes(m.cells()[1]) == b"A-GT"
    # ... unless it is one longer, in which case its last character is cut (strlen-1 rule)
    m = rr.MSA.from_text(b"ACGT\nTTTTX")
    assert (m.rows, m.cols) == (2, 4) and bytes(m.cells()[1]) == b"TTTT"
    # a NUL ends the line (strlen semantics)
    m = rr.MSA.from_text(b"ACGT\nAC\x00T\nGGGG\n")
    assert m.rows == 2
    # single line without newline: width is strlen-1
    m = rr.MSA.from_text(b"ACGTA")
    assert (m.rows, m.cols) == (1, 4)
    # empty input
    m = rr.MSA.from_text(b"")
    assert (m.rows, m.cols) == (0, 0)
    with pytest.raises(rr.RRError) as e:
        rr.MSA.read(str(tmp_path / "does_not_exist"))
    assert e.value.code == -1 and "MA is missing." in str(e.value)


def _kept_rows(data: bytes):
    """the row rules of Einlesen (291, 299) restated on bytes: lines split at \n, strlen semantics at NUL"""
    lines = data.split(b"\n")
    lines = [l + b"\n" for l in lines[:-1]] + ([lines[-1]] if lines[-1] else [])
    if not lines:
        return 0, []
    sl = lambda l: len(l.split(b"\0")[0]) if b"\0" in l else len(l)
    cols = sl(lines[0]) - 1
    return max(cols, 0), [l[:cols] for l in lines if sl(l) - 1 == cols]


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_parser_parallel_index_matches_restated_rules(seed, tmp_path):
    """> 16 MB of text takes the multi-threaded line index; the file-backed MSA (no host copy until cells() is asked
    for) and the in-memory parse must keep exactly the rows the reference would"""
    rng = np.random.default_rng(seed)
    cols = 3000 + seed
    out = []
    for r in range(6500):
        k = rng.integers(0, 40)
        line = bytes(rng.choice(np.frombuffer(b"ACGT- ", dtype=np.uint8), cols).tobytes())
        if k == 0:
            line = line[: cols - 1 - int(rng.integers(0, 5))]                # too short
        elif k == 1:
            line = line + b"A" * int(rng.integers(1, 9))                      # too long
        elif k == 2:
            line = line[:100] + b"\0" + line[101:]                            # strlen stops early: dropped
        elif k == 3:
            line = line + b"\0garbage"                                       # strlen == cols (the newline comes after the NUL): dropped
        elif k == 4:
            line = b""
        out.append(line + b"\n")
    data = b"".join(out)
    if seed == 2:
        data = data[:-1]                                                     # no trailing newline on the last line
    assert len(data) > (1 << 24)
    cols_want, rows_want = _kept_rows(data)
    p = tmp_path / "MSAreal"
    p.write_bytes(data)
    for m in (rr.MSA.read(str(p)), rr.MSA.from_text(data)):
        assert m.cols == cols_want and m.rows == len(rows_want)
        c = m.cells()
        assert c.shape == (len(rows_want), cols_want)
        assert c.tobytes() == b"".join(rows_want)
        m.close()


def test_lnfact_table_bitwise_equal_to_oracle():
    t = rr.lnfact_table(20000)
    for n in list(range(0, 400)) + [1000, 4096, 13700, 19999]:
        assert t[n] == O.lnfact(n), n


def test_score_host_bitwise_equal_to_oracle():
    rng = np.random.default_rng(3)
    n = 0
    while n < 3000:
        cov = int(rng.integers(2, 5000))
        gr1 = int(rng.integers(1, cov + 1))
        gr2 = int(rng.integers(1, cov + 1))
        lo, hi = max(0, gr1 + gr2 - cov), min(gr1, gr2)
        s = int(rng.integers(lo, hi + 1))
        si, sj = gr1 + int(rng.integers(0, 50)), gr2 + int(rng.integers(0, 50))
        a, b = rr.score_host(s, gr1, gr2, cov, si, sj), O.score(s, gr1, gr2, cov, si, sj)
        assert a == b or (a != a and b != b), (s, gr1, gr2, cov, a, b)
        n += 1
    # saturated branch and the s = minimum-of-support corner (lower tail with pdf == 0)
    assert rr.score_host(300, 300, 320, 640, 300, 320) == O.score(300, 300, 320, 640, 300, 320) > 98
    assert rr.score_host(900, 2000, 1800, 13700, 2000, 1800) == 98.0 + 1800.0 / 3800.0
    assert rr.score_host(5, 10, 15, 20, 10, 15) == O.score(5, 10, 15, 20, 10, 15)


def test_pruning_bounds_are_upper_bounds():
    """the two bounds of rr_score.h never fall below the exact score (before caps)"""
    rng = np.random.default_rng(4)
    checked = med = 0
    for _ in range(20000):
        cov = int(rng.integers(2, 3000))
        gr1 = int(rng.integers(1, cov + 1))
        gr2 = int(rng.integers(1, cov + 1))
        lo, hi = max(1, gr1 + gr2 - cov), min(gr1, gr2)
        if lo > hi:
            continue
        mean = gr1 * gr2 / cov
        s = int(np.clip(int(mean + rng.normal() * 3 * max(1.0, mean ** 0.5)), lo, hi))
        # sizes chosen so that the saturated branch returns its largest possible value, 98 + 2s/(2s) = 99
        z = O.score(s, gr1, gr2, cov, s, s)
        u = rr.score_bound_host(s, gr1, gr2, cov)
        assert u >= z, (s, gr1, gr2, cov, z, u)
        checked += 1
        if rr.below_median_host(s, gr1, gr2, cov):
            assert z <= 0.30103001, (s, gr1, gr2, cov, z)
            med += 1
    assert checked > 5000 and med > 500


def test_score_at_the_lower_end_of_the_support_is_exactly_zero():
    """the kernels drop pairs with s = gr1 + gr2 - cov >= 1 without evaluating them: P[X >= s] = 1 there, and the GSL
    chain returns exactly 1 (the lower-tail sum starts from pdf(s - 1) = 0), i.e. a score of (minus) zero that can
    never replace a maximum - checked on the oracle's shim and on the library's host build of the score"""
    rng = np.random.default_rng(11)
    n = 0
    for _ in range(4000):
        cov = int(rng.integers(1, 6000))
        gr1 = int(rng.integers(1, cov + 1))
        gr2 = int(rng.integers(max(1, cov - gr1 + 1), cov + 1))   # gr1 + gr2 > cov
        s = gr1 + gr2 - cov
        assert 1 <= s <= min(gr1, gr2)
        for z in (O.score(s, gr1, gr2, cov, gr1 + 3, gr2 + 5), rr.score_host(s, gr1, gr2, cov, gr1 + 3, gr2 + 5)):
            assert z == 0.0, (s, gr1, gr2, cov, z)
        n += 1
    for cov, gr1 in ((2493, 7), (1961, 1688), (70, 1)):            # gr2 = cov: pairs seen at config 2
        assert O.score(gr1, gr1, cov, cov, 10, 4000) == 0.0 == rr.score_host(gr1, gr1, cov, cov, 10, 4000)
    assert n == 4000


def test_bound_covers_the_saturation_window():
    """raw scores in (98, 99) are replaced by 98 + F, which may be LARGER than the raw score: the effective
    bound must not drop below 99 there"""
    import math
    hits = 0
    for cov, gr1, gr2 in [(640, 300, 320), (700, 330, 340), (800, 380, 390), (1000, 470, 480), (900, 400, 450)]:
        for s in range(max(1, gr1 + gr2 - cov), min(gr1, gr2) + 1):
            raw = -math.log10(max(O.hyper_Q(s - 1, gr2, cov - gr2, gr1), 1e-300))
            if 97.5 < raw < 99.5:
                z = O.score(s, gr1, gr2, cov, s, s)       # F = 1 -> 99 when saturated
                assert rr.score_bound_host(s, gr1, gr2, cov) >= z
                hits += 98 < raw < 99
    assert hits >= 3


def test_median_bound_exhaustive_small_populations():
    for cov in range(2, 41):
        for gr1 in range(1, cov + 1):
            for gr2 in range(1, cov + 1):
                for s in range(max(1, gr1 + gr2 - cov), min(gr1, gr2) + 1):
                    z = O.score(s, gr1, gr2, cov, s, s)
                    if rr.below_median_host(s, gr1, gr2, cov):
                        assert z <= 0.30103001
                    assert rr.score_bound_host(s, gr1, gr2, cov) >= z


def _brute_breakcols(start, end, cols, mincov):
    out = np.zeros(cols, dtype=np.int32)
    for ii in range(cols):
        jj = ii + 20
        while jj < cols:
            cov = int(((start <= ii) & (end >= jj) & (end >= ii)).sum())
            if cov < mincov:
                break
            jj += 1
        out[ii] = max(jj, ii + 20)
    return out


@pytest.mark.parametrize("mincov", [0, 1, 3, 8])
def test_breakcols_from_spans(mincov):
    rng = np.random.default_rng(9 + mincov)
    cols, rows = 150, 40
    start = rng.integers(0, cols, rows).astype(np.int32)
    end = np.minimum(cols - 1, start + rng.integers(0, 120, rows)).astype(np.int32)
    start[5], end[5] = 2 ** 31 - 1, -1  # an uncovered row
    got = rr.breakcols_from_spans(start, end, cols, mincov)
    want = _brute_breakcols(start, end, cols, mincov)
    assert (np.minimum(got, np.maximum(cols, np.arange(cols) + 20)) == np.minimum(want, np.maximum(cols, np.arange(cols) + 20))).all()


@pytest.mark.parametrize("seed,kunit,ti,tj", [(1, 256, 24, 48), (2, 128, 24, 48), (3, 32, 8, 32), (4, 256, 24, 48)])
def test_contraction_ranges_cover_every_contributing_row(seed, kunit, ti, tj):
    """K-range skipping is exact only if every row that covers a site of the row block AND a site of the column block
    lies inside the rank ranges the plan keeps (one per length class of rows) - checked exhaustively on random
    spans with a few MSA-spanning rows, rows are ordered as rr_pack orders them"""
    rng = np.random.default_rng(seed)
    R, N = (1500, 2600) if seed != 4 else (700, 1900)   # seed 4: below 1024 rows -> a single class
    ln = np.minimum(N, (rng.gamma(2.0, 300.0, R) + 30).astype(np.int64))
    ln[rng.integers(0, R, 12)] = N                                        # rows spanning everything
    st = (rng.random(R) * (N - ln + 1)).astype(np.int64)
    en = st + ln - 1
    perm, cls, cs = rr.rank_rows(st, en)                                   # rr_pack's order and classes
    ncls = len(cs) - 1
    st, en = st[perm], en[perm]
    assert cs[0] == 0 and cs[-1] == R and (np.diff(cs) >= 0).all() and (cs[1:-1] % 256 == 0).all()
    if R < 1024:
        assert (cs[:-1] == 0).all()                                       # everything in the last class
    for c in range(ncls):
        assert (cls[cs[c]:cs[c + 1]] == c).all()
        assert (np.diff(st[cs[c]:cs[c + 1]]) >= 0).all()                  # each class sorted by span start
    k_lo, k_hi, nrb = rr.contraction_ranges(st, en, N, cs, ti, tj, kunit)
    assert nrb == (N - 20 + ti - 1) // ti
    rank = np.arange(R)
    checked = 0
    for rb in range(0, nrb, 7):
        s_lo, s_hi = rb * ti, min(rb * ti + ti - 1, N - 21)
        for cb in range((s_lo + 20) // tj, k_lo.shape[1], 5):
            c_lo = cb * tj
            contributes = (st <= s_hi) & (en >= max(c_lo, s_lo))         # covers some site of both blocks
            for c in range(ncls):
                r = rank[contributes & (cls == c)]
                if len(r):
                    assert k_lo[c, cb] * kunit <= r.min() and r.max() < k_hi[c, rb] * kunit, (rb, cb, c)
                    checked += 1
                # and the ranges stay inside their class (up to the k-unit rounding)
                assert k_hi[c, rb] * kunit <= cs[c + 1] + kunit - 1 and k_lo[c, cb] * kunit >= cs[c] - (cs[c] % kunit)
    assert checked > 40


def test_cli_without_gpu_reports_and_fails(tmp_path):
    exe = os.path.join(ROOT, "repeatresolver_b200", "bin", "MaxCorrelation")
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 0 and "Usage: ./MaxCorrelation MSApath <options>" in r.stdout
    r = subprocess.run([exe, "nope"], capture_output=True, text=True, cwd=tmp_path)
    assert r.returncode == 1 and "MA is missing." in r.stdout
    # -p 0 selects a different routine in the reference (AllMaxCorrsRechner, MaxCorrelation.c:1010-1013): refused, nothing written
    (tmp_path / "M0").write_bytes(golden_msa("kat_appendix_g"))
    r = subprocess.run([exe, "M0", "-c", "30", "-p", "0"], capture_output=True, text=True, cwd=tmp_path)
    assert r.returncode == 1 and "-p 0" in r.stderr and not (tmp_path / "MaxCorrsOf_M0").exists()
    if rr.device_count() == 0:
        (tmp_path / "M").write_bytes(golden_msa("kat_appendix_g"))
        r = subprocess.run([exe, "M", "-c", "30"], capture_output=True, text=True, cwd=tmp_path)
        assert r.returncode == 1 and "no CUDA device" in r.stderr
        assert not (tmp_path / "MaxCorrsOf_M").exists()


def test_msagen_deterministic_and_shaped():
    a = rr.MsaGen(copies=3, coverage=10, repeat_len=400, seed=5, flank=200, min_overlap=50, threads=1)
    b = rr.MsaGen(copies=3, coverage=10, repeat_len=400, seed=5, flank=200, min_overlap=50, threads=7)
    assert (a.rows, a.cols) == (b.rows, b.cols) and (a.codes() == b.codes()).all()
    codes = a.codes()
    assert codes.max() <= 5 and a.cols >= 400
    # every row is one contiguous span, with no leading / trailing gap
    for r in range(a.rows):
        cov = np.nonzero(codes[r] < 5)[0]
        assert len(cov) and cov[-1] - cov[0] + 1 == len(cov)
        assert codes[r, cov[0]] != 4 and codes[r, cov[-1]] != 4
    t = a.text().split(b"\n")
    assert len(t) == a.rows + 1 and set(b"".join(t)) <= set(b"ACGT- ")


# ---- binary side format of the result --------------------------------------------------------
def reference_text_window(text_path, von, bis):
    """MaxCorrsEinlesen (RepeatResolver.c:609-646) restated: line i is kept when von <= i/5 <= bis"""
    out = []
    with open(text_path) as f:
        for i, line in enumerate(f):
            if von <= i // 5 <= bis:
                out.append(float(line))
    return np.array(out)


def test_binary_side_format_round_trip_and_window(tmp_path):
    rng = np.random.default_rng(12)
    N = 57
    M = np.round(rng.uniform(0, 99, 5 * N), 9)
    M[rng.random(5 * N) < 0.3] = 0.0
    M[3] = 98.897959183673464                      # a saturated value with more than six decimals
    A = rng.integers(-1, 5 * N, 5 * N).astype(np.int32)
    tp, bp = str(tmp_path / "MaxCorrsOf_X"), str(tmp_path / "MaxCorrsBinOf_X")
    rr.MaxCorrsRausschreiben(M, tp)
    rr.MaxCorrsRausschreiben_bin(M, bp, A)
    assert os.path.getsize(bp) == 24 + 5 * N * 12
    for von, bis in ((0, N - 1), (0, 10 ** 6), (5, 9), (56, 56), (56, 70), (57, 80), (-3, 2), (9, 5)):
        Mb, Ab = rr.MaxCorrsEinlesen_bin(bp, von, bis)
        lo, hi = 5 * max(von, 0), min(5 * N, 5 * (bis + 1))
        assert np.array_equal(Mb, M[lo:hi]) and np.array_equal(Ab, A[lo:hi]), (von, bis)       # full precision
        Mt, _ = rr.MaxCorrsEinlesen_bin(bp, von, bis, as_text=True)
        assert np.array_equal(Mt, reference_text_window(tp, von, bis)), (von, bis)               # what the text reader sees
    assert rr.MaxCorrsEinlesen_bin(bp, 0, 0, as_text=True)[0][3] == 98.897959
    # without partners
    rr.MaxCorrsRausschreiben_bin(M, bp)
    assert os.path.getsize(bp) == 24 + 5 * N * 8
    Mb, Ab = rr.MaxCorrsEinlesen_bin(bp, 2, 3)
    assert np.array_equal(Mb, M[10:20]) and (Ab == -1).all()
    # empty result
    rr.MaxCorrsRausschreiben_bin(np.zeros(0), bp)
    assert len(rr.MaxCorrsEinlesen_bin(bp, 0, 5)[0]) == 0


def test_binary_side_format_rejects_other_files(tmp_path):
    p = tmp_path / "junk"
    p.write_bytes(b"0.000000\n" * 10)
    with pytest.raises(rr.RRError):
        rr.MaxCorrsEinlesen_bin(str(p), 0, 1)
    with pytest.raises(rr.RRError):
        rr.MaxCorrsEinlesen_bin(str(tmp_path / "missing"), 0, 1)
    good = tmp_path / "good"
    rr.MaxCorrsRausschreiben_bin(np.arange(50, dtype=np.float64), str(good), np.arange(50, dtype=np.int32))
    data = good.read_bytes()
    (tmp_path / "short").write_bytes(data[:24 + 100])
    with pytest.raises(rr.RRError):
        rr.MaxCorrsEinlesen_bin(str(tmp_path / "short"), 0, 9)
    (tmp_path / "noargs").write_bytes(data[:24 + 400 + 40])
    M, _ = rr.MaxCorrsEinlesen_bin(str(tmp_path / "noargs"), 0, 1)      # values intact, partners of the window intact
    assert list(M) == list(range(10))
    with pytest.raises(rr.RRError):
        rr.MaxCorrsEinlesen_bin(str(tmp_path / "noargs"), 5, 9)         # partners of that window are cut off


def test_dropoff_cutoff_host_rule():
    """rr_dropoff_cutoff_host (the host half of rr_group_refinement) against Dropoff_Cutoff's loop (RepeatResolver.c:1488-1509)
    restated on random non-increasing member counts, including zero denominators, ties and short cliques"""
    rng = np.random.default_rng(21)

    def rule(sizes, signumber, c):
        drop_c = max(1, c)
        start, min_drop = drop_c, 1000000.0
        for i in range(start, len(sizes) - 1):
            den = min(float(signumber) - float(sizes[i]), float(sizes[i]))
            if den > 0:
                drop = (float(sizes[i - 1]) - float(sizes[i + 1])) / den
                if drop < min_drop:
                    min_drop, drop_c = drop, i
        return drop_c, min_drop

    for trial in range(400):
        signumber = int(rng.integers(1, 3000))
        n = int(rng.integers(0, 101))
        sizes = np.sort(rng.integers(0, signumber + 1, n))[::-1].astype(np.uint32)
        if trial % 5 == 0 and n:
            sizes[:n // 2] = signumber                                      # every read in the first members: denominator 0
        if trial % 7 == 0 and n:
            sizes[n // 3:] = sizes[n // 3]                                  # a plateau: ties, the first minimum wins
        for c in (0, 1, 3):
            got = rr.dropoff_cutoff_host(sizes, signumber, c)
            assert got == rule(sizes, signumber, c), (trial, c)
    assert rr.dropoff_cutoff_host([], 10) == (1, 1000000.0)
    assert rr.GroupPrecision(np.array([0x3fffffff | (1 << 31), 0], dtype=np.uint64), 70) == (30 + 29, 1)


def test_kmeans_dissolution_through_the_score_table():
    """rr_kmeans looks the dissolution's scores up in a table over the clusters that can ever become admissible (initially two
    reads or more, and cluster 0, which collects the reads nothing scores for); the result must be the on-the-fly one's -
    random signatures with a few families, many one-read clusters, reads that match nothing, every mingroup"""
    rng = np.random.default_rng(33)
    for trial in range(12):
        anzahl = int(rng.integers(3, 260))
        scv = int(rng.integers(1, 5))
        fam = rng.integers(0, 2 ** 63, (4, scv), dtype=np.int64).astype(np.uint64)
        sig = fam[rng.integers(0, 4, anzahl)] ^ (rng.integers(0, 2 ** 63, (anzahl, scv), dtype=np.int64).astype(np.uint64)
                                                 & rng.integers(0, 2 ** 63, (anzahl, scv), dtype=np.int64).astype(np.uint64)
                                                 & rng.integers(0, 2 ** 63, (anzahl, scv), dtype=np.int64).astype(np.uint64))
        cen = sig.copy()
        if trial % 3 == 0:
            sig[::7] = 0
            cen[::5] = ~np.uint64(0)                                         # centroids some reads score 0 against
        hubs = rng.choice(anzahl, max(1, anzahl // 4), replace=False)
        cluster = np.where(rng.random(anzahl) < 0.7, hubs[rng.integers(0, len(hubs), anzahl)], rng.integers(0, anzahl, anzahl)).astype(np.int32)
        for mingroup in (2, 3, 5, 9):
            a, na = rr.kmeans_finish(sig, cen, cluster, mingroup)
            b, nb = rr.debug.kmeans_finish_table(sig, cen, cluster, mingroup)
            assert na == nb and np.array_equal(a, b), (trial, mingroup)


def test_text_writer_is_printf_exact(tmp_path):
    """rr_maxcorr_write formats "%f" itself (exact binary value to six decimals, ties to even, one block write per 32 k lines):
    byte-identical to printf - here glibc's snprintf through ctypes and Python's own correctly rounded "%f" - on exact ties,
    values around every rounding edge, subnormals, negatives, values beyond the integer path (>= 2^52, inf, nan: written through
    fprintf) and random doubles of every magnitude; rr_argmax_write likewise for "%d" """
    import ctypes as C
    rng = np.random.default_rng(5)
    vals = [0.0, -0.0, 0.0078125, 0.5, 1.5e-6, 2.5e-6, 0.5e-6, 0.4999999e-6, 98.897959, 99.0, 98.0, 1e-300, 5e-324, -1.25, 123456789.125,
            2.0 ** 52 - 1, 2.0 ** 52, 1e15, 1e22, float("inf"), float("-inf"), 4503599627370495.5, 0.0000015, 1.0000005, 9.9999995,
            99.9999995, 0.9999995]
    vals += [k / 2.0 ** n for n in range(1, 30) for k in (1, 3, 5, 7, 9, 11)]        # exact ties at the seventh decimal and near misses
    a = np.concatenate([np.array(vals), rng.random(150000) * 100, rng.random(50000) * 1e-4, 10.0 ** rng.uniform(-12, 14, 50000),
                        -rng.random(1000) * 50, np.round(rng.random(50000) * 100, 6) + 0.5e-6,
                        rng.integers(0, 2 ** 63, 20000).astype(np.uint64).view(np.float64)])
    a = np.ascontiguousarray(a[~np.isnan(a)], dtype=np.float64)
    p = str(tmp_path / "MaxCorrsOf_x")
    rr.MaxCorrsRausschreiben(a, p)
    got = open(p, "rb").read()
    assert got == b"".join(b"%f\n" % v for v in a)
    libc = C.CDLL(None)
    libc.snprintf.restype = C.c_int
    buf = C.create_string_buffer(512)
    lines = got.split(b"\n")
    for i in list(range(len(vals))) + list(rng.integers(0, len(a), 3000)):
        libc.snprintf(buf, C.c_size_t(512), b"%f", C.c_double(a[i]))
        assert lines[i] == buf.value, (i, a[i])
    rr.MaxCorrsRausschreiben(np.array([np.nan, 1.0]), p)                     # nan goes through fprintf as well
    assert open(p, "rb").read() in (b"nan\n1.000000\n", b"-nan\n1.000000\n")
    from repeatresolver_b200._lib import lib
    A = np.concatenate([[-1, 0, 9, 10, 2 ** 31 - 1, -2 ** 31], rng.integers(-1, 3_000_000, 70000)]).astype(np.int32)
    assert lib.rr_argmax_write(p.encode(), A.ctypes.data, len(A)) == 0
    assert open(p, "rb").read() == b"".join(b"%d\n" % v for v in A)


def _einlesen_file(tmp_path):
    """a result file with everything MaxCorrsEinlesen's loop can meet: plain "%f" lines, leading blanks, exponents, hex floats,
    inf / nan, trailing text, a line longer than fgets' 99 characters (counts as several lines) and no newline at the end"""
    rng = np.random.default_rng(2)
    lines = [b"%f" % v for v in rng.random(43) * 99] + [b"   7.250000", b"\t1e2", b"0x1.8p3", b"inf", b"-3.5abc", b"98.897959 trailing",
                                                        b"1" * 120 + b".5", b"0.000001", b"12"]
    p = tmp_path / "MaxCorrsOf_M"
    p.write_bytes(b"\n".join(lines))
    return str(p)


def test_maxcorrs_einlesen_window_and_line_rules(tmp_path):
    """rr_maxcorr_read_text / MaxCorrsEinlesen: fgets(100) lines, sscanf("%lf") values, the window von <= i/5 <= bis"""
    p = _einlesen_file(tmp_path)
    all_ = rr.MaxCorrsEinlesen(p, 0, 10 ** 6)
    # the 122-character line is two fgets lines: 99 ones, then 21 ones and ".5"
    assert len(all_) == 43 + 6 + 2 + 2
    assert all_[43] == 7.25 and all_[44] == 100.0 and all_[45] == 12.0 and np.isinf(all_[46]) and all_[47] == -3.5 and all_[48] == 98.897959
    assert all_[49] == float("1" * 99) and all_[50] == float("1" * 21 + ".5") and all_[51] == 0.000001 and all_[52] == 12.0
    for von, bis in ((0, 0), (2, 3), (9, 9), (10, 10), (10, 50), (11, 12), (3, 2)):
        want = all_[[i for i in range(len(all_)) if von <= i // 5 <= bis]]
        got = rr.MaxCorrsEinlesen(p, von, bis)
        assert np.array_equal(got, want, equal_nan=True), (von, bis)
    assert rr.MaxCorrsEinlesen(p + ".missing", 0, 3) is None                  # 619
    empty = tmp_path / "MaxCorrsOf_E"
    empty.write_bytes(b"\n\nabc\n")
    assert list(rr.MaxCorrsEinlesen(str(empty), 0, 0)) == [0.0, 0.0, 0.0]      # no number: 0.0 (the reference: uninitialised)
    # the text written by the library reads back to the values "%f" carries
    M = np.random.default_rng(3).random(500) * 99
    q = str(tmp_path / "MaxCorrsOf_Q")
    rr.MaxCorrsRausschreiben(M, q)
    assert np.array_equal(rr.MaxCorrsEinlesen(q, 0, 99), np.array([float("%f" % v) for v in M]))


@pytest.mark.skipif(not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "ref_einlesen_driver")),
                    reason="oracle/_ref is built in the build container only")
def test_maxcorrs_einlesen_against_the_unmodified_reference(tmp_path):
    import subprocess
    p = _einlesen_file(tmp_path)
    drv = os.path.join(ROOT, "oracle", "_ref", "ref_einlesen_driver")
    for von, bis in ((0, 10 ** 6), (0, 0), (2, 3), (9, 10), (10, 50)):
        got = rr.MaxCorrsEinlesen(p, von, bis)
        out = subprocess.run([drv, p, str(von), str(bis), "40", str(len(got))], capture_output=True, text=True)
        assert out.returncode == 0, out.stderr
        ref = np.array([float.fromhex(l) if "n" not in l.lower() or "inf" in l.lower() else float(l) for l in out.stdout.split()])
        assert np.array_equal(got, ref, equal_nan=True), (von, bis)
    out = subprocess.run([drv, p + ".missing", "0", "3", "40", "1"], capture_output=True, text=True)
    assert out.stdout.strip() == "NULL"


def test_coverage_restriction():
    """RepeatResolver.c:4001-4013 restated literally against the vectorised mirror"""
    rng = np.random.default_rng(9)
    for trial in range(20):
        N = int(rng.integers(1, 60))
        cov = rng.integers(0, 50, N)
        if trial % 4 == 0:
            cov[:] = 10 * int(rng.integers(1, 5))                            # exactly at the 9/10 edge for some columns
            cov[::3] = cov[0] * 9 // 10
        M = rng.random(5 * N) * 99
        want = M.copy()
        maxcov = 0
        for i in range(N):
            if cov[i] > maxcov:
                maxcov = int(cov[i])
        for i in range(5 * N):
            if int(cov[i // 5]) * 10 < maxcov * 9:
                want[i] = 0.0
        assert np.array_equal(rr.coverage_restriction(M, cov), want)
    assert len(rr.coverage_restriction([], [])) == 0
    with pytest.raises(ValueError):
        rr.coverage_restriction([1.0], [1])


def test_partition_file_and_completion(tmp_path):
    """Unterteilung_Rausschreiben / UnterteilungEinlesen / UnterteilungsKomplettierung (RepeatResolver.c:568-607, 1845-1865)"""
    u = np.array([3, 0, -1, 12, 7], dtype=np.int32)
    p = str(tmp_path / "KmeansSubdivisionOf_x")
    rr.Unterteilung_Rausschreiben(u, p)
    assert open(p, "rb").read() == b"3\n0\n-1\n12\n7"                        # no newline after the last (579-582)
    assert np.array_equal(rr.UnterteilungEinlesen(p), u)
    open(p, "wb").write(b" 4\n+5 x\n\nabc\n-6\n")
    assert list(rr.UnterteilungEinlesen(p)) == [4, 5, 0, 0, -6]
    assert rr.UnterteilungEinlesen(p + ".missing") is None
    aus = np.array([1, -1, 1, 1, -1, 1, 1])
    assert list(rr.UnterteilungsKomplettierung(u, aus)) == [3, -1, 0, -1, -1, 12, 7]
    with pytest.raises(ValueError):
        rr.UnterteilungsKomplettierung(u[:3], aus)
