"""The kernel bodies of csrc/rr_scan_bitset.cu (the AND+POPC variant of the hot path with the fused epilogue of
rr_device.cuh), rr_cliquer.cu, rr_relvars.cu and rr_kmeans.cu, compiled by the HOST compiler against the
stand-in CUDA header of tests/emu (every CUDA thread a pthread, __syncthreads and the warp primitives as barriers) and run
on small inputs against the oracle.  This checks the kernels' logic - indexing, staging, skips, reductions, tails - on the
CPU, in the build container, every round; it is test infrastructure, says nothing about speed and does not replace the GPU
tests.  All of these kernels are also validated on a B200 (tests/test_zz_gpu_*.py); what runs here is the same source."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import repeatresolver_b200 as rr
from conftest import ROOT, golden_msa
import oracle_lib as O
from test_oracle_cliquer import cliquer_cases
from test_oracle_kmeans import kmeans_cases
from test_oracle_relvars import partition_by_site, relvars_cases, window_codes

EMU_DIR = os.path.join(ROOT, "tests", "emu")
CSRC = os.path.join(ROOT, "repeatresolver_b200", "csrc")


@pytest.fixture(scope="module")
def emu():
    if os.environ.get("RR_EMU_LIB"):                  # e.g. an AddressSanitizer build of the same driver (tools/emu_asan.sh)
        out = os.environ["RR_EMU_LIB"]
        srcs = []
    else:
        out = os.path.join(EMU_DIR, "_build", "libemu.so")
        srcs = [os.path.join(EMU_DIR, f) for f in ("emu_driver.cpp", "cuda_runtime.h")] + \
               [os.path.join(CSRC, f) for f in ("rr_pack.cu", "rr_scan_bitset.cu", "rr_device.cuh", "rr_plan.cpp", "rr_cliquer.cu", "rr_relvars.cu",
                                                "rr_kmeans.cu", "rr_score.h", "rr_kmeans.h", "rr_kernels.h")]
    if not os.path.exists(out) or any(os.path.getmtime(s) > os.path.getmtime(out) for s in srcs):
        os.makedirs(os.path.dirname(out), exist_ok=True)
        # -Bsymbolic: librr_maxcorr.so (loaded RTLD_GLOBAL by the package) exports nvcc's host stubs under the very names of
        # the kernels; the emulation must call its own bodies, not those stubs
        subprocess.check_call(["g++", "-O1", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-Wl,-Bsymbolic", "-I" + EMU_DIR,
                               "-I" + CSRC, "-o", out, os.path.join(EMU_DIR, "emu_driver.cpp"), "-lpthread"])
    lib = C.CDLL(out)
    vp, i, d, u64, u32 = C.c_void_p, C.c_int, C.c_double, C.c_ulonglong, C.c_uint
    lib.emu_cliquer.argtypes = [vp, vp, vp, vp, i, vp, i, i, i, i, d, d, vp, vp, u64, vp]
    lib.emu_relvars_pairs.argtypes = [vp, vp, i, vp, i, vp, vp, i, vp, d, vp, vp, u32, vp]
    lib.emu_masked_sizes.argtypes = [vp, vp, C.c_longlong, i, vp]
    lib.emu_kmeans_sweeps.argtypes = [vp, i, i, i, vp, vp, vp]
    return lib


def pack_bits(codes):
    """the layout rr_k_pack_bits writes: bits[5N][W32], covbits[N][W32], bit r%32 of word r/32 = read r (any row order
    gives the same counts), W32 a multiple of 4"""
    R, N = codes.shape
    W32 = 4 * ((R + 127) // 128)
    member = np.zeros((5 * N, W32 * 32), dtype=np.uint8)
    for k in range(5):
        member[k::5, :R] = (codes == k).T
    cover = np.zeros((N, W32 * 32), dtype=np.uint8)
    cover[:, :R] = (codes < 5).T
    as_words = lambda m: np.ascontiguousarray(np.packbits(m, axis=1, bitorder="little").view(np.uint32))
    return as_words(member), as_words(cover), W32


def two_family_msa(R, N, seed, variant_every=3):
    """reads with random spans over N sites, two families that differ at every variant_every-th site, 4 % noise"""
    rng = np.random.default_rng(seed)
    base = rng.integers(0, 4, N)
    alt = (base + 1 + rng.integers(0, 3, N)) % 4
    fam = rng.random(R) < 0.35
    codes = np.where(fam[:, None] & (np.arange(N) % variant_every == 0)[None, :], alt[None, :], base[None, :]).astype(np.uint8)
    noise = rng.random((R, N)) < 0.04
    codes[noise] = rng.integers(0, 5, int(noise.sum()))
    start = rng.integers(0, N // 2, R)
    end = np.minimum(N - 1, start + rng.integers(N // 3, N, R))
    col = np.arange(N)[None, :]
    codes[(col < start[:, None]) | (col > end[:, None])] = 5
    return codes[np.argsort(start, kind="stable")]


def run_cliquer(emu, codes, queries, mincov, maxclique, greedy, anfang=0, ende=None):
    o = O.Oracle.from_codes(codes)
    bits, cov, W32 = pack_bits(codes)
    gs = o.gsize().astype(np.int32)
    lnf = rr.lnfact_table(codes.shape[0] + 2)
    q = np.ascontiguousarray(queries, dtype=np.int32)
    ende = codes.shape[1] if ende is None else ende
    cap = len(q) * 5 * codes.shape[1]
    cand = np.zeros(cap, dtype=rr.HIT_DTYPE)
    hits = np.zeros(cap, dtype=rr.HIT_DTYPE)
    counters = np.zeros(2, dtype=np.uint64)
    thr = min(greedy - 1e-9 * max(1.0, abs(greedy)), 97.89)
    rc = emu.emu_cliquer(bits.ctypes.data, cov.ctypes.data, gs.ctypes.data, lnf.ctypes.data, W32, q.ctypes.data, len(q),
                         anfang, ende, mincov // 4, greedy, thr, cand.ctypes.data, hits.ctypes.data,
                         cap, counters.ctypes.data)
    assert rc == 0 and counters[1] <= counters[0] <= cap
    cand, hits = cand[:int(counters[0])], hits[:int(counters[1])]
    # every listed pair carries the oracle's four counts (argument order of 1217: Group1 = the candidate)
    for h in cand[:: max(1, len(cand) // 200)]:
        assert [h["s"], h["gr1"], h["gr2"], h["cov"]] == o.counts(int(h["group"]), int(q[h["slot"]])), h
    members, scores, n = rr.cliquer_from_hits(q, hits, gs, mincov, maxclique, greedy)
    return o, members, scores, n, len(cand), len(hits)


def test_cliquer_count_kernel_on_the_golden_cases(emu):
    for name in ("tree_small", "saturated"):
        case = cliquer_cases()[name]
        codes = window_codes(golden_msa(name), case["von"], case["bis"])[:, :330]       # a slab and a partial one
        queries = [int(x) for x in case["queries"] if int(x) < 5 * codes.shape[1]][:5]
        o, members, scores, n, nc, nh = run_cliquer(emu, codes, queries, case["mincov"], case["maxclique"], case["greedy"])
        assert nh > 0
        for k, qq in enumerate(queries):
            m0, z0 = o.cliquer(qq, case["mincov"], case["maxclique"], case["greedy"])
            assert list(members[k, :n[k]]) == list(m0) and np.array_equal(scores[k, :n[k]], z0), (name, qq)


def test_cliquer_count_kernel_with_greedy_inside_the_saturation_band(emu):
    """a bound above 98 proves nothing (486), a bound below it prunes against greedy = 98.1: only saturated pairs survive"""
    case = cliquer_cases()["saturated"]
    codes = window_codes(golden_msa("saturated"), case["von"], case["bis"])[:, :200]
    o = O.Oracle.from_codes(codes)
    M0, _, _ = o.scan(case["mincov"])
    queries = [int(q) for q in np.argsort(-M0, kind="stable")[:6]]
    o, members, scores, n, nc, nh = run_cliquer(emu, codes, queries, case["mincov"], 6, 98.1)
    sat = 0
    for k, qq in enumerate(queries):
        m0, z0 = o.cliquer(qq, case["mincov"], 6, 98.1)
        assert list(members[k, :n[k]]) == list(m0) and np.array_equal(scores[k, :n[k]], z0), qq
        sat += len(m0) - 1
    assert sat > 0 and (scores[:, 1:][scores[:, 1:] > 0] > 98.1).all()


def test_cliquer_count_kernel_with_several_chunks_and_ragged_coverage(emu):
    """more than 1024 reads = two 32-word chunks per bitset; spans so that chunks are skipped on either side"""
    codes = two_family_msa(1300, 70, seed=41)
    o = O.Oracle.from_codes(codes)
    gs = o.gsize()
    cand = np.flatnonzero((gs > 40) & (gs < 600))
    queries = [int(x) for x in cand[:: len(cand) // 9][:9]] + [0]                      # 10 queries: partial last block
    o, members, scores, n, nc, nh = run_cliquer(emu, codes, queries, 30, 8, 3.0)
    sizes = set()
    for k, qq in enumerate(queries):
        m0, z0 = o.cliquer(qq, 30, 8, 3.0)
        assert list(members[k, :n[k]]) == list(m0) and np.array_equal(scores[k, :n[k]], z0), qq
        sizes.add(len(m0))
    assert max(sizes) == 8 and nh > 50
    # a sub-range of candidate columns
    o, members, scores, n, _, _ = run_cliquer(emu, codes, queries[:3], 30, 5, 2.0, anfang=11, ende=52)
    for k, qq in enumerate(queries[:3]):
        m0, z0 = o.cliquer(qq, 30, 5, 2.0, 11, 52)
        assert list(members[k, :n[k]]) == list(m0) and np.array_equal(scores[k, :n[k]], z0), qq


def run_relvars(emu, codes, ut, u_no, M, cutoff, mingroup, masked=False):
    """masked=False: the part's rows packed on their own (rr_relative_vars with RR_RELVARS_KERNEL=1);
    masked=True: the whole MSA packed once, the part as a bitset (rr_relative_vars_packed)"""
    part = ut == u_no
    cov_u = int(part.sum())
    if cov_u < mingroup:
        return np.zeros(0, dtype=np.int32), 0
    if masked:
        bits, _, W32 = pack_bits(codes)
        mbits = np.zeros(W32 * 32, dtype=np.uint8)
        mbits[:len(part)] = part
        umask = np.ascontiguousarray(np.packbits(mbits, bitorder="little").view(np.uint32))
        sub = codes[part]
        gu = np.stack([(sub == k).sum(0) for k in range(5)], 1).reshape(-1).astype(np.int32)
        head = min(len(gu), 480)                                          # one warp per group: the first 60 blocks are enough here
        got = np.zeros(head, dtype=np.int32)
        emu.emu_masked_sizes(bits.ctypes.data, umask.ctypes.data, head, W32, got.ctypes.data)
        assert np.array_equal(got, gu[:head])
        umask_ptr = umask.ctypes.data
    else:
        sub = np.ascontiguousarray(codes[part])
        bits, _, W32 = pack_bits(sub)
        gu = np.stack([(sub == k).sum(0) for k in range(5)], 1).reshape(-1).astype(np.int32)
        umask_ptr = None
    sel = rr.relative_vars_from_counts(M, gu, cov_u, cutoff, mingroup)
    if len(sel) == 0:
        return sel, 0
    first = np.searchsorted(sel, sel + 100, side="left").astype(np.int32)
    lnf = rr.lnfact_table(codes.shape[0] + 2)
    mark = np.zeros(len(sel), dtype=np.uint8)
    unsure = np.zeros((4096, 4), dtype=np.int32)
    count = np.zeros(1, dtype=np.uint32)
    emu.emu_relvars_pairs(bits.ctypes.data, umask_ptr, W32, sel.ctypes.data, len(sel), first.ctypes.data, gu.ctypes.data, cov_u,
                          lnf.ctypes.data, cutoff, mark.ctypes.data, unsure.ctypes.data, 4096, count.ctypes.data)
    assert count[0] <= 4096
    for a, b, s, _ in unsure[:int(count[0])]:                           # what rr_abi.cu does with the undecided pairs
        if rr.relative_score_host(int(s), int(gu[sel[b]]), int(gu[sel[a]]), cov_u) > cutoff:
            mark[a] = mark[b] = 1
    return sel[mark > 0], int(count[0])


@pytest.mark.parametrize("masked", [False, True], ids=["part_packed", "whole_msa_masked"])
def test_relvars_pair_kernel_on_the_golden_cases(emu, masked):
    checked = 0
    for name, case in sorted(relvars_cases().items()):
        codes = window_codes(golden_msa(name), case["von"], case["bis"])
        o = O.Oracle.from_codes(codes)
        M, _, _ = o.scan(case["mincov"])
        ut, _ = partition_by_site(codes, M)
        for u_no, want in case["parts"].items():
            got, _ = run_relvars(emu, codes, ut, int(u_no), M, case["cutoff"], case["mingroup"], masked)
            assert list(got) == want["vars"], (name, u_no)
            checked += len(got) > 0
    assert checked >= 4


@pytest.mark.parametrize("masked", [False, True], ids=["part_packed", "whole_msa_masked"])
def test_relvars_pair_kernel_with_several_chunks_and_a_cutoff_hit_exactly(emu, masked):
    codes = two_family_msa(1200, 90, seed=43, variant_every=4)
    o = O.Oracle.from_codes(codes)
    rng = np.random.default_rng(2)
    M = np.where(rng.random(5 * codes.shape[1]) < 0.8, 9.0, 0.0)
    ut = (rng.random(codes.shape[0]) < 0.1).astype(np.int32)             # part 0 holds ~1080 reads: two chunks
    for cutoff, mingroup in ((3.0, 8), (8.5, 30)):
        got, _ = run_relvars(emu, codes, ut, 0, M, cutoff, mingroup, masked)
        want = o.relative_vars(ut, 0, M, cutoff, mingroup)
        assert list(got) == list(want) and len(want) > 5, (cutoff, mingroup)
    # a cutoff equal to a score that occurs: the pair is undecided on the device and settled by the host (Z > cutoff is strict)
    sub = codes[ut == 0]
    gu = np.stack([(sub == k).sum(0) for k in range(5)], 1).reshape(-1)
    sel = rr.relative_vars_from_counts(M, gu, len(sub), 3.0, 8)
    a, b = int(sel[0]), int(sel[np.searchsorted(sel, sel[0] + 100)])
    ga, gb = sub[:, a // 5] == a % 5, sub[:, b // 5] == b % 5
    z = O.relative_score(int((ga & gb).sum()), int(gb.sum()), int(ga.sum()), len(sub))
    if z > 0.5:
        got, undecided = run_relvars(emu, codes, ut, 0, M, z, 8, masked)
        assert list(got) == list(o.relative_vars(ut, 0, M, z, 8)) and undecided >= 1


@pytest.mark.parametrize("tile_reads", [16, 64])      # rows per score panel
def test_kmeans_sweeps_on_the_golden_cases(emu, tile_reads):
    for name in sorted(kmeans_cases()):
        rel = relvars_cases()[name]
        codes = window_codes(golden_msa(name), rel["von"], rel["bis"])
        o = O.Oracle.from_codes(codes)
        M, _, _ = o.scan(rel["mincov"])
        ut, _ = partition_by_site(codes, M)
        msa = rr.MSA.from_cells(codes, codes=True)
        for key, want in kmeans_cases()[name].items():
            u_no, mingroup = (int(x) for x in key.split("/"))
            reads, sig = rr.kmeans_signatures(msa, ut, u_no, rel["parts"][str(u_no)]["vars"])
            n, scv = sig.shape
            best_j = np.zeros((n, 5), dtype=np.int32)
            cen = np.zeros_like(sig)
            cluster = np.zeros(n, dtype=np.int32)
            assert emu.emu_kmeans_sweeps(sig.ctypes.data, n, scv, tile_reads, best_j.ctypes.data, cen.ctypes.data, cluster.ctypes.data) == 0
            for i in range(0, n, max(1, n // 10)):
                assert list(best_j[i]) == list(rr.kmeans_top5_host(sig, i))
            final, split = rr.kmeans_finish(sig, cen, cluster, mingroup)
            after = ut.copy()
            after[reads] = final + ut.max() + 1
            assert split == want["split"] and list(after) == want["after"], (name, key)
            # the dissolution through the device's score table: the kernel's table for the clusters that can become admissible
            final_t, split_t = rr.debug.kmeans_finish_table(sig, cen, cluster, mingroup)
            assert split_t == split and np.array_equal(final_t, final), (name, key)
            J = np.array([j for j in range(n) if j == 0 or np.count_nonzero(cluster == j) >= 2], dtype=np.int32)
            S = np.full((n, len(J)), -1, dtype=np.int32)
            emu.emu_kmeans_scores.restype = None
            emu.emu_kmeans_scores(C.c_void_p(sig.ctypes.data), C.c_void_p(cen.ctypes.data), C.c_void_p(J.ctypes.data), C.c_int(len(J)),
                                  C.c_int(n), C.c_int(scv), C.c_void_p(S.ctypes.data))
            want_S = scv * 64 - np.array([[sum(bin(int(x)).count("1") for x in (cen[j] ^ sig[i])) for j in J] for i in range(n)])
            assert np.array_equal(S, want_S), (name, key)
        msa.close()


def test_kmeans_sweeps_over_several_tiles_and_panels(emu):
    """random signatures, 150 reads (three 64-read tiles on either side, the last one partial) of 70 words (three staged
    chunks of 32, the last one partial), panels of 37 and 150 rows: the five kept reads against the host rule, the
    centroids against the majority, the assignment against numpy, and the dissolution's table for a gathered cluster list"""
    rng = np.random.default_rng(23)
    n, scv = 150, 70
    fam = rng.integers(0, 2 ** 63, (5, scv), dtype=np.int64).astype(np.uint64)
    noise = (rng.integers(0, 2 ** 63, (n, scv), dtype=np.int64).astype(np.uint64) & rng.integers(0, 2 ** 63, (n, scv), dtype=np.int64).astype(np.uint64)
             & rng.integers(0, 2 ** 63, (n, scv), dtype=np.int64).astype(np.uint64))
    sig = np.ascontiguousarray(fam[rng.integers(0, 5, n)] ^ noise)
    sig[7] = sig[3]                                                          # equal scores: the order of the reads decides
    sig[140] = sig[3]
    pop = np.array([bin(x).count("1") for x in range(256)], dtype=np.int64)

    def match(a, b):                                                         # [..., scv] uint64 -> 64 * scv - Hamming distance
        x = np.ascontiguousarray(a ^ b)
        return scv * 64 - pop[x.view(np.uint8)].reshape(x.shape[:-1] + (-1,)).sum(-1)

    results = []
    for panel_rows in (37, 150):
        best_j = np.full((n, 5), -7, dtype=np.int32)
        cen = np.zeros_like(sig)
        cluster = np.full(n, -7, dtype=np.int32)
        assert emu.emu_kmeans_sweeps(sig.ctypes.data, n, scv, panel_rows, best_j.ctypes.data, cen.ctypes.data, cluster.ctypes.data) == 0
        results.append((best_j, cen, cluster))
    best_j, cen, cluster = results[0]
    for other in results[1:]:
        assert all(np.array_equal(a, b) for a, b in zip(results[0], other))
    for i in range(n):
        assert list(best_j[i]) == list(rr.kmeans_top5_host(sig, i)), i
        b = best_j[i]
        assert all(int(cen[i, z]) == rr.kmeans_majority5_host(*(int(sig[k, z]) for k in b)) for z in (0, 31, 32, 69))
    scores = match(cen[None, :, :], sig[:, None, :])                         # [i][j]
    for i in range(n):
        row = scores[i].copy()
        row[i] = -1                                                          # not itself (2717)
        want = int(np.argmax(row)) if row.max() > 0 else 0                   # first best
        assert cluster[i] == want, i
    J = np.array([0, 3, 64, 65, 127, 128, 149], dtype=np.int32)
    S = np.full((n, len(J)), -1, dtype=np.int32)
    emu.emu_kmeans_scores.restype = None
    emu.emu_kmeans_scores(C.c_void_p(sig.ctypes.data), C.c_void_p(cen.ctypes.data), C.c_void_p(J.ctypes.data), C.c_int(len(J)), C.c_int(n),
                          C.c_int(scv), C.c_void_p(S.ctypes.data))
    assert np.array_equal(S, scores[:, J])


def test_kmeans_signature_kernel_equals_the_host_signatures(emu):
    """rr_k_km_signatures (one warp per read and 32 groups, one ballot per half word) against rr_kmeans_signatures on the
    golden parts - code cells and raw characters, group counts that leave a partial half word, an empty last word
    (n_vars % 64 == 0) and no groups at all"""
    vp, i = C.c_void_p, C.c_int
    emu.emu_kmeans_signatures.argtypes = [vp, i, i, vp, i, i, i, vp]
    emu.emu_kmeans_signatures.restype = None
    rng = np.random.default_rng(17)
    checked = 0
    for name in sorted(kmeans_cases())[:1]:
        rel = relvars_cases()[name]
        codes = window_codes(golden_msa(name), rel["von"], rel["bis"])
        o = O.Oracle.from_codes(codes)
        M, _, _ = o.scan(rel["mincov"])
        ut, _ = partition_by_site(codes, M)
        chars = np.frombuffer(b"ACGT- ", dtype=np.uint8)[codes]
        chars[::3] = np.frombuffer(b"acgt_n", dtype=np.uint8)[codes[::3]]        # lower case, '_' and an unknown symbol
        for as_codes, cells in ((1, codes), (0, chars)):
            msa = rr.MSA.from_cells(cells, codes=bool(as_codes))
            for u_no in sorted(set(int(k.split("/")[0]) for k in kmeans_cases()[name]))[:1]:
                base = np.array(rel["parts"][str(u_no)]["vars"], dtype=np.int32)[:200]
                sets = (base, base[:0], np.sort(rng.choice(5 * codes.shape[1], 128, replace=False)).astype(np.int32),
                        np.sort(rng.choice(5 * codes.shape[1], 301, replace=False)).astype(np.int32))
                for vars_ in (sets if as_codes else sets[2:]):
                    reads, want = rr.kmeans_signatures(msa, ut, u_no, vars_)
                    rows = np.ascontiguousarray(cells[reads])
                    got = np.full_like(want, 0xdeadbeefdeadbeef)
                    emu.emu_kmeans_signatures(rows.ctypes.data, cells.shape[1], as_codes, vars_.ctypes.data, len(vars_), len(reads),
                                              want.shape[1], got.ctypes.data)
                    assert np.array_equal(got, want), (name, as_codes, u_no, len(vars_))
                    checked += 1
            msa.close()
    assert checked >= 5


# ---- the AND+POPC variant of the scan itself (rr_k_scan_bitset with the fused epilogue of rr_device.cuh) ---------------
def ranked(codes):
    """rows in the order rr_pack leaves them: (length class, span start, span end); returns codes, start, end, class_start"""
    R, N = codes.shape
    cov = codes < 5
    start = cov.argmax(1).astype(np.int32)
    end = (N - 1 - cov[:, ::-1].argmax(1)).astype(np.int32)
    perm, _, cs = rr.rank_rows(start, end)
    return np.ascontiguousarray(codes[perm]), start[perm].copy(), end[perm].copy(), np.ascontiguousarray(cs, dtype=np.int32)


def run_scan_bitset(emu, codes, mincov, flags=0, blocks=3, part=(0, 1)):
    rc, start, end, cs = ranked(codes)
    R, N = rc.shape
    bits, _, W32 = pack_bits(rc)
    gs = np.stack([(rc == k).sum(0) for k in range(5)], 1).reshape(-1).astype(np.int32)
    coverage = (rc < 5).sum(0).astype(np.int32)
    brk = np.ascontiguousarray(rr.breakcols_from_spans(start, end, N, mincov), dtype=np.int32)
    lnf = rr.lnfact_table(R + 2)
    best = np.zeros(5 * N, dtype=[("z", "<u8"), ("p", "<u8")])
    best["p"] = np.uint64(2 ** 64 - 1)
    counters = np.zeros(8, dtype=np.uint64)
    pairs = emu.emu_scan_bitset(R, N, W32, mincov, flags, bits.ctypes.data, gs.ctypes.data, coverage.ctypes.data, brk.ctypes.data,
                                start.ctypes.data, end.ctypes.data, cs.ctypes.data, len(cs) - 1, lnf.ctypes.data, best.ctypes.data, counters.ctypes.data,
                                blocks, part[0], part[1])
    M = best["z"].view(np.float64).copy()
    A = np.where(best["p"] == np.uint64(2 ** 64 - 1), -1, best["p"].astype(np.int64)).astype(np.int32)
    return M, A, int(pairs), counters


@pytest.fixture(scope="module")
def emu_scan(emu):
    vp, i, u32 = C.c_void_p, C.c_int, C.c_uint
    emu.emu_scan_bitset.restype = C.c_longlong
    emu.emu_scan_bitset.argtypes = [i, i, i, i, u32, vp, vp, vp, vp, vp, vp, vp, i, vp, vp, vp, i, i, i]
    return emu


@pytest.mark.parametrize("name,cov", [("kat_appendix_g", 20), ("kat_appendix_g", 44), ("tree_small", 10), ("ragged", 20)])
def test_bitset_scan_kernel_on_golden_cases(emu_scan, name, cov, tmp_path):
    """the hot path's second implementation, kernel body and fused epilogue, on the CPU against the oracle: pair count,
    maxima (same libm on both sides: equal as doubles), arg-max; with and without pruning"""
    text = golden_msa(name)
    o = O.Oracle.from_text(text, tmp_path)
    msa = rr.MSA.from_text(text)
    cells = msa.cells().copy()
    table = np.full(256, 5, dtype=np.uint8)
    for ch, k in ((b"aA", 0), (b"cC", 1), (b"gG", 2), (b"tT", 3), (b"-_", 4)):
        for c in ch:
            table[c] = k
    codes = table[cells]
    msa.close()
    M0, A0, P0 = o.scan(cov)
    for flags in (0, rr.FLAG_NO_PRUNE):
        M, A, pairs, counters = run_scan_bitset(emu_scan, codes, cov, flags)
        assert pairs == P0 and int(counters[0]) == P0
        assert np.array_equal(M, M0), np.abs(M - M0).max()
        assert np.array_equal(A, A0)
        if flags:
            assert int(counters[1]) >= int((M0 > 0).sum()) // 2            # exact evaluations happened


def test_bitset_scan_kernel_length_classes_and_parts(emu_scan):
    """1 300 reads: several length classes of rows (boundaries on whole 256-row blocks), word ranges per class, and a two-way partition of the
    row blocks whose element-wise max (ties to the smaller partner) is the full result"""
    codes = two_family_msa(1300, 75, seed=47)
    cs = ranked(codes)[3]
    assert len(cs) >= 3 and (cs[1:-1] % 256 == 0).all() and 0 < cs[1] < 1300
    o = O.Oracle.from_codes(codes)
    M0, A0, P0 = o.scan(30)
    assert (M0 > 0).sum() > 50
    M, A, pairs, counters = run_scan_bitset(emu_scan, codes, 30)
    assert pairs == P0 == int(counters[0]) and np.array_equal(M, M0) and np.array_equal(A, A0)
    assert int(counters[1]) < P0                                          # pruning skipped exact evaluations
    Mm, Am, Pm = np.zeros_like(M0), np.full_like(A0, -1), 0
    for part in range(2):
        Mp, Ap, pp, cp = run_scan_bitset(emu_scan, codes, 30, part=(part, 2))
        assert pp == int(cp[0])
        better = (Mp > Mm) | ((Mp == Mm) & (Mp > 0) & (Ap >= 0) & ((Am < 0) | (Ap < Am)))
        Mm, Am = np.where(better, Mp, Mm), np.where(better, Ap, Am)
        Pm += pp
    assert Pm == P0 and np.array_equal(Mm, M0) and np.array_equal(Am, A0)


def test_bitset_scan_kernel_on_a_window_of_the_real_pipeline_msa(emu_scan):
    """1 000 columns from the middle of the MSA the reference's own pipeline produced (tests/golden/real_pipeline_msareal):
    real coverage structure - 312 reads with spans of 20 to 15 909 columns, a third of the cells blank"""
    text = golden_msa("real_pipeline_msareal")
    rows = text.split(b"\n")[:-1]
    cells = np.frombuffer(b"".join(rows), dtype=np.uint8).reshape(len(rows), -1)[:, 9000:10000]
    table = np.full(256, 5, dtype=np.uint8)
    for ch, k in ((b"aA", 0), (b"cC", 1), (b"gG", 2), (b"tT", 3), (b"-_", 4)):
        for c in ch:
            table[c] = k
    codes = table[cells]
    codes = np.ascontiguousarray(codes[(codes < 5).any(1)])              # reads that reach the window
    assert codes.shape[0] > 150
    o = O.Oracle.from_codes(codes)
    M0, A0, P0 = o.scan(30)
    assert P0 > 2e5 and (M0 > 3).sum() > 20
    M, A, pairs, counters = run_scan_bitset(emu_scan, codes, 30, blocks=4)
    assert pairs == P0 == int(counters[0]) and np.array_equal(M, M0) and np.array_equal(A, A0)


# ---- the packing kernels (rr_pack.cu) and the whole device pipeline of the bitset variant ---------------------------------
@pytest.fixture(scope="module")
def emu_pack(emu_scan):
    vp, i, i64 = C.c_void_p, C.c_int, C.c_longlong
    e = emu_scan
    e.emu_row_spans.argtypes = [vp, i, i, i, vp, vp, vp]
    e.emu_pack_bits.argtypes = [vp, vp, i, i, i, vp, vp, i, i, i]
    e.emu_bitset_sizes.argtypes = [vp, i64, i, vp]
    e.emu_pair_counts.argtypes = [vp, vp, i, i64, vp, vp, vp]
    e.emu_general_break.argtypes = [vp, i, i, i, vp]
    e.emu_bits_to_operand.argtypes = [vp, i64, i, vp, i64, i]
    return e


def device_pack(e, cells, codes_flag):
    """what pack_impl (rr_abi.cu) does, with the kernels emulated: spans -> row order -> bitsets -> sizes"""
    cells = np.ascontiguousarray(cells, dtype=np.uint8)
    R, N = cells.shape
    start, end, ncov = (np.zeros(R, dtype=np.int32) for _ in range(3))
    e.emu_row_spans(cells.ctypes.data, R, N, codes_flag, start.ctypes.data, end.ctypes.data, ncov.ctypes.data)
    perm, _, cs = rr.rank_rows(start, end)
    perm = perm.astype(np.int32)
    cs = np.ascontiguousarray(cs, dtype=np.int32)
    W32 = 4 * ((R + 127) // 128)
    bits = np.zeros((5 * N, W32), dtype=np.uint32)
    cov = np.zeros((N, W32), dtype=np.uint32)
    # as a row-sliced pack (rr_pack_rows): two slices of the rows, each packed into buffers of its own, merged by OR
    cut = R // 3
    for lo, hi in ((0, cut), (cut, R)):
        b1, c1 = np.zeros_like(bits), np.zeros_like(cov)
        part = np.ascontiguousarray(cells[lo:hi])
        if hi > lo:
            e.emu_pack_bits(part.ctypes.data, perm.ctypes.data, R, N, codes_flag, b1.ctypes.data, c1.ctypes.data, W32, lo, hi)
        assert not (bits & b1).any() and not (cov & c1).any()          # disjoint: OR = sum
        bits |= b1
        cov |= c1
    whole_b, whole_c = np.zeros_like(bits), np.zeros_like(cov)
    e.emu_pack_bits(cells.ctypes.data, perm.ctypes.data, R, N, codes_flag, whole_b.ctypes.data, whole_c.ctypes.data, W32, 0, R)
    assert np.array_equal(whole_b, bits) and np.array_equal(whole_c, cov)
    gs = np.zeros(5 * N, dtype=np.int32)
    cv = np.zeros(N, dtype=np.int32)
    e.emu_bitset_sizes(bits.ctypes.data, 5 * N, W32, gs.ctypes.data)
    e.emu_bitset_sizes(cov.ctypes.data, N, W32, cv.ctypes.data)
    return dict(R=R, N=N, W32=W32, start=start, end=end, ncov=ncov, perm=perm, cs=cs, bits=bits, cov=cov, gs=gs, cv=cv)


def test_packing_kernels_on_raw_text(emu_pack, tmp_path):
    """raw characters (lower case, '_' gaps, junk) through rr_k_row_spans / rr_k_pack_bits / rr_k_bitset_sizes /
    rr_k_pair_counts / rr_k_general_break / rr_k_bits_to_operand against numpy and the oracle"""
    text = golden_msa("ragged")
    o = O.Oracle.from_text(text, tmp_path)
    msa = rr.MSA.from_text(text)
    cells = msa.cells().copy()
    msa.close()
    table = np.full(256, 5, dtype=np.uint8)
    for ch, k in ((b"aA", 0), (b"cC", 1), (b"gG", 2), (b"tT", 3), (b"-_", 4)):
        for c in ch:
            table[c] = k
    codes = table[cells]
    p = device_pack(emu_pack, cells, 0)
    covd = codes < 5
    assert np.array_equal(p["start"], covd.argmax(1)) and np.array_equal(p["end"], codes.shape[1] - 1 - covd[:, ::-1].argmax(1))
    assert np.array_equal(p["ncov"], covd.sum(1))
    want_bits, want_cov, W32 = pack_bits(codes[p["perm"]])
    assert W32 == p["W32"] and np.array_equal(p["bits"], want_bits) and np.array_equal(p["cov"], want_cov)
    assert np.array_equal(p["gs"], o.gsize()) and np.array_equal(p["cv"], o.coverage())
    # the four counts of explicit pairs
    rng = np.random.default_rng(3)
    gi = rng.integers(0, 5 * p["N"], 300).astype(np.int32)
    gj = rng.integers(0, 5 * p["N"], 300).astype(np.int32)
    out = np.zeros((300, 4), dtype=np.int32)
    emu_pack.emu_pair_counts(p["bits"].ctypes.data, p["cov"].ctypes.data, W32, 300, gi.ctypes.data, gj.ctypes.data, out.ctypes.data)
    for k in range(0, 300, 3):
        assert list(out[k]) == o.counts(int(gi[k]), int(gj[k]))
    # the general first-break kernel against the closed form for single-span rows
    # (on the first 90 sites only: a warp walks the columns one shuffle reduction at a time, slow under emulation)
    n90 = 90
    brk = np.zeros(n90, dtype=np.int32)
    emu_pack.emu_general_break(p["cov"].ctypes.data, W32, n90, 20, brk.ctypes.data)
    shared = covd[:, :n90].T.astype(np.int32) @ covd[:, :n90].astype(np.int32)       # |C[ii] & C[jj]|
    for ii in range(n90):
        below = [jj for jj in range(ii + 20, n90) if shared[ii, jj] < 20]
        assert brk[ii] == (below[0] if below else max(n90, ii + 20)), ii                # MaxCorrelation.c:804-810
    # the 0/2 operands of the tensor path (a product is 4, see rr_scan_umma.cu): int8 and packed e2m1 (2.0 = 0b0100), K-major
    Kp = 256 * ((p["R"] + 255) // 256)
    member = np.zeros((5 * p["N"], Kp), dtype=np.uint8)
    rc = codes[p["perm"]]
    for k in range(5):
        member[k::5, :p["R"]] = (rc == k).T
    xb = np.zeros((5 * p["N"], Kp), dtype=np.int8)
    emu_pack.emu_bits_to_operand(p["bits"].ctypes.data, 5 * p["N"], p["W32"], xb.ctypes.data, Kp, 0)
    assert np.array_equal(xb.view(np.uint8), member * 2)
    x4 = np.zeros((5 * p["N"], Kp // 2), dtype=np.uint8)
    emu_pack.emu_bits_to_operand(p["bits"].ctypes.data, 5 * p["N"], p["W32"], x4.ctypes.data, Kp, 1)
    assert np.array_equal(x4, (member[:, 0::2] * 4) | (member[:, 1::2] * 4 << 4))


@pytest.mark.parametrize("name,cov", [("kat_appendix_g", 30), ("initial_aligner_style", 30)])
def test_whole_device_pipeline_of_the_bitset_variant(emu_pack, name, cov, tmp_path):
    """cells as read from the text -> spans -> row order -> bitsets -> sizes -> host plan -> scan kernel: every device
    step of the hot path's AND+POPC variant emulated, result = the oracle's and, formatted, the unmodified reference's file"""
    from conftest import golden_maxcorrs
    text = golden_msa(name)
    o = O.Oracle.from_text(text, tmp_path)
    msa = rr.MSA.from_text(text)
    cells = msa.cells().copy()
    msa.close()
    p = device_pack(emu_pack, cells, 0)
    R, N = p["R"], p["N"]
    start, end = p["start"][p["perm"]].copy(), p["end"][p["perm"]].copy()
    brk = np.ascontiguousarray(rr.breakcols_from_spans(start, end, N, cov), dtype=np.int32)
    lnf = rr.lnfact_table(R + 2)
    best = np.zeros(5 * N, dtype=[("z", "<u8"), ("p", "<u8")])
    best["p"] = np.uint64(2 ** 64 - 1)
    counters = np.zeros(8, dtype=np.uint64)
    pairs = emu_pack.emu_scan_bitset(R, N, p["W32"], cov, 0, p["bits"].ctypes.data, p["gs"].ctypes.data, p["cv"].ctypes.data,
                                     brk.ctypes.data, start.ctypes.data, end.ctypes.data, p["cs"].ctypes.data, len(p["cs"]) - 1, lnf.ctypes.data,
                                     best.ctypes.data, counters.ctypes.data, 3, 0, 1)
    M = best["z"].view(np.float64)
    M0, A0, P0 = o.scan(cov)
    assert pairs == P0 == int(counters[0]) and np.array_equal(M, M0)
    assert O.fmt_lines(M) == golden_maxcorrs(name, cov)


def test_bitset_scan_kernel_edge_shapes(emu_scan):
    """shapes without a single pair test (fewer rows than the coverage floor, fewer than 21 columns, one row, one base
    everywhere), and small random shapes with unusual coverage floors"""
    rng = np.random.default_rng(0)
    for codes, cov in [(np.zeros((5, 100), np.uint8), 30), (rng.integers(0, 5, (64, 20)).astype(np.uint8), 4),
                       (np.zeros((1, 50), np.uint8), 1), (np.zeros((40, 60), np.uint8), 10)]:
        M, A, pairs, counters = run_scan_bitset(emu_scan, codes, cov)
        assert pairs == 0 and int(counters[0]) == 0 and not M.any() and (A == -1).all()
    for R, N, cov in [(64, 21, 4), (33, 45, 1), (70, 40, 0), (130, 64, 9)]:
        codes = rng.integers(0, 6, (R, N)).astype(np.uint8)
        codes[:, : N // 2][rng.random((R, N // 2)) < 0.5] = 0             # some structure: a dominant base
        keep = (codes < 5).any(1)
        codes = codes[keep]
        # rows must be single spans for the closed-form first-break: blank only outside [first, last]
        cov_mask = codes < 5
        first, last = cov_mask.argmax(1), codes.shape[1] - 1 - cov_mask[:, ::-1].argmax(1)
        inner = (np.arange(codes.shape[1])[None, :] >= first[:, None]) & (np.arange(codes.shape[1])[None, :] <= last[:, None])
        codes = np.where(inner & (codes == 5), 4, codes).astype(np.uint8)
        o = O.Oracle.from_codes(codes)
        M0, A0, P0 = o.scan(cov)
        M, A, pairs, counters = run_scan_bitset(emu_scan, codes, cov)
        assert pairs == P0 == int(counters[0]) and np.array_equal(M, M0) and np.array_equal(A, A0), (R, N, cov)


@pytest.mark.parametrize("kind,seed", [("Distributed", 52), ("EquiDistant", 53)])
def test_bitset_scan_kernel_copy_families_with_exact_ties(emu_scan, kind, seed):
    """BASELINE.json configs[2] in miniature: the other two copy-difference structures, with a stretch of columns
    duplicated so that different partners attain exactly the same maximum - the smallest partner must win whatever the
    order in which the (really concurrent) emulated threads reach the 16-byte compare-and-swap"""
    g = rr.MsaGen(type=kind, copies=4, coverage=14, repeat_len=260, diff=0.04, seed=seed, flank=120, min_overlap=50)
    codes = g.codes()
    codes[:, 200:215] = codes[:, 90:105]
    cov = codes < 5                                                       # keep rows single spans after the copy
    first, last = cov.argmax(1), codes.shape[1] - 1 - cov[:, ::-1].argmax(1)
    col = np.arange(codes.shape[1])[None, :]
    inner = (col >= first[:, None]) & (col <= last[:, None])
    codes = np.where(inner & (codes == 5), 4, np.where(inner, codes, 5)).astype(np.uint8)
    o = O.Oracle.from_codes(codes)
    M0, A0, P0 = o.scan(12)
    # a partner inside the source stretch has a twin in the copy that attains the same score: the smaller id won
    ties = int(((M0 > 0) & (A0 >= 90 * 5) & (A0 < 105 * 5)).sum())
    M, A, pairs, counters = run_scan_bitset(emu_scan, codes, 12, blocks=5)
    assert pairs == P0 == int(counters[0]) and np.array_equal(M, M0) and np.array_equal(A, A0)
    assert ties > 0 and not ((A0 >= 200 * 5) & (A0 < 215 * 5) & (np.arange(len(A0)) < 90 * 5)).any()


def test_pruning_tiers_of_the_fused_epilogue_never_drop_a_pair_that_matters(emu):
    """rr_tier1_f32 (FP32, float table, rounding margin) and rr_tier2 (windowed partial sum) of rr_device.cuh, called as
    rr_scan_umma.cu calls them, at depths up to 40 000 reads: a pair whose exact score reaches the running maximum must
    survive both; against a maximum well above the score most pairs must be dropped (the tiers do prune)"""
    vp = C.c_void_p
    emu.emu_tiers.argtypes = [vp, C.c_int, C.c_longlong, vp, vp, vp, vp, vp]
    rng = np.random.default_rng(23)
    for max_cov in (300, 4000, 40000):
        lnf = rr.lnfact_table(max_cov + 2)
        quads, scores = [], []
        while len(quads) < 4000:
            cov = int(rng.integers(max(2, max_cov // 50), max_cov + 1))
            gr1 = int(rng.integers(1, cov + 1)) if rng.random() < 0.5 else int(rng.integers(1, max(2, cov // 20)))
            gr2 = int(rng.integers(1, cov + 1)) if rng.random() < 0.5 else int(rng.integers(1, max(2, cov // 20)))
            lo, hi = max(1, gr1 + gr2 - cov), min(gr1, gr2)
            if lo > hi:
                continue
            mean = gr1 * gr2 / cov
            s = int(min(hi, max(lo, round(mean + rng.normal() * 3 * (mean ** 0.5 + 1)))))      # around the mean and in the tail
            z = O.score(s, gr1, gr2, cov, gr1 + 3, gr2 + 5)
            quads.append((s, gr1, gr2, cov))
            scores.append(z)
        q = np.array(quads, dtype=np.uint32)
        z = np.array(scores)
        k1, k2, kq = np.zeros(len(q), dtype=np.uint8), np.zeros(len(q), dtype=np.uint8), np.zeros(len(q), dtype=np.uint8)
        # at the lower end of the support the score is exactly 0 and the kernels drop the pair before the tiers (tier 1b)
        live = z > 0
        for best in (z, z * (1 - 1e-12), z * 0.999, np.maximum(z - 1e-6, 0)):
            best = np.ascontiguousarray(best)                            # keep the buffer alive across the call
            emu.emu_tiers(lnf.ctypes.data, max_cov, len(q), q.ctypes.data, best.ctypes.data, k1.ctypes.data, k2.ctypes.data, kq.ctypes.data)
            assert kq[live].all(), "rr_tier1_q (the scaled form the tcgen05 kernel uses) dropped a pair that matters"
            assert (kq != k1).mean() < 0.01          # the two forms differ only where round(mean) != floor(mean) matters
            assert k1[live].all() and k2[live].all(), (max_cov, int((~k1[live].astype(bool)).sum()), int((~k2[live].astype(bool)).sum()))
        far = np.ascontiguousarray(z + 3.0)
        emu.emu_tiers(lnf.ctypes.data, max_cov, len(q), q.ctypes.data, far.ctypes.data, k1.ctypes.data, k2.ctypes.data, kq.ctypes.data)
        unsat = z < 90
        assert kq[unsat].mean() < 0.5
        assert k1[unsat].mean() < 0.5 and (k1 & k2)[unsat].mean() < 0.3, (max_cov, k1[unsat].mean(), (k1 & k2)[unsat].mean())


def test_lower_bound_of_the_deferred_evaluation_is_sound_and_tight(emu):
    """rr_tier2_interval (rr_device.cuh): the value the scan kernel raises a group's maximum by BEFORE the pair's exact
    score exists must never exceed that score (it prunes other pairs and must lose against the pair's own exact value),
    at depths up to 40 000 reads, near the mean, in the tail and in the saturated range; and it must be tight where the
    score matters (within 1e-3 relative + 1e-3 for most pairs above the mean), or the thresholds it gives are useless"""
    vp = C.c_void_p
    emu.emu_tier2_interval.argtypes = [vp, C.c_longlong, vp, vp, vp, vp]
    rng = np.random.default_rng(29)
    for max_cov in (60, 300, 4000, 40000):
        lnf = rr.lnfact_table(max_cov + 2)
        quads, scores = [], []
        while len(quads) < 5000:
            cov = int(rng.integers(max(2, max_cov // 50), max_cov + 1))
            gr1 = int(rng.integers(1, cov + 1)) if rng.random() < 0.5 else int(rng.integers(1, max(2, cov // 20)))
            gr2 = int(rng.integers(1, cov + 1)) if rng.random() < 0.5 else int(rng.integers(1, max(2, cov // 20)))
            lo, hi = max(1, gr1 + gr2 - cov), min(gr1, gr2)
            if lo > hi:
                continue
            mean = gr1 * gr2 / cov
            kind = rng.random()
            if kind < 0.5:
                s = int(min(hi, max(lo, round(mean + abs(rng.normal()) * 4 * (mean ** 0.5 + 1)))))   # in the upper tail
            elif kind < 0.8:
                s = int(rng.integers(lo, hi + 1))                                                    # anywhere, incl. saturated
            else:
                s = hi                                                                               # the end of the support
            quads.append((s, gr1, gr2, cov))
            scores.append(O.score(s, gr1, gr2, cov, gr1 + 3, gr2 + 5))
        q = np.array(quads, dtype=np.uint32)
        z = np.array(scores)
        keep = np.zeros(len(q), dtype=np.uint8)
        zlb = np.zeros(len(q), dtype=np.float64)
        best = np.zeros(len(q), dtype=np.float64)                         # no maximum yet: everything survives
        emu.emu_tier2_interval(lnf.ctypes.data, len(q), q.ctypes.data, best.ctypes.data, keep.ctypes.data, zlb.ctypes.data)
        assert keep.all()
        assert (zlb >= 0).all() and (zlb <= 98.0).all()
        bad = zlb > z
        assert not bad.any(), (max_cov, q[bad][:5], zlb[bad][:5], z[bad][:5])
        assert (zlb < z)[zlb > 0].all()                                   # strictly below: the exact value must win the update
        useful = (z > 1.0) & (z < 98.0)
        tight = zlb[useful] >= z[useful] * (1 - 1e-3) - 1e-3
        assert tight.mean() > 0.9, (max_cov, tight.mean())
        sat = z >= 98.0
        if sat.any():
            assert (zlb[sat] == 98.0).mean() > 0.9                        # saturated scores give the clamp (432: 98 + F > 98)
        # against a maximum equal to the pair's own score the pair must survive (it may tie or beat it)
        emu.emu_tier2_interval(lnf.ctypes.data, len(q), q.ctypes.data, np.ascontiguousarray(z).ctypes.data, keep.ctypes.data, zlb.ctypes.data)
        assert keep[z > 0].all()


def test_relvars_default_composition_with_the_emulated_kernels(emu_pack):
    """what rr_relative_vars does by default, step for step, with its device steps emulated: the part's rows packed as an
    MSA of their own (row spans, bitsets, sizes = |G & U|), |Gi & Gj & U| as the first count of rr_k_pair_counts for
    (later group, earlier group), selection and marks by the library's host half"""
    case = relvars_cases()["saturated"]
    codes = window_codes(golden_msa("saturated"), case["von"], case["bis"])
    o = O.Oracle.from_codes(codes)
    M, _, _ = o.scan(case["mincov"])
    ut, _ = partition_by_site(codes, M)
    for u_no, want in case["parts"].items():
        sub = np.ascontiguousarray(codes[ut == int(u_no)])
        if len(sub) > 200:                                               # rr_k_row_spans is one block per read: the small part is enough
            continue
        p = device_pack(emu_pack, sub, 1)
        sel = rr.relative_vars_from_counts(M, p["gs"], len(sub), case["cutoff"], case["mingroup"])
        first = np.searchsorted(sel, sel + 100, side="left")
        pa = np.concatenate([np.full(len(sel) - f, a) for a, f in enumerate(first)]).astype(np.int64)
        pb = np.concatenate([np.arange(f, len(sel)) for f in first]).astype(np.int64)
        gi, gj = sel[pb].astype(np.int32), sel[pa].astype(np.int32)      # Group1 = the later group (2465)
        out = np.zeros((len(gi), 4), dtype=np.int32)
        emu_pack.emu_pair_counts(p["bits"].ctypes.data, p["cov"].ctypes.data, p["W32"], len(gi), gi.ctypes.data, gj.ctypes.data, out.ctypes.data)
        S = np.zeros((len(sel), len(sel)), dtype=np.int32)
        S[pa, pb] = out[:, 0]
        got = rr.relative_vars_from_counts(M, p["gs"], len(sub), case["cutoff"], case["mingroup"], lambda s_: S)
        assert list(got) == want["vars"], u_no


# ---- CliqueGroup / CliqueCoverage (rr_k_clique_members + rr_k_rank_bits_to_rows, csrc/rr_cliquer.cu) ---------------------
def test_clique_group_kernels_against_the_oracle(emu_pack):
    """bit-sliced member counts and the rank -> row transposition on a packed MSA with several 32-read words, cliques of 0 to
    100 members (repeats allowed: the reference counts a group as often as it is listed), every kind of cutoff"""
    rng = np.random.default_rng(5)
    codes = two_family_msa(150, 60, seed=9)
    R, N = codes.shape
    p = device_pack(emu_pack, codes, 1)
    W32 = p["W32"]
    sizes = [0, 1, 2, 5, 12, 31, 32, 64, 100]
    stride = 100
    members = np.full((len(sizes) * 3, stride), -1, dtype=np.int32)
    nm = np.zeros(len(members), dtype=np.int32)
    cut = np.zeros(len(members), dtype=np.int32)
    for k, n in enumerate(sizes * 3):
        members[k, :n] = rng.integers(0, 5 * N, n)
        if n >= 5:
            members[k, 1] = members[k, 0]                                   # a repeated member
        nm[k] = n
        cut[k] = [-1, 0, max(n // 3, 1)][k // len(sizes)] if k % 7 else [127, n, n - 1][k // len(sizes)]
    rank_of_row = np.zeros(R, dtype=np.int32)
    rank_of_row[p["perm"]] = np.arange(R, dtype=np.int32)
    sc = R // 64 + 1
    vp, i, i64 = C.c_void_p, C.c_int, C.c_longlong
    emu_pack.emu_clique_members.argtypes = [vp, vp, i, i64, vp, i, vp, vp, i, vp, i, i, vp, vp]
    for which, fn in ((0, O.clique_group), (1, O.clique_coverage)):
        tmp = np.zeros((len(members), W32), dtype=np.uint32)
        out = np.full((len(members), 2 * sc), 0xdeadbeef, dtype=np.uint32)
        emu_pack.emu_clique_members(p["bits"].ctypes.data, p["cov"].ctypes.data, W32, len(members), members.ctypes.data, stride,
                                    nm.ctypes.data, cut.ctypes.data, which, rank_of_row.ctypes.data, R, 2 * sc, tmp.ctypes.data,
                                    out.ctypes.data)
        got = out.view("<u8")
        for k in range(len(members)):
            want = O.bitset_words(fn(codes, members[k, :nm[k]], int(cut[k])))
            assert np.array_equal(got[k], want), (which, k, int(nm[k]), int(cut[k]))


# ---- Dropoff_Cutoff's member counts (rr_k_clique_sizes, csrc/rr_cliquer.cu) --------------------------------------------
@pytest.mark.parametrize("grid_x", [2])
def test_clique_sizes_kernel_against_the_oracle(emu_pack, grid_x):
    """sizes[k] = reads contained in more than k of a clique's first n members (RepeatResolver.c:1472-1486) on a packed MSA of
    several 32-read words (a grid that leaves warps without words; many words: the random-bitset test below), cliques of 0 to 100
    members with repeats; the cutoff rule on those counts equals the restatement's"""
    rng = np.random.default_rng(11)
    codes = two_family_msa(200, 40, seed=4)
    R, N = codes.shape
    p = device_pack(emu_pack, codes, 1)
    W32 = p["W32"]
    stride = 101
    ns = [0, 1, 2, 6, 7, 13, 30, 64, 100]
    members = np.full((len(ns), stride), -1, dtype=np.int32)
    for k, n in enumerate(ns):
        members[k, :n] = rng.integers(5, 5 * N, n)
        if n >= 6:
            members[k, 2] = members[k, 0]                                   # a repeated member counts twice
    nm = np.array(ns, dtype=np.int32)
    sizes = np.zeros((len(ns), stride), dtype=np.uint32)
    vp, i, i64 = C.c_void_p, C.c_int, C.c_longlong
    emu_pack.emu_clique_sizes.argtypes = [vp, i, i64, vp, i, vp, vp, i]
    emu_pack.emu_clique_sizes(p["bits"].ctypes.data, W32, len(ns), members.ctypes.data, stride, nm.ctypes.data, sizes.ctypes.data, grid_x)
    for k, n in enumerate(ns):
        score = np.zeros(R, dtype=np.int64)
        for g in members[k, :n]:
            score += codes[:, g // 5] == g % 5
        want = [int(np.count_nonzero(score > c)) for c in range(n)]
        assert list(sizes[k, :n]) == want, (k, n)
        assert not sizes[k, n:].any()
        if n > 5:
            assert rr.dropoff_cutoff_host(sizes[k, :n], R, 0) == O.dropoff_cutoff(codes, members[k], n, 0), (k, n)


def test_clique_sizes_kernel_walks_every_word(emu_pack):
    """random bitsets of 300 words: one block of 128 threads needs three passes, two blocks two (the second with idle warps)"""
    rng = np.random.default_rng(12)
    G, W32, stride = 60, 300, 41
    bits = rng.integers(0, 2 ** 32, (G, W32), dtype=np.uint64).astype(np.uint32) & rng.integers(0, 2 ** 32, (G, W32), dtype=np.uint64).astype(np.uint32)
    ns = [40, 7, 1, 0, 23]
    members = np.full((len(ns), stride), -1, dtype=np.int32)
    for k, n in enumerate(ns):
        members[k, :n] = rng.integers(0, G, n)
    nm = np.array(ns, dtype=np.int32)
    unpacked = np.unpackbits(bits.view(np.uint8), axis=1, bitorder="little").astype(np.int64)      # [G][32 * W32]
    vp, i, i64 = C.c_void_p, C.c_int, C.c_longlong
    emu_pack.emu_clique_sizes.argtypes = [vp, i, i64, vp, i, vp, vp, i]
    for grid_x in (1, 2):
        sizes = np.zeros((len(ns), stride), dtype=np.uint32)
        emu_pack.emu_clique_sizes(bits.ctypes.data, W32, len(ns), members.ctypes.data, stride, nm.ctypes.data, sizes.ctypes.data, grid_x)
        for k, n in enumerate(ns):
            score = unpacked[members[k, :n]].sum(axis=0) if n else np.zeros(32 * W32, dtype=np.int64)
            assert list(sizes[k, :n]) == [int(np.count_nonzero(score > c)) for c in range(n)], (grid_x, k)
            assert not sizes[k, n:].any()
