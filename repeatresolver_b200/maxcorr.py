"""Host-side mirror of the reference's MaxCorrelation interface, on top of the C ABI.

Names follow /root/reference/MaxCorrelation.c: reading an MSA is `Einlesen` (270-393), the
scan is `Parallel_AllMaxCorrsRechner` (839-908), writing is `MaxCorrsRausschreiben`
(516-532), and `MaxCorrelation(path, c, p)` is the program's `main` (916-1026).  Every call
goes through librr_maxcorr.so; nothing here computes.
"""
import ctypes as C
import os

import numpy as np

from . import _lib
from ._lib import CliquerStats, ScanOpts, ScanStats, lib

VARIANTS = {"auto": 0, "bitset": 1, "umma": 2, "umma_f4": 3, "umma_mxf4": 4}
VARIANT_NAMES = {v: k for k, v in VARIANTS.items()}
FLAG_NO_PRUNE = 1
FLAG_HOST_FINALIZE = 2
FLAG_GENERAL_BREAK = 4
FLAG_SEED_ONLY = 8
FLAG_SKIP_SEED = 16


class RRError(RuntimeError):
    def __init__(self, code, where):
        self.code = code
        super().__init__(f"{where} failed with {code}: {lib.rr_last_error().decode(errors='replace')}")


def _check(rc, where):
    if rc != 0:
        raise RRError(rc, where)


def variant_available(variant):
    return bool(lib.rr_variant_available(VARIANTS[variant] if isinstance(variant, str) else variant))


def device_count():
    return lib.rr_device_count()


class MSA:
    """The kept rows of an MSA (the reading half of Einlesen)."""

    def __init__(self, handle):
        self._h = C.c_void_p(handle)

    @classmethod
    def read(cls, path):
        h = C.c_void_p()
        _check(lib.rr_msa_read(os.fsencode(path), C.byref(h)), "rr_msa_read")
        return cls(h.value)

    @classmethod
    def read_window(cls, path, von, bis):
        """RepeatResolver.c's reader (Einlesen 293-429, rr_msa_read_window): (the MSA of the columns von .. bis of the reads with
        a symbol at both ends of the window, Ausgelassen: int8 per line of the file, 1 = kept, -1 = left out)"""
        h = C.c_void_p()
        n = C.c_int64(0)
        p = os.fsencode(path)
        rc = lib.rr_msa_read_window(p, int(von), int(bis), C.byref(h), None, 0, C.byref(n))
        _check(rc, "rr_msa_read_window")
        lib.rr_msa_free(h)                                                   # the first call counted the lines
        aus = np.zeros(n.value, dtype=np.int8)
        h = C.c_void_p()
        _check(lib.rr_msa_read_window(p, int(von), int(bis), C.byref(h), aus.ctypes.data, len(aus), C.byref(n)), "rr_msa_read_window")
        return cls(h.value), aus

    @classmethod
    def from_text(cls, text):
        if isinstance(text, str):
            text = text.encode("latin1")
        h = C.c_void_p()
        _check(lib.rr_msa_from_text(text, len(text), C.byref(h)), "rr_msa_from_text")
        return cls(h.value)

    @classmethod
    def from_cells(cls, cells, codes=True):
        cells = np.ascontiguousarray(cells, dtype=np.uint8)
        assert cells.ndim == 2
        h = C.c_void_p()
        _check(lib.rr_msa_from_cells(cells.ctypes.data, cells.shape[0], cells.shape[1], int(codes), C.byref(h)),
               "rr_msa_from_cells")
        return cls(h.value)

    @classmethod
    def alloc(cls, rows, cols, codes=True):
        h = C.c_void_p()
        _check(lib.rr_msa_alloc(rows, cols, int(codes), C.byref(h)), "rr_msa_alloc")
        return cls(h.value)

    @property
    def rows(self):
        return lib.rr_msa_rows(self._h)

    @property
    def cols(self):
        return lib.rr_msa_cols(self._h)

    def cells(self):
        """numpy view of the [rows][cols] cell matrix (host memory owned by the handle)."""
        n = self.rows * self.cols
        if n == 0:
            return np.zeros((self.rows, self.cols), dtype=np.uint8)
        buf = (C.c_uint8 * n).from_address(lib.rr_msa_cells(self._h))
        return np.frombuffer(buf, dtype=np.uint8).reshape(self.rows, self.cols)

    def close(self):
        if self._h:
            lib.rr_msa_free(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Packed:
    """One GPU's packed copy of an MSA (the packing half of Einlesen, on the device)."""

    def __init__(self, msa, device=0, rows=None):
        """rows = None: the whole MSA packed on `device` (rr_pack).  rows = (lo, hi): only those rows are uploaded
        (rr_pack_rows); the caller then exchanges the spans (slice_spans / set_spans), merges the bitsets of all slices
        (bits_device) and calls finish() - repeatresolver_b200.dist.pack_over_ranks does that between ranks."""
        h = C.c_void_p()
        self.rows, self.cols = msa.rows, msa.cols
        if rows is None:
            _check(lib.rr_pack(msa._h, device, C.byref(h)), "rr_pack")
            self.slice = (0, msa.rows)
        else:
            _check(lib.rr_pack_rows(msa._h, device, int(rows[0]), int(rows[1]), C.byref(h)), "rr_pack_rows")
            self.slice = (int(rows[0]), int(rows[1]))
        self._h = h

    def slice_spans(self):
        """[3][hi - lo] int32: first / last covered column and covered cells of the uploaded rows"""
        n = self.slice[1] - self.slice[0]
        sp = np.zeros((3, max(n, 1)), dtype=np.int32)
        _check(lib.rr_pack_slice_spans(self._h, sp[0].ctypes.data, sp[1].ctypes.data, sp[2].ctypes.data), "rr_pack_slice_spans")
        return sp[:, :n]

    def set_spans(self, spans):
        sp = np.ascontiguousarray(spans, dtype=np.int32)
        assert sp.shape == (3, self.rows)
        _check(lib.rr_pack_set_spans(self._h, sp[0].ctypes.data, sp[1].ctypes.data, sp[2].ctypes.data), "rr_pack_set_spans")

    def bits_device(self):
        """(device pointer, bytes) of this slice's group + coverage bitsets: to be OR-merged (integer SUM) over the slices"""
        p, n = C.c_void_p(), C.c_size_t(0)
        _check(lib.rr_pack_bits_device(self._h, C.byref(p), C.byref(n)), "rr_pack_bits_device")
        return p.value, n.value

    def finish(self):
        _check(lib.rr_pack_finish(self._h), "rr_pack_finish")

    def scan(self, mincov=30, variant="auto", flags=0, part_index=0, part_count=1):
        opts = ScanOpts(mincov, VARIANTS[variant] if isinstance(variant, str) else variant, flags, part_index, part_count)
        st = ScanStats()
        _check(lib.rr_scan(self._h, C.byref(opts), C.byref(st)), "rr_scan")
        return st.as_dict()

    def fetch(self):
        G = 5 * self.cols
        M = np.zeros(G, dtype=np.float64)
        A = np.zeros(G, dtype=np.int32)
        _check(lib.rr_scan_fetch(self._h, M.ctypes.data, A.ctypes.data), "rr_scan_fetch")
        return M, A

    def finalize(self, M, A):
        """RR_FLAG_HOST_FINALIZE as a call (rr_scan_finalize): the maxima of groups with a partner re-evaluated with the
        host libm from device-side counts, in place; M / A as fetched or merged over parts"""
        assert M.dtype == np.float64 and A.dtype == np.int32 and M.flags.c_contiguous and A.flags.c_contiguous
        assert len(M) == 5 * self.cols and len(A) == 5 * self.cols
        _check(lib.rr_scan_finalize(self._h, M.ctypes.data, A.ctypes.data), "rr_scan_finalize")
        return M

    def set_thresholds(self, thr):
        thr = np.ascontiguousarray(thr, dtype=np.float64)
        assert len(thr) == 5 * self.cols
        _check(lib.rr_scan_set_thresholds(self._h, thr.ctypes.data), "rr_scan_set_thresholds")

    def values_to_device(self, device_ptr):
        """write the 5N running maxima to a device buffer (e.g. torch_tensor.data_ptr())"""
        _check(lib.rr_scan_values_device(self._h, C.c_void_p(device_ptr)), "rr_scan_values_device")

    def set_thresholds_device(self, device_ptr):
        _check(lib.rr_scan_set_thresholds_device(self._h, C.c_void_p(device_ptr)), "rr_scan_set_thresholds_device")

    def cliquer(self, query_group, mincov=30, maxclique=30, greedy=3.0, anfang=0, ende=None):
        """Cliquer (RepeatResolver.c:1179-1240) for one query group, plain path (rr_cliquer): (members incl. the
        query, scores with scores[0] = 100)"""
        members = np.full(maxclique + 1, -1, dtype=np.int32)
        scores = np.zeros(maxclique, dtype=np.float64)
        n = C.c_int(0)
        _check(lib.rr_cliquer(self._h, int(query_group), anfang, 2 ** 30 if ende is None else ende, mincov, maxclique, greedy,
                              members.ctypes.data, scores.ctypes.data, C.byref(n)), "rr_cliquer")
        return members[:n.value].copy(), scores[:n.value].copy()

    def cliquer_batch(self, query_groups, mincov=30, maxclique=30, greedy=3.0, anfang=0, ende=None):
        """Cliquer for all query groups of a Group_Refinement pass (1647-1649) on the device (rr_cliquer_batch):
        (members [nq][maxclique+1] in the reference's -1 terminated layout, scores [nq][maxclique], n_members [nq],
        stats dict)"""
        q = np.ascontiguousarray(query_groups, dtype=np.int32).ravel()
        members = np.full((len(q), maxclique + 1), -1, dtype=np.int32)
        scores = np.zeros((len(q), maxclique), dtype=np.float64)
        n = np.zeros(len(q), dtype=np.int32)
        st = CliquerStats()
        _check(lib.rr_cliquer_batch(self._h, len(q), q.ctypes.data, anfang, 2 ** 30 if ende is None else ende, mincov, maxclique,
                                    greedy, members.ctypes.data, scores.ctypes.data, n.ctypes.data, C.byref(st)),
               "rr_cliquer_batch")
        return members, scores, n, st.as_dict()

    def clique_groups(self, cliques, cutoffs, want_groups=True, want_coverage=True):
        """CliqueGroup / CliqueCoverage (RepeatResolver.c:976-1008, 1064-1096) for a batch of cliques on the device
        (rr_clique_groups).  cliques: [n][stride] int32 in the reference's layout (members first, the first negative entry
        ends a clique, 986-993) or a list of member lists; cutoffs: [n].  Returns (groups, coverage): uint64 [n][rows/64+1]
        bitsets over the MSA's rows (read r = bit r % 64 of word r / 64, GrAdd 211-217), None for a kind not asked for."""
        if isinstance(cliques, np.ndarray) and cliques.ndim == 2:
            mem = np.ascontiguousarray(cliques, dtype=np.int32)
            neg = mem < 0
            nm = np.where(neg.any(1), neg.argmax(1), mem.shape[1]).astype(np.int32)
        else:
            lists = [[int(g) for g in c] for c in cliques]
            stride = max([len(c) for c in lists] + [1])
            mem = np.full((len(lists), stride), -1, dtype=np.int32)
            for k, c in enumerate(lists):
                mem[k, :len(c)] = c
            nm = np.array([len(c) for c in lists], dtype=np.int32)
        cut = np.ascontiguousarray(np.broadcast_to(np.asarray(cutoffs, dtype=np.int32), (len(mem),)), dtype=np.int32)
        sc = self.rows // 64 + 1
        G = np.zeros((len(mem), sc), dtype=np.uint64) if want_groups else None
        Cv = np.zeros((len(mem), sc), dtype=np.uint64) if want_coverage else None
        _check(lib.rr_clique_groups(self._h, len(mem), mem.ctypes.data, max(mem.shape[1], 1) if mem.ndim == 2 else 1, nm.ctypes.data,
                                    cut.ctypes.data, G.ctypes.data if G is not None else None, Cv.ctypes.data if Cv is not None else None),
               "rr_clique_groups")
        return G, Cv

    def group_refinement(self, MaxCorrs, cutoff, mincov=30, maxclique=30, greedy=3.0, anfang=0, ende=None, want_groups=True,
                         want_coverage=True):
        """Group_Refinement (RepeatResolver.c:1634-1693) for every group above the cutoff (rr_group_refinement): Cliquer,
        Sizes, Dropoff_Cutoff, CliqueGroup and CliqueCoverage.  MaxCorrs is not modified; returns a dict with
        MaxCorrs (the copy with the entries of groups whose Sizes <= 5 zeroed, 1685), groups [nq] (ascending), Cliques
        [nq][maxclique+1], Sizes, Cutoffs, Drop_Off [nq], C_Groups / C_Coverage uint64 [nq][rows/64+1] (all zero where the
        reference leaves NULL, i.e. Sizes <= 5; None if not asked for), stats."""
        M = np.array(MaxCorrs, dtype=np.float64)
        if len(M) != 5 * self.cols:
            raise ValueError("MaxCorrs must hold 5 * cols values")
        nq = int(np.count_nonzero(M > cutoff))
        sc = self.rows // 64 + 1
        groups = np.zeros(nq, dtype=np.int32)
        cliques = np.full((nq, maxclique + 1), -1, dtype=np.int32)
        sizes = np.zeros(nq, dtype=np.int32)
        cutoffs = np.zeros(nq, dtype=np.int32)
        drop = np.zeros(nq, dtype=np.float64)
        G = np.zeros((nq, sc), dtype=np.uint64) if want_groups else None
        Cv = np.zeros((nq, sc), dtype=np.uint64) if want_coverage else None
        n = C.c_int64(0)
        st = CliquerStats()
        _check(lib.rr_group_refinement(self._h, M.ctypes.data, float(cutoff), anfang, 2 ** 30 if ende is None else ende, mincov, maxclique,
                                       float(greedy), nq, groups.ctypes.data, cliques.ctypes.data, sizes.ctypes.data, cutoffs.ctypes.data,
                                       drop.ctypes.data, G.ctypes.data if G is not None else None,
                                       Cv.ctypes.data if Cv is not None else None, C.byref(n), C.byref(st)), "rr_group_refinement")
        assert n.value == nq
        return {"MaxCorrs": M, "groups": groups, "Cliques": cliques, "Sizes": sizes, "Cutoffs": cutoffs, "Drop_Off": drop,
                "C_Groups": G, "C_Coverage": Cv, "stats": st.as_dict()}

    def relative_vars(self, Unterteilung, u_no, MaxCorrs, cutoff, mingroup, with_pairs=False):
        """Relative_Vars (RepeatResolver.c:2424-2493) on this packed MSA, the part applied as a mask
        (rr_relative_vars_packed): ascending group ids"""
        u = np.ascontiguousarray(Unterteilung, dtype=np.int32)
        M = np.ascontiguousarray(MaxCorrs, dtype=np.float64)
        assert len(u) == self.rows and len(M) == 5 * self.cols
        out = np.zeros(5 * self.cols + 1, dtype=np.int32)
        n, pairs = C.c_int(0), C.c_int64(0)
        _check(lib.rr_relative_vars_packed(self._h, u.ctypes.data, int(u_no), M.ctypes.data, float(cutoff), int(mingroup),
                                           out.ctypes.data, C.byref(n), C.byref(pairs)), "rr_relative_vars_packed")
        return (out[:n.value].copy(), pairs.value) if with_pairs else out[:n.value].copy()

    def pair_counts(self, gi, gj):
        gi = np.ascontiguousarray(gi, dtype=np.int32)
        gj = np.ascontiguousarray(gj, dtype=np.int32)
        out = np.zeros((len(gi), 4), dtype=np.int32)
        _check(lib.rr_pair_counts(self._h, len(gi), gi.ctypes.data, gj.ctypes.data, out.ctypes.data), "rr_pair_counts")
        return out

    def timer_start(self):
        _check(lib.rr_timer_start(self._h), "rr_timer_start")

    def timer_stop(self):
        ms = C.c_float()
        _check(lib.rr_timer_stop(self._h, C.byref(ms)), "rr_timer_stop")
        return ms.value

    def sizes(self):
        gs = np.zeros(5 * self.cols, dtype=np.int32)
        cv = np.zeros(self.cols, dtype=np.int32)
        _check(lib.rr_packed_sizes(self._h, gs.ctypes.data, cv.ctypes.data), "rr_packed_sizes")
        return gs, cv

    def close(self):
        if self._h:
            lib.rr_packed_free(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def Einlesen(path, von=None, bis=None):
    """MaxCorrelation.c:270-393 (the whole file) or, with a window, RepeatResolver.c:293-429 (the columns von .. bis of the reads
    with a symbol at both ends of the window): the MSA, in the second form together with Ausgelassen (1 / -1 per line)"""
    if von is None and bis is None:
        return MSA.read(path)
    return MSA.read_window(path, 0 if von is None else von, 2 ** 31 - 2 if bis is None else bis)


def Parallel_AllMaxCorrsRechner(msa, mincov=30, n_gpus=1, variant="auto", flags=FLAG_HOST_FINALIZE):
    """MaxCorrelation.c:839-908 on n_gpus B200s.  Returns (MaxCorrs[5N], argmax[5N], stats)."""
    G = 5 * msa.cols
    M = np.zeros(G, dtype=np.float64)
    A = np.zeros(G, dtype=np.int32)
    st = ScanStats()
    _check(lib.rr_maxcorr_run(msa._h, mincov, n_gpus, VARIANTS[variant] if isinstance(variant, str) else variant,
                              flags, M.ctypes.data, A.ctypes.data, C.byref(st)), "rr_maxcorr_run")
    return M, A, st.as_dict()


def MaxCorrsRausschreiben(MaxCorrs, outputfile):
    """MaxCorrelation.c:516-532."""
    M = np.ascontiguousarray(MaxCorrs, dtype=np.float64)
    _check(lib.rr_maxcorr_write(os.fsencode(outputfile), M.ctypes.data, len(M)), "rr_maxcorr_write")


def MaxCorrsRausschreiben_bin(MaxCorrs, outputfile, argmax=None):
    """the binary side format "MaxCorrsBinOf_<MSA>" (rr_maxcorr_write_bin): full-precision values, optional partners"""
    M = np.ascontiguousarray(MaxCorrs, dtype=np.float64)
    A = None if argmax is None else np.ascontiguousarray(argmax, dtype=np.int32)
    assert A is None or len(A) == len(M)
    _check(lib.rr_maxcorr_write_bin(os.fsencode(outputfile), M.ctypes.data, None if A is None else A.ctypes.data, len(M)),
           "rr_maxcorr_write_bin")


def MaxCorrsEinlesen_bin(inputfile, von, bis, as_text=False):
    """MaxCorrsEinlesen (RepeatResolver.c:609-646) on the binary side format: the groups of columns von..bis inclusive,
    (MaxCorrs, argmax); as_text rounds the values as the "%f" text file would"""
    n = max(0, 5 * (bis - max(von, 0) + 1))
    M = np.zeros(n, dtype=np.float64)
    A = np.zeros(n, dtype=np.int32)
    got = C.c_int64(0)
    _check(lib.rr_maxcorr_read_bin(os.fsencode(inputfile), von, bis, int(as_text), M.ctypes.data, A.ctypes.data, C.byref(got)),
           "rr_maxcorr_read_bin")
    return M[:got.value].copy(), A[:got.value].copy()


def MaxCorrsEinlesen(inputfile, von, bis):
    """RepeatResolver.c:609-646: the values of MaxCorrsOf_<MSA> for the columns von..bis (inclusive; five lines per column),
    None if the file cannot be opened (619).  The reference sizes its array from the MSA window it has read; here the result
    has as many entries as the window holds lines (rr_maxcorr_read_text)."""
    n = C.c_int64(0)
    path = os.fsencode(inputfile)
    if lib.rr_maxcorr_read_text(path, int(von), int(bis), None, 0, C.byref(n)) != 0:
        return None
    out = np.zeros(n.value, dtype=np.float64)
    _check(lib.rr_maxcorr_read_text(path, int(von), int(bis), out.ctypes.data, len(out), C.byref(n)), "rr_maxcorr_read_text")
    return out[:n.value]


def MaxCorrelation(msa_path, c=30, p=1, variant="auto", flags=FLAG_HOST_FINALIZE, outdir=None):
    """The program: read <msa_path>, scan on p GPUs with coverage floor c, write
    MaxCorrsOf_<msa_path> (MaxCorrelation.c:991-993, 1014).  Returns (path written, stats)."""
    msa = Einlesen(msa_path)
    M, A, st = Parallel_AllMaxCorrsRechner(msa, c, max(1, p), variant, flags)
    name = "MaxCorrsOf_" + os.path.basename(msa_path) if outdir is not None else "MaxCorrsOf_" + msa_path
    out = os.path.join(outdir, name) if outdir is not None else name
    MaxCorrsRausschreiben(M, out)
    msa.close()
    return out, st


def Cliquer(packed, anfang, ende, mincov, maxclique, greedy, a):
    """RepeatResolver.c:1179-1240 with the reference's argument order: the clique of group `a` among the groups of
    columns [anfang, ende) as the reference's int[maxclique+1] (Clique[0] = a, best partner first, -1 after the last)."""
    members, _, _, _ = packed.cliquer_batch([a], mincov, maxclique, greedy, anfang, ende)
    return members[0]


def CliqueGroup(packed, Clique, c):
    """RepeatResolver.c:976-1008 with the reference's argument order: the reads contained in more than c of the clique's
    groups, as the reference's unsigned long[sc] bitset (uint64 array)."""
    return packed.clique_groups(np.asarray(Clique, dtype=np.int32).reshape(1, -1), [c], want_coverage=False)[0][0]


def CliqueCoverage(packed, Clique, c):
    """RepeatResolver.c:1064-1096: the reads covered at more than c of the sites of the clique's groups (uint64 bitset)."""
    return packed.clique_groups(np.asarray(Clique, dtype=np.int32).reshape(1, -1), [c], want_groups=False)[1][0]


def group_reads(bitset, rows):
    """the reads of a reference-layout bitset (GrElement, RepeatResolver.c:262-269) as ascending row numbers"""
    bits = np.unpackbits(np.ascontiguousarray(bitset, dtype="<u8").view(np.uint8), bitorder="little")[:rows]
    return np.flatnonzero(bits)


def Group_Refinement_Cliques(packed, MaxCorrs, cutoff, anfang, ende, mincov, maxclique, greedy):
    """The Cliquer calls of Group_Refinement (RepeatResolver.c:1638-1650) in one device batch: for every group i with
    MaxCorrs[i] > cutoff its clique and Sizes[i] (1650: members before the first entry <= 0).  Returns
    (groups [nq], Cliques [nq][maxclique+1], Sizes [nq], scores [nq][maxclique], stats)."""
    M = np.asarray(MaxCorrs, dtype=np.float64)
    groups = np.flatnonzero(M > cutoff).astype(np.int32)
    members, scores, _, st = packed.cliquer_batch(groups, mincov, maxclique, greedy, anfang, ende)
    sizes = np.array([int(np.argmax(row <= 0)) for row in members], dtype=np.int32)     # while(Cliques[i][Sizes[i]]>0)
    return groups, members, sizes, scores, st


def coverage_restriction(MaxCorrs, Coverage):
    """RepeatResolver.c:4001-4013, the step of its main between reading MaxCorrs and Group_Refinement: the values of every
    column covered by fewer than 9/10 of the deepest column's reads are set to 0 (Coverage[i/5] * 10 < maxcov * 9, integer
    arithmetic).  Coverage: reads per column (Packed.sizes()[1]).  Returns a copy."""
    M = np.array(MaxCorrs, dtype=np.float64)
    cov = np.asarray(Coverage, dtype=np.int64)
    if len(M) != 5 * len(cov):
        raise ValueError("MaxCorrs must hold 5 values per column")
    maxcov = max(int(cov.max()) if len(cov) else 0, 0)                       # 4002: starts at 0
    M[np.repeat(cov * 10 < maxcov * 9, 5)] = 0.0
    return M


def Group_Refinement(packed, MaxCorrs, cutoff, anfang, ende, mincov, maxclique, greedy):
    """RepeatResolver.c:1634-1693 with the reference's argument order.  The reference fills its global arrays indexed by
    group; here they come back for the groups above the cutoff only (see Packed.group_refinement), and MaxCorrs - which
    the reference modifies in place - as the "MaxCorrs" entry of the result."""
    return packed.group_refinement(MaxCorrs, cutoff, mincov, maxclique, greedy, anfang, ende)


def Parallel_Group_Refinement(packed, MaxCorrs, cutoff, anfang, ende, mincov, maxclique, greedy, NTHREADS=1):
    """RepeatResolver.c:1770-1821: the same work spread over threads - and, because the thread arguments travel in an int
    array (1793-1794), with cutoff and greedy truncated towards zero.  NTHREADS is accepted and ignored: all groups go to the
    device in one batch."""
    return packed.group_refinement(MaxCorrs, float(int(cutoff)), mincov, maxclique, float(int(greedy)), anfang, ende)


def GroupPrecision(bitset, signumber):
    """RepeatResolver.c:1098-1131, the two numbers of its "Group Precision maj / min" printout: over the blocks of 30
    consecutive reads (reads of one copy in the simulator's order), members and non-members of the majority side / of the
    minority side."""
    bits = np.unpackbits(np.ascontiguousarray(bitset, dtype="<u8").view(np.uint8), bitorder="little")[:signumber]
    n = signumber // 30
    drin = bits[:n * 30].reshape(n, 30).sum(axis=1).astype(np.int64)
    drau = 30 - drin
    return int(np.where(drin > drau, drin, drau).sum()), int(np.where(drin > drau, drau, drin).sum())


def dropoff_cutoff_host(sizes, signumber, c=0):
    """Dropoff_Cutoff's rule (1488-1509) on given member counts (host half of rr_group_refinement): (cutoff, Drop_Off)"""
    s = np.ascontiguousarray(sizes, dtype=np.uint32)
    d = C.c_double(0.0)
    cut = lib.rr_dropoff_cutoff_host(s.ctypes.data, len(s), int(signumber), int(c), C.byref(d))
    return cut, d.value


def Relative_Vars(msa, Unterteilung, u_no, MaxCorrs, cutoff, mingroup, device=0):
    """RepeatResolver.c:2424-2493 with the reference's argument order (first version, rr_relative_vars): the groups that
    vary inside part u_no of the read partition, ascending ids (the reference's array without its -1 terminator)."""
    u = np.ascontiguousarray(Unterteilung, dtype=np.int32)
    M = np.ascontiguousarray(MaxCorrs, dtype=np.float64)
    assert len(u) == msa.rows and len(M) == 5 * msa.cols
    out = np.zeros(5 * msa.cols + 1, dtype=np.int32)
    n, pairs = C.c_int(0), C.c_int64(0)
    _check(lib.rr_relative_vars(msa._h, device, u.ctypes.data, int(u_no), M.ctypes.data, float(cutoff), int(mingroup),
                                out.ctypes.data, C.byref(n), C.byref(pairs)), "rr_relative_vars")
    assert out[n.value] == -1
    return out[:n.value].copy()


def relative_score_host(s, gr1, gr2, cov):
    """Relative_Group_Significance on counts (RepeatResolver.c:506-523), host libm"""
    return lib.rr_relative_score_host(s, gr1, gr2, cov)


def relative_vars_from_counts(maxcorrs, gsize_u, cov_u, cutoff, mingroup, triple_counts=None):
    """the host half of rr_relative_vars (test hook).  triple_counts: callable(sel) -> [n_sel][n_sel] int32 of
    |G_sel[a] & G_sel[b] & U|; without it only the selection is returned."""
    M = np.ascontiguousarray(maxcorrs, dtype=np.float64)
    gu = np.ascontiguousarray(gsize_u, dtype=np.int32)
    sel = np.zeros(len(M), dtype=np.int32)
    n_sel = C.c_int(0)
    _check(lib.rr_relative_vars_from_counts(len(M), M.ctypes.data, gu.ctypes.data, int(cov_u), float(cutoff), int(mingroup), None,
                                            sel.ctypes.data, C.byref(n_sel), None, None), "rr_relative_vars_from_counts")
    sel = sel[:n_sel.value].copy()
    if triple_counts is None:
        return sel
    S = np.ascontiguousarray(triple_counts(sel), dtype=np.int32)
    assert S.shape == (len(sel), len(sel))
    out = np.zeros(len(M) + 1, dtype=np.int32)
    n = C.c_int(0)
    _check(lib.rr_relative_vars_from_counts(len(M), M.ctypes.data, gu.ctypes.data, int(cov_u), float(cutoff), int(mingroup),
                                            S.ctypes.data, None, C.byref(n_sel), out.ctypes.data, C.byref(n)),
           "rr_relative_vars_from_counts")
    return out[:n.value].copy()


def Kmeans(msa, Unterteilung, u_no, Vars, mingroup, device=0):
    """RepeatResolver.c:2604-2821 with the reference's argument order (rr_kmeans): splits part u_no of the
    read partition; returns (number of non-empty clusters, the new partition) - the reference updates Unterteilung in place"""
    u = np.array(Unterteilung, dtype=np.int32)
    v = np.ascontiguousarray(Vars, dtype=np.int32)
    assert len(u) == msa.rows
    n = C.c_int(0)
    _check(lib.rr_kmeans(msa._h, device, u.ctypes.data, int(u_no), v.ctypes.data, len(v), int(mingroup), C.byref(n)), "rr_kmeans")
    return n.value, u


def Unterteilungskomprimierung(Unterteilung):
    """RepeatResolver.c:1823-1843: the part numbers renamed 0, 1, 2, ... in the order of their first read; negative entries
    (reads left out) stay.  Returns (number of parts, the renamed partition) - the reference renames in place."""
    u = np.array(Unterteilung, dtype=np.int32)
    keep = u > -1
    vals, first = np.unique(u[keep], return_index=True)
    order = np.argsort(first, kind="stable")                                 # parts by first appearance
    rename = np.zeros(int(vals.max()) + 1 if len(vals) else 1, dtype=np.int32)
    rename[vals[order]] = np.arange(len(vals), dtype=np.int32)
    u[keep] = rename[u[keep]]
    return len(vals), u


def UnterteilungsKomplettierung(Unterteilung, Ausgelassen):
    """RepeatResolver.c:1845-1865: the partition of the reads that were read (Ausgelassen[i] == 1) spread back over all reads
    of the file; the reads left out by the window rule (Ausgelassen[i] == -1) get -1."""
    a = np.asarray(Ausgelassen)
    u = np.asarray(Unterteilung, dtype=np.int32)
    if int(np.count_nonzero(a == 1)) != len(u):
        raise ValueError("one part number per read with Ausgelassen == 1")
    out = np.zeros(len(a), dtype=np.int32)                                   # other marks: the reference leaves the slot unset
    out[a == 1] = u
    out[a == -1] = -1
    return out


def Unterteilung_Rausschreiben(Unterteilung, outputfile):
    """RepeatResolver.c:568-585: one part number per line, no newline after the last"""
    with open(outputfile, "w") as f:
        f.write("\n".join("%d" % int(x) for x in Unterteilung))


def UnterteilungEinlesen(inputfile):
    """RepeatResolver.c:587-607: one number per line as sscanf("%d") reads it (lines of up to 99 characters; a line without a
    number keeps 0 here where the reference leaves the slot unset); None if the file cannot be opened"""
    import re
    try:
        with open(inputfile, "rb") as f:
            data = f.read()
    except OSError:
        return None
    out = []
    for line in data.split(b"\n") if data else []:
        pieces = [line[k:k + 99] for k in range(0, len(line), 99)] or [line]  # fgets(buffer, 100, ...)
        for piece in pieces:
            m = re.match(rb"\s*([+-]?\d+)", piece)
            out.append(int(m.group(1)) if m else 0)
    if data.endswith(b"\n"):
        out.pop()                                                            # no line after the last newline
    return np.array(out, dtype=np.int32)


def Kmeans_Subdivision(msa, Unterteilung, MaxCorrs, cutoff, mingroup, device=0, relative_vars=None, kmeans=None):
    """RepeatResolver.c:3382-3404 (called by main at 4065), the caller of Relative_Vars and Kmeans: the partition is renamed
    (Unterteilungskomprimierung), every part k that exists at that moment and holds more than 2 * mingroup reads is split by
    Kmeans on the groups Relative_Vars selects for it - parts are visited in order and the clusters a split creates get
    numbers above all existing ones, so they are not visited again -, and the result is renamed once more.  Returns (number
    of parts, the new partition); the reference updates Unterteilung in place.  relative_vars / kmeans default to the
    device paths (Relative_Vars, Kmeans of this module); the CPU tests pass the oracle's with the same signatures."""
    relative_vars = relative_vars or (lambda m, u, k, M, c, g: Relative_Vars(m, u, k, M, c, g, device))
    kmeans = kmeans or (lambda m, u, k, v, g: Kmeans(m, u, k, v, g, device))
    number, u = Unterteilungskomprimierung(Unterteilung)
    M = np.ascontiguousarray(MaxCorrs, dtype=np.float64)
    for k in range(number):
        if int(np.count_nonzero(u == k)) > mingroup * 2:                     # 3391
            Vars = relative_vars(msa, u, k, M, cutoff, mingroup)
            _, u = kmeans(msa, u, k, Vars, mingroup)
            u = np.asarray(u, dtype=np.int32)
    return Unterteilungskomprimierung(u)


def kmeans_signatures(msa, Unterteilung, u_no, Vars):
    """host piece of rr_kmeans (test hook): (reads of the part, signatures [anzahl][n_vars/64+1] uint64)"""
    u = np.ascontiguousarray(Unterteilung, dtype=np.int32)
    v = np.ascontiguousarray(Vars, dtype=np.int32)
    reads = np.zeros(max(msa.rows, 1), dtype=np.int32)
    n = C.c_int(0)
    _check(lib.rr_kmeans_signatures(msa._h, u.ctypes.data, int(u_no), v.ctypes.data, len(v), reads.ctypes.data, C.byref(n), None),
           "rr_kmeans_signatures")
    sig = np.zeros((n.value, len(v) // 64 + 1), dtype=np.uint64)
    _check(lib.rr_kmeans_signatures(msa._h, u.ctypes.data, int(u_no), v.ctypes.data, len(v), reads.ctypes.data, C.byref(n),
                                    sig.ctypes.data), "rr_kmeans_signatures")
    return reads[:n.value].copy(), sig


def kmeans_top5_host(sig, i):
    sig = np.ascontiguousarray(sig, dtype=np.uint64)
    out = np.zeros(5, dtype=np.int32)
    _check(lib.rr_kmeans_top5_host(sig.shape[0], sig.shape[1], sig.ctypes.data, int(i), out.ctypes.data), "rr_kmeans_top5_host")
    return out


def kmeans_majority5_host(a, b, c, d, e):
    return int(lib.rr_kmeans_majority5_host(int(a), int(b), int(c), int(d), int(e)))


def kmeans_finish(sig, cen, cluster, mingroup):
    """host piece of rr_kmeans (test hook): the dissolution of small clusters, (final clusters, number of non-empty ones)"""
    sig = np.ascontiguousarray(sig, dtype=np.uint64)
    cen = np.ascontiguousarray(cen, dtype=np.uint64)
    cl = np.ascontiguousarray(cluster, dtype=np.int32)
    out = np.zeros(len(cl), dtype=np.int32)
    n = C.c_int(0)
    _check(lib.rr_kmeans_finish(sig.shape[0], sig.shape[1], sig.ctypes.data, cen.ctypes.data, cl.ctypes.data, int(mingroup),
                                out.ctypes.data, C.byref(n)), "rr_kmeans_finish")
    return out, n.value


def launch_count():
    return int(lib.rr_launch_count())


def lnfact_table(n):
    t = np.zeros(n, dtype=np.float64)
    lib.rr_lnfact_table(t.ctypes.data, n)
    return t


def score_host(s, gr1, gr2, cov, sizei, sizej):
    return lib.rr_score_host(s, gr1, gr2, cov, sizei, sizej)


def score_bound_host(s, gr1, gr2, cov):
    return lib.rr_score_bound_host(s, gr1, gr2, cov)


def below_median_host(s, gr1, gr2, cov):
    return bool(lib.rr_below_median_host(s, gr1, gr2, cov))


def group_score_host(s, gr1, gr2, cov, sizei, sizej):
    """Group_PositiveSignificance (RepeatResolver.c:472-488) on counts, host libm"""
    return lib.rr_group_score_host(s, gr1, gr2, cov, sizei, sizej)


def cliquer_from_counts(query_group, groups, counts, sizes, size_query, mincov=30, maxclique=30, greedy=3.0):
    """the host half of Cliquer on given counts (test hook): (members incl. the query, scores with scores[0] = 100)"""
    groups = np.ascontiguousarray(groups, dtype=np.int32)
    counts = np.ascontiguousarray(counts, dtype=np.int32)
    sizes = np.ascontiguousarray(sizes, dtype=np.int32)
    assert counts.shape == (len(groups), 4) and sizes.shape == groups.shape
    members = np.full(maxclique + 1, -1, dtype=np.int32)
    scores = np.zeros(maxclique, dtype=np.float64)
    n = C.c_int(0)
    _check(lib.rr_cliquer_from_counts(int(query_group), len(groups), groups.ctypes.data, counts.ctypes.data, sizes.ctypes.data,
                                      int(size_query), mincov, maxclique, greedy, members.ctypes.data, scores.ctypes.data,
                                      C.byref(n)), "rr_cliquer_from_counts")
    return members[:n.value].copy(), scores[:n.value].copy()


HIT_DTYPE = np.dtype([("slot", "<i4"), ("group", "<i4"), ("s", "<i4"), ("gr1", "<i4"), ("gr2", "<i4"), ("cov", "<i4"),
                      ("z", "<f8")])


def cliquer_from_hits(query_groups, hits, gsize, mincov=30, maxclique=30, greedy=3.0):
    """the host half of rr_cliquer_batch on a given hit list (test hook; hits: array of HIT_DTYPE)"""
    q = np.ascontiguousarray(query_groups, dtype=np.int32)
    hits = np.ascontiguousarray(hits, dtype=HIT_DTYPE)
    gsize = np.ascontiguousarray(gsize, dtype=np.int32)
    members = np.full((len(q), maxclique + 1), -1, dtype=np.int32)
    scores = np.zeros((len(q), maxclique), dtype=np.float64)
    n = np.zeros(len(q), dtype=np.int32)
    _check(lib.rr_cliquer_from_hits(len(q), q.ctypes.data, len(hits), hits.ctypes.data, gsize.ctypes.data, len(gsize), mincov,
                                    maxclique, greedy, members.ctypes.data, scores.ctypes.data, n.ctypes.data),
           "rr_cliquer_from_hits")
    return members, scores, n


def length_classes(rows):
    """rank boundaries [n_classes + 1] of the length classes rr_pack sorts `rows` rows into (rr_length_classes)"""
    cs = np.zeros(9, dtype=np.int32)
    n = lib.rr_length_classes(int(rows), cs.ctypes.data)
    if n < 1:
        raise RRError(n, "rr_length_classes")
    return cs[:n + 1].copy()


def rank_rows(start, end):
    """the row order of rr_pack for spans start[r]..end[r]: (length class, span start, span end), classes by span length
    (ties: start, then original order); returns (perm [rank -> row], class of every rank, class_start)"""
    start = np.asarray(start, dtype=np.int64)
    end = np.asarray(end, dtype=np.int64)
    R = len(start)
    cs = length_classes(R)
    by_len = np.lexsort((np.arange(R), start, end - start))
    cls = np.zeros(R, dtype=np.int64)
    for c in range(len(cs) - 1):
        cls[by_len[cs[c]:cs[c + 1]]] = c
    perm = np.lexsort((np.arange(R), end, start, cls))
    return perm, cls[perm], cs


def contraction_ranges(start, end, cols, class_start, ti, tj, kunit):
    """k_lo [n_classes][ceil(cols/tj)], k_hi [n_classes][n_rowblocks] of the scan plan for rows given in rank order (test hook)"""
    start = np.ascontiguousarray(start, dtype=np.int32)
    end = np.ascontiguousarray(end, dtype=np.int32)
    cs = np.ascontiguousarray(class_start, dtype=np.int32)
    ncls = len(cs) - 1
    ncb = max((cols + tj - 1) // tj, 1)
    nrb_max = max((cols + ti - 1) // ti, 1)
    k_lo = np.zeros(ncls * ncb, dtype=np.int32)
    k_hi = np.zeros(ncls * nrb_max, dtype=np.int32)
    nrb = C.c_int(0)
    _check(lib.rr_contraction_ranges(start.ctypes.data, end.ctypes.data, len(start), cols, ncls, cs.ctypes.data, ti, tj, kunit,
                                     k_lo.ctypes.data, k_hi.ctypes.data, C.byref(nrb)), "rr_contraction_ranges")
    n = max(nrb.value, 1)
    return k_lo.reshape(ncls, ncb), k_hi[:ncls * n].reshape(ncls, n), nrb.value


def breakcols_from_spans(start, end, cols, mincov):
    start = np.ascontiguousarray(start, dtype=np.int32)
    end = np.ascontiguousarray(end, dtype=np.int32)
    out = np.zeros(cols, dtype=np.int32)
    _check(lib.rr_breakcols_from_spans(start.ctypes.data, end.ctypes.data, len(start), cols, mincov, out.ctypes.data),
           "rr_breakcols_from_spans")
    return out
