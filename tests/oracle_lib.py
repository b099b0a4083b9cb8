"""ctypes binding of oracle/liboracle.so -- the CPU restatement used as the checker.
Test infrastructure: imported only from tests/, __graft_entry__.smoke() and bench.py."""
import ctypes as C
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_lib = None


def lib():
    global _lib
    if _lib is None:
        path = os.path.join(ROOT, "oracle", "liboracle.so")
        if not os.path.exists(path):
            import subprocess
            subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "liboracle.so"])
        L = C.CDLL(path)
        L.rr_oracle_load.restype = C.c_void_p
        L.rr_oracle_load.argtypes = [C.c_char_p]
        L.rr_oracle_from_codes.restype = C.c_void_p
        L.rr_oracle_from_codes.argtypes = [C.c_void_p, C.c_int, C.c_int]
        L.rr_oracle_free.argtypes = [C.c_void_p]
        L.rr_oracle_R.argtypes = [C.c_void_p]
        L.rr_oracle_N.argtypes = [C.c_void_p]
        L.rr_oracle_gsize.restype = C.POINTER(C.c_int)
        L.rr_oracle_gsize.argtypes = [C.c_void_p]
        L.rr_oracle_coverage.restype = C.POINTER(C.c_int)
        L.rr_oracle_coverage.argtypes = [C.c_void_p]
        L.rr_oracle_counts.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        L.rr_oracle_count_matrix.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        L.rr_oracle_count_matrix.restype = None
        L.rr_oracle_score.restype = C.c_double
        L.rr_oracle_score.argtypes = [C.c_uint, C.c_uint, C.c_uint, C.c_uint, C.c_int, C.c_int]
        L.rr_oracle_scan.restype = C.c_int64
        L.rr_oracle_scan.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        L.rr_oracle_count_pairs.restype = C.c_int64
        L.rr_oracle_count_pairs.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int]
        L.rr_oracle_write.argtypes = [C.c_char_p, C.c_void_p, C.c_int]
        L.rr_oracle_group_score.restype = C.c_double
        L.rr_oracle_group_score.argtypes = [C.c_uint, C.c_uint, C.c_uint, C.c_uint, C.c_int, C.c_int]
        L.rr_oracle_cliquer.restype = C.c_int
        L.rr_oracle_cliquer.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, C.c_int, C.c_void_p, C.c_void_p]
        L.rr_oracle_relative_score.restype = C.c_double
        L.rr_oracle_relative_score.argtypes = [C.c_uint] * 4
        L.rr_oracle_relative_vars.restype = C.c_int
        L.rr_oracle_relative_vars.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_double, C.c_int, C.c_void_p]
        L.rr_oracle_kmeans.restype = C.c_int
        L.rr_oracle_kmeans.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int]
        L.gsl_cdf_hypergeometric_P.restype = C.c_double
        L.gsl_cdf_hypergeometric_P.argtypes = [C.c_uint] * 4
        L.rr_oracle_lnfact.restype = C.c_double
        L.rr_oracle_lnfact.argtypes = [C.c_uint]
        L.gsl_cdf_hypergeometric_Q.restype = C.c_double
        L.gsl_cdf_hypergeometric_Q.argtypes = [C.c_uint] * 4
        _lib = L
    return _lib


class Oracle:
    def __init__(self, handle):
        if not handle:
            raise OSError("MA is missing.")
        self._h = C.c_void_p(handle)
        self.R = lib().rr_oracle_R(self._h)
        self.N = lib().rr_oracle_N(self._h)

    @classmethod
    def load(cls, path):
        return cls(lib().rr_oracle_load(os.fsencode(path)))

    @classmethod
    def from_text(cls, text, tmp_path):
        p = os.path.join(str(tmp_path), "oracle_in.msa")
        with open(p, "wb") as f:
            f.write(text)
        return cls.load(p)

    @classmethod
    def from_codes(cls, codes):
        codes = np.ascontiguousarray(codes, dtype=np.uint8)
        return cls(lib().rr_oracle_from_codes(codes.ctypes.data, codes.shape[0], codes.shape[1]))

    def gsize(self):
        return np.ctypeslib.as_array(lib().rr_oracle_gsize(self._h), shape=(5 * self.N,)).copy()

    def coverage(self):
        return np.ctypeslib.as_array(lib().rr_oracle_coverage(self._h), shape=(self.N,)).copy()

    def counts(self, i, j):
        out = (C.c_int * 4)()
        lib().rr_oracle_counts(self._h, int(i), int(j), out)
        return list(out)

    def count_matrix(self, rows, cols):
        """|G_r & G_c| (Schnitt, MaxCorrelation.c:114-125) for every pair of two lists of group ids: [len(rows)][len(cols)]"""
        r = np.ascontiguousarray(rows, dtype=np.int32)
        c = np.ascontiguousarray(cols, dtype=np.int32)
        out = np.zeros((len(r), len(c)), dtype=np.int32)
        lib().rr_oracle_count_matrix(self._h, len(r), r.ctypes.data, len(c), c.ctypes.data, out.ctypes.data)
        return out

    def scan(self, mincov=30, modulus=None, res_lo=0, res_hi=None, threads=8):
        """full scan when modulus is None (threads pthreads), else rows ii % modulus in [res_lo,res_hi)"""
        if modulus is None:
            modulus, res_lo, res_hi = threads, 0, threads
        G = 5 * self.N
        M = np.zeros(G + 1, dtype=np.float64)
        A = np.zeros(G + 1, dtype=np.int32)
        P = lib().rr_oracle_scan(self._h, mincov, modulus, res_lo, res_hi, M.ctypes.data, A.ctypes.data)
        return M[:G], A[:G], int(P)

    def count_pairs(self, mincov=30, modulus=1, res_lo=0, res_hi=1):
        """PositiveSignificance calls (MaxCorrelation.c:820) of the rows ii % modulus in [res_lo, res_hi), from the loops and
        filters alone (no score is evaluated); one thread per residue"""
        return int(lib().rr_oracle_count_pairs(self._h, mincov, modulus, res_lo, res_hi))

    def cliquer(self, a, mincov=30, maxclique=30, greedy=3.0, anfang=0, ende=None):
        """RepeatResolver.c:1179-1240 for query group a: (members incl. a, scores with best[0] = 100)"""
        clique = np.full(maxclique + 1, -1, dtype=np.int32)
        best = np.zeros(maxclique, dtype=np.float64)
        n = lib().rr_oracle_cliquer(self._h, anfang, self.N if ende is None else ende, mincov, maxclique, greedy, int(a),
                                    clique.ctypes.data, best.ctypes.data)
        return clique[:n].copy(), best[:n].copy()

    def relative_vars(self, unterteilung, u_no, maxcorrs, cutoff, mingroup):
        """RepeatResolver.c:2424-2493: the groups that vary inside part u_no of the read partition (ascending ids)"""
        u = np.ascontiguousarray(unterteilung, dtype=np.int32)
        m = np.ascontiguousarray(maxcorrs, dtype=np.float64)
        assert len(u) == self.R and len(m) == 5 * self.N
        out = np.zeros(5 * self.N + 1, dtype=np.int32)
        n = lib().rr_oracle_relative_vars(self._h, u.ctypes.data, int(u_no), m.ctypes.data, float(cutoff), int(mingroup),
                                          out.ctypes.data)
        assert out[n] == -1
        return out[:n].copy()

    def kmeans(self, unterteilung, u_no, vars_, mingroup):
        """RepeatResolver.c:2604-2821: (number of non-empty clusters, the partition after the split)"""
        u = np.array(unterteilung, dtype=np.int32)
        v = np.ascontiguousarray(vars_, dtype=np.int32)
        assert len(u) == self.R
        n = lib().rr_oracle_kmeans(self._h, u.ctypes.data, int(u_no), v.ctypes.data, len(v), int(mingroup))
        return n, u

    def close(self):
        if self._h:
            lib().rr_oracle_free(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def score(s, gr1, gr2, cov, sizei=0, sizej=0):
    return lib().rr_oracle_score(s, gr1, gr2, cov, sizei, sizej)


def group_score(s, gr1, gr2, cov, sizei=0, sizej=0):
    return lib().rr_oracle_group_score(s, gr1, gr2, cov, sizei, sizej)


def relative_score(s, gr1, gr2, cov):
    """Relative_Group_Significance on counts (RepeatResolver.c:506-523, 490-504)"""
    return lib().rr_oracle_relative_score(s, gr1, gr2, cov)


def hyper_P(k, n1, n2, t):
    return lib().gsl_cdf_hypergeometric_P(k, n1, n2, t)


def hyper_Q(k, n1, n2, t):
    return lib().gsl_cdf_hypergeometric_Q(k, n1, n2, t)


def lnfact(n):
    return lib().rr_oracle_lnfact(n)


def fmt_lines(M):
    """the reference's "%f\\n" text (MaxCorrsRausschreiben, 527-530)"""
    return "".join("%f\n" % v for v in M).encode()


# ---- CliqueGroup / CliqueCoverage (RepeatResolver.c:976-1008, 1064-1096), plain numpy on the code matrix ----
def clique_group(codes, clique, c):
    """the reads contained in more than c of the clique's groups (976-1008): per read the number of members g with
    Signatures[read][g / 5] == g % 5 (GrElement of Groups[g], filled at 405-411); bool [rows]"""
    codes = np.asarray(codes)
    score = np.zeros(codes.shape[0], dtype=np.int64)
    for g in clique:
        if g < 0:
            break                                    # the first negative entry ends the clique (986-993)
        score += codes[:, g // 5] == g % 5
    return score > c


def clique_coverage(codes, clique, c):
    """the reads covered (a symbol, not a blank) at more than c of the sites of the clique's groups (1064-1096)"""
    codes = np.asarray(codes)
    score = np.zeros(codes.shape[0], dtype=np.int64)
    for g in clique:
        if g < 0:
            break
        score += codes[:, g // 5] < 5
    return score > c


def dropoff_cutoff(codes, clique, size, c=0):
    """Dropoff_Cutoff (RepeatResolver.c:1460-1522) on the first `size` members of a clique (Sizes[c_i], 1650): sizes[k] = reads
    contained in more than k of them (1472-1486); the cutoff k in [max(1, c), size - 1) that minimises
    (sizes[k-1] - sizes[k+1]) / min(signumber - sizes[k], sizes[k]) where that denominator is positive, the first such k on
    ties (1494-1509).  Returns (cutoff, Drop_Off)."""
    codes = np.asarray(codes)
    signumber = codes.shape[0]
    score = np.zeros(signumber, dtype=np.int64)
    for g in clique[:size]:
        score += codes[:, g // 5] == g % 5
    sizes = [float(np.count_nonzero(score > k)) for k in range(size)]
    drop_c = max(1, c)
    min_drop = 1000000.0
    for i in range(drop_c, size - 1):
        den = min(float(signumber) - sizes[i], sizes[i])
        if den > 0:
            drop = (sizes[i - 1] - sizes[i + 1]) / den
            if drop < min_drop:
                min_drop, drop_c = drop, i
    return drop_c, min_drop


def group_precision(mask):
    """GroupPrecision (1098-1131): over blocks of 30 consecutive reads, the majority and the minority counts (maj, min)"""
    mask = np.asarray(mask, dtype=bool)
    n = len(mask) // 30
    drin = mask[:n * 30].reshape(n, 30).sum(axis=1)
    drau = 30 - drin
    return int(np.where(drin > drau, drin, drau).sum()), int(np.where(drin > drau, drau, drin).sum())


def group_refinement(oracle, codes, maxcorrs, cutoff, mincov, maxclique, greedy, anfang=0, ende=None):
    """Group_Refinement (1634-1693) on the oracle's Cliquer: for every group i with MaxCorrs[i] > cutoff its clique, Sizes[i]
    (members before the first entry <= 0, 1650) and - where Sizes[i] > 5 - the cutoff of Dropoff_Cutoff(i, 0) (the results of
    BestCutoff and KorrMaxCutoff are overwritten at 1662 and have no side effect), Drop_Off[i], CliqueGroup and CliqueCoverage
    at that cutoff; MaxCorrs[i] = 0 where Sizes[i] <= 5 (1685).  Returns (MaxCorrs after, {i: dict})."""
    M = np.array(maxcorrs, dtype=np.float64)
    res = {}
    for i in np.flatnonzero(M > cutoff):
        members, _ = oracle.cliquer(int(i), mincov, maxclique, greedy, anfang, ende)
        clique = np.full(maxclique + 1, -1, dtype=np.int32)
        clique[:len(members)] = members
        size = 0
        while clique[size] > 0:
            size += 1
        r = {"clique": clique, "size": size, "cutoff": 0, "drop_off": 1000.0, "group": None, "coverage": None}
        if size > 5:
            r["cutoff"], r["drop_off"] = dropoff_cutoff(codes, clique, size, 0)
            r["group"] = clique_group(codes, clique, r["cutoff"])
            r["coverage"] = clique_coverage(codes, clique, r["cutoff"])
        else:
            M[i] = 0.0
        res[int(i)] = r
    return M, res


def bitset_words(mask):
    """a bool [rows] as the reference's group words: read r = bit r % 64 of word r / 64 (GrAdd 211-217), rows / 64 + 1 words"""
    mask = np.asarray(mask, dtype=bool)
    sc = len(mask) // 64 + 1
    bits = np.zeros(sc * 64, dtype=np.uint8)
    bits[:len(mask)] = mask
    return np.packbits(bits, bitorder="little").view("<u8").copy()
