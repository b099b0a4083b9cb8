"""Test and measurement hooks of librr_maxcorr.so (include/rr_debug.h).  Not part of the drop-in interface."""
import ctypes as C

import numpy as np

from ._lib import lib, ScanOpts
from .maxcorr import VARIANTS, _check

DEBUG_MMA_ONLY = 0x400      # scan flag: producer + MMA pipeline only (timing decomposition; leaves no result)


def set_cliquer_cap(cap):
    """capacity (records) of the candidate / hit lists of rr_cliquer_batch; 0 restores the default"""
    lib.rr_debug_set_cliquer_cap(int(cap))


def set_deferred_cap(cap):
    """capacity (candidates) of the deferred-evaluation list of the tcgen05 scan; 0 restores the default"""
    lib.rr_debug_set_deferred_cap(int(cap))


def umma_tiles(packed, mincov=30, variant="umma_mxf4", part_index=0, part_count=1):
    """(row tiles, column tiles) of the plan of the last scan with these options"""
    opts = ScanOpts(mincov, VARIANTS[variant], 0, part_index, part_count)
    nr, nc = C.c_int(0), C.c_int(0)
    _check(lib.rr_debug_umma_counts(packed._h, C.byref(opts), 0, 0, None, None, None, C.byref(nr), C.byref(nc)), "rr_debug_umma_counts")
    return nr.value, nc.value


def umma_counts(packed, row_tile, col_tile, mincov=30, variant="umma_mxf4", part_index=0, part_count=1):
    """The accumulator of one (row tile, column tile) pair of the tcgen05 scan kernel as its epilogue reads it from
    TMEM: (counts [128][240], row group ids [128] (-1 = padding row), column group ids [240] (-1 = beyond the MSA)).
    Call after packed.scan(...) with the same options."""
    opts = ScanOpts(mincov, VARIANTS[variant], 0, part_index, part_count)
    counts = np.zeros((128, 240), dtype=np.int32)
    rg = np.zeros(128, dtype=np.int32)
    cg = np.zeros(240, dtype=np.int32)
    _check(lib.rr_debug_umma_counts(packed._h, C.byref(opts), int(row_tile), int(col_tile), counts.ctypes.data, rg.ctypes.data,
                                    cg.ctypes.data, None, None), "rr_debug_umma_counts")
    return counts, rg, cg


def mma_peak(variant="umma_mxf4", device=0, kblocks_per_sm=16384, reps=5):
    """The tensor pipe's rate for the MMA kind / tile shape / operand layout of the scan (bare tcgen05.mma loop on every
    SM, csrc/rr_scan_umma.cu): {"tflops": 2 * MACs / s / 1e12, "ms": best launch, "macs": per launch}"""
    ms, macs = C.c_float(0), C.c_double(0)
    _check(lib.rr_debug_mma_peak(device, VARIANTS[variant], kblocks_per_sm, reps, C.byref(ms), C.byref(macs)), "rr_debug_mma_peak")
    return {"tflops": 2.0 * macs.value / (ms.value * 1e-3) / 1e12, "ms": ms.value, "macs": macs.value}


def mma_peak_shape(variant="umma_mxf4", cta_group=2, n_cols=240, device=0, kblocks_per_sm=16384, reps=5):
    """mma_peak for another issue form: cta_group 2 = a CTA pair issuing M = 256 x n_cols (each CTA holds half of B)"""
    ms, macs = C.c_float(0), C.c_double(0)
    _check(lib.rr_debug_mma_peak_shape(device, VARIANTS[variant], cta_group, n_cols, kblocks_per_sm, reps, C.byref(ms), C.byref(macs)),
           "rr_debug_mma_peak_shape")
    return {"tflops": 2.0 * macs.value / (ms.value * 1e-3) / 1e12, "ms": ms.value, "macs": macs.value}


def kmeans_finish_table(sig, cen, cluster, mingroup):
    """rr_kmeans_finish with the scores looked up in the table rr_kmeans makes on the device (filled on the host here):
    (final clusters, number of non-empty ones)"""
    sig = np.ascontiguousarray(sig, dtype=np.uint64)
    cen = np.ascontiguousarray(cen, dtype=np.uint64)
    cl = np.ascontiguousarray(cluster, dtype=np.int32)
    out = np.zeros(len(cl), dtype=np.int32)
    n = C.c_int(0)
    _check(lib.rr_debug_kmeans_finish_table(sig.shape[0], sig.shape[1], sig.ctypes.data, cen.ctypes.data, cl.ctypes.data, int(mingroup),
                                            out.ctypes.data, C.byref(n)), "rr_debug_kmeans_finish_table")
    return out, n.value
