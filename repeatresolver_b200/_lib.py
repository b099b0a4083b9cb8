"""ctypes loader for the in-tree shared libraries.

librr_maxcorr.so is the C ABI declared in include/rr_maxcorr.h (CUDA, sm_100a).  It is
loaded eagerly and loading fails loudly: there is no Python or CPU implementation of the
scan behind this package.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))


class ScanOpts(C.Structure):
    _fields_ = [("mincov", C.c_int), ("variant", C.c_int), ("flags", C.c_uint),
                ("part_index", C.c_int), ("part_count", C.c_int)]


class ScanStats(C.Structure):
    _fields_ = [("pair_tests", C.c_int64), ("exact_evals", C.c_int64), ("bound_evals", C.c_int64),
                ("work_units", C.c_int64), ("executed_ops", C.c_int64), ("variant", C.c_int),
                ("rows", C.c_int), ("cols", C.c_int), ("row_sites", C.c_int), ("general_break", C.c_int),
                ("h2d_ms", C.c_float), ("pack_ms", C.c_float), ("prepare_ms", C.c_float),
                ("kernel_ms", C.c_float), ("fetch_ms", C.c_float), ("finalize_ms", C.c_float)]

    def as_dict(self):
        return {name: getattr(self, name) for name, _ in self._fields_}


class CliquerStats(C.Structure):
    _fields_ = [("pairs", C.c_int64), ("candidates", C.c_int64), ("hits", C.c_int64), ("host_evals", C.c_int64),
                ("kernel_ms", C.c_float), ("launches", C.c_int), ("retries", C.c_int)]

    def as_dict(self):
        return {name: getattr(self, name) for name, _ in self._fields_}


class MsagenParams(C.Structure):
    _fields_ = [("type", C.c_int), ("copies", C.c_int), ("coverage", C.c_int), ("repeat_len", C.c_int),
                ("diff", C.c_double), ("seed", C.c_uint64), ("flank", C.c_int), ("min_overlap", C.c_int),
                ("max_reads", C.c_int), ("threads", C.c_int)]


def _load(name):
    path = os.path.join(_HERE, name)
    if not os.path.exists(path):
        raise ImportError(
            f"{path} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            f"or `make -C repeatresolver_b200/csrc` (nvcc, sm_100a). There is no fallback path.")
    return C.CDLL(path, mode=C.RTLD_GLOBAL)


lib = _load("librr_maxcorr.so")
gen = _load("librr_msagen.so")

_vp, _i, _i64, _u32, _d = C.c_void_p, C.c_int, C.c_int64, C.c_uint32, C.c_double
_P = C.POINTER


def _sig(fn, res, args):
    fn.restype = res
    fn.argtypes = args


# every symbol include/rr_maxcorr.h declares
ABI_SYMBOLS = [
    "rr_msa_read", "rr_msa_read_window", "rr_msa_from_text", "rr_msa_from_cells", "rr_msa_alloc", "rr_msa_rows", "rr_msa_cols",
    "rr_msa_cells", "rr_msa_free", "rr_device_count", "rr_variant_available", "rr_pack", "rr_pack_rows", "rr_pack_slice_spans", "rr_pack_set_spans", "rr_pack_bits_device", "rr_pack_finish", "rr_packed_free", "rr_scan", "rr_scan_fetch", "rr_scan_finalize", "rr_scan_set_thresholds", "rr_scan_values_device", "rr_scan_set_thresholds_device",
    "rr_pair_counts", "rr_packed_sizes", "rr_maxcorr_run", "rr_maxcorr_write", "rr_argmax_write", "rr_maxcorr_write_bin", "rr_maxcorr_read_bin", "rr_maxcorr_read_text", "rr_lnfact",
    "rr_lnfact_table", "rr_score_host", "rr_score_bound_host", "rr_below_median_host", "rr_breakcols_from_spans", "rr_contraction_ranges", "rr_length_classes", "rr_cliquer", "rr_cliquer_batch", "rr_clique_groups", "rr_group_refinement", "rr_dropoff_cutoff_host", "rr_cliquer_from_counts", "rr_cliquer_from_hits", "rr_relative_vars", "rr_relative_vars_packed", "rr_relative_vars_from_counts", "rr_relative_score_host", "rr_kmeans", "rr_kmeans_signatures", "rr_kmeans_finish", "rr_kmeans_top5_host", "rr_kmeans_majority5_host", "rr_group_score_host",
    "rr_timer_start", "rr_timer_stop", "rr_launch_count", "rr_last_error", "rr_version",
]
MSAGEN_SYMBOLS = ["rr_msagen_create", "rr_msagen_free", "rr_msagen_rows", "rr_msagen_cols", "rr_msagen_read_copy",
                  "rr_msagen_fill_codes", "rr_msagen_fill_text", "rr_msagen_write"]

_sig(lib.rr_msa_read, _i, [C.c_char_p, _P(_vp)])
_sig(lib.rr_msa_read_window, _i, [C.c_char_p, _i, _i, _P(_vp), _vp, _i64, _P(_i64)])
_sig(lib.rr_msa_from_text, _i, [C.c_char_p, C.c_size_t, _P(_vp)])
_sig(lib.rr_msa_from_cells, _i, [_vp, _i, _i, _i, _P(_vp)])
_sig(lib.rr_msa_alloc, _i, [_i, _i, _i, _P(_vp)])
_sig(lib.rr_msa_rows, _i, [_vp])
_sig(lib.rr_msa_cols, _i, [_vp])
_sig(lib.rr_msa_cells, _vp, [_vp])
_sig(lib.rr_msa_free, None, [_vp])
_sig(lib.rr_device_count, _i, [])
_sig(lib.rr_variant_available, _i, [_i])
_sig(lib.rr_pack, _i, [_vp, _i, _P(_vp)])
_sig(lib.rr_pack_rows, _i, [_vp, _i, _i, _i, _P(_vp)])
_sig(lib.rr_pack_slice_spans, _i, [_vp, _vp, _vp, _vp])
_sig(lib.rr_pack_set_spans, _i, [_vp, _vp, _vp, _vp])
_sig(lib.rr_pack_bits_device, _i, [_vp, _P(_vp), _P(C.c_size_t)])
_sig(lib.rr_pack_finish, _i, [_vp])
_sig(lib.rr_packed_free, None, [_vp])
_sig(lib.rr_scan, _i, [_vp, _P(ScanOpts), _P(ScanStats)])
_sig(lib.rr_scan_fetch, _i, [_vp, _vp, _vp])
_sig(lib.rr_scan_finalize, _i, [_vp, _vp, _vp])
_sig(lib.rr_scan_set_thresholds, _i, [_vp, _vp])
_sig(lib.rr_scan_values_device, _i, [_vp, _vp])
_sig(lib.rr_scan_set_thresholds_device, _i, [_vp, _vp])
_sig(lib.rr_pair_counts, _i, [_vp, _i64, _vp, _vp, _vp])
_sig(lib.rr_packed_sizes, _i, [_vp, _vp, _vp])
_sig(lib.rr_maxcorr_run, _i, [_vp, _i, _i, _i, C.c_uint, _vp, _vp, _P(ScanStats)])
_sig(lib.rr_maxcorr_write, _i, [C.c_char_p, _vp, _i64])
_sig(lib.rr_argmax_write, _i, [C.c_char_p, _vp, _i64])
_sig(lib.rr_maxcorr_write_bin, _i, [C.c_char_p, _vp, _vp, _i64])
_sig(lib.rr_maxcorr_read_bin, _i, [C.c_char_p, _i, _i, _i, _vp, _vp, _P(_i64)])
_sig(lib.rr_maxcorr_read_text, _i, [C.c_char_p, _i, _i, _vp, _i64, _P(_i64)])
_sig(lib.rr_lnfact, _d, [C.c_uint])
_sig(lib.rr_lnfact_table, None, [_vp, C.c_size_t])
_sig(lib.rr_score_host, _d, [_u32, _u32, _u32, _u32, C.c_int32, C.c_int32])
_sig(lib.rr_score_bound_host, _d, [_u32, _u32, _u32, _u32])
_sig(lib.rr_below_median_host, _i, [_u32, _u32, _u32, _u32])
_sig(lib.rr_breakcols_from_spans, _i, [_vp, _vp, _i, _i, _i, _vp])
_sig(lib.rr_cliquer, _i, [_vp, _i, _i, _i, _i, _i, _d, _vp, _vp, _P(_i)])
_sig(lib.rr_cliquer_batch, _i, [_vp, _i64, _vp, _i, _i, _i, _i, _d, _vp, _vp, _vp, _P(CliquerStats)])
_sig(lib.rr_relative_vars, _i, [_vp, _i, _vp, _i, _vp, _d, _i, _vp, _P(_i), _P(_i64)])
_sig(lib.rr_relative_vars_packed, _i, [_vp, _vp, _i, _vp, _d, _i, _vp, _P(_i), _P(_i64)])
_sig(lib.rr_relative_vars_from_counts, _i, [_i64, _vp, _vp, _i, _d, _i, _vp, _vp, _P(_i), _vp, _P(_i)])
_sig(lib.rr_relative_score_host, _d, [_u32, _u32, _u32, _u32])
_sig(lib.rr_kmeans, _i, [_vp, _i, _vp, _i, _vp, _i, _i, _P(_i)])
_sig(lib.rr_kmeans_signatures, _i, [_vp, _vp, _i, _vp, _i, _vp, _P(_i), _vp])
_sig(lib.rr_kmeans_finish, _i, [_i, _i, _vp, _vp, _vp, _i, _vp, _P(_i)])
_sig(lib.rr_kmeans_top5_host, _i, [_i, _i, _vp, _i, _vp])
_sig(lib.rr_kmeans_majority5_host, C.c_uint64, [C.c_uint64] * 5)
_sig(lib.rr_cliquer_from_hits, _i, [_i64, _vp, _i64, _vp, _vp, _i64, _i, _i, _d, _vp, _vp, _vp])
_sig(lib.rr_cliquer_from_counts, _i, [_i, _i64, _vp, _vp, _vp, _i, _i, _i, _d, _vp, _vp, _P(_i)])
_sig(lib.rr_clique_groups, _i, [_vp, _i64, _vp, _i, _vp, _vp, _vp, _vp])
_sig(lib.rr_group_refinement, _i, [_vp, _vp, _d, _i, _i, _i, _i, _d, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _P(_i64), _P(CliquerStats)])
_sig(lib.rr_dropoff_cutoff_host, _i, [_vp, _i, _i, _i, _P(_d)])
_sig(lib.rr_group_score_host, _d, [_u32, _u32, _u32, _u32, C.c_int32, C.c_int32])
_sig(lib.rr_contraction_ranges, _i, [_vp, _vp, _i, _i, _i, _vp, _i, _i, _i, _vp, _vp, _P(_i)])
_sig(lib.rr_length_classes, _i, [_i, _vp])
_sig(lib.rr_timer_start, _i, [_vp])
_sig(lib.rr_timer_stop, _i, [_vp, _P(C.c_float)])
_sig(lib.rr_launch_count, _i64, [])
_sig(lib.rr_last_error, C.c_char_p, [])
_sig(lib.rr_version, C.c_char_p, [])

# include/rr_debug.h: test / measurement hooks, not part of the drop-in boundary
DEBUG_SYMBOLS = ["rr_debug_set_cliquer_cap", "rr_debug_set_deferred_cap", "rr_debug_umma_counts", "rr_debug_mma_peak", "rr_debug_mma_peak_shape", "rr_debug_kmeans_finish_table"]
_sig(lib.rr_debug_set_cliquer_cap, None, [C.c_ulonglong])
_sig(lib.rr_debug_set_deferred_cap, None, [C.c_ulonglong])
_sig(lib.rr_debug_mma_peak, _i, [_i, _i, _i, _i, _P(C.c_float), _P(C.c_double)])
_sig(lib.rr_debug_mma_peak_shape, _i, [_i, _i, _i, _i, _i, _i, _P(C.c_float), _P(C.c_double)])
_sig(lib.rr_debug_kmeans_finish_table, _i, [_i, _i, _vp, _vp, _vp, _i, _vp, _P(_i)])
_sig(lib.rr_debug_umma_counts, _i, [_vp, _P(ScanOpts), _i, _i, _vp, _vp, _vp, _P(_i), _P(_i)])

_sig(gen.rr_msagen_create, _vp, [_P(MsagenParams)])
_sig(gen.rr_msagen_free, None, [_vp])
_sig(gen.rr_msagen_rows, _i, [_vp])
_sig(gen.rr_msagen_cols, _i, [_vp])
_sig(gen.rr_msagen_read_copy, _P(C.c_int), [_vp])
_sig(gen.rr_msagen_fill_codes, None, [_vp, _vp])
_sig(gen.rr_msagen_fill_text, None, [_vp, _vp])
_sig(gen.rr_msagen_write, _i, [_vp, C.c_char_p])
