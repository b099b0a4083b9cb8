"""repeatresolver_b200 -- B200-native MaxCorrelation scan (drop-in for
PhilippBongartz/RepeatResolver's MaxCorrelation.c hot path).

The package is a thin host-side mirror of the reference interface over the C ABI in
include/rr_maxcorr.h (librr_maxcorr.so: hand-written sm_100a CUDA).  Importing it loads
the shared library and fails if it has not been built.
"""
from .maxcorr import (MSA, Packed, RRError, Einlesen, Parallel_AllMaxCorrsRechner, MaxCorrsRausschreiben, MaxCorrsRausschreiben_bin, MaxCorrsEinlesen_bin, MaxCorrsEinlesen,
                      MaxCorrelation, Cliquer, CliqueGroup, CliqueCoverage, group_reads, Group_Refinement_Cliques, coverage_restriction, Group_Refinement, Parallel_Group_Refinement, GroupPrecision, dropoff_cutoff_host, Relative_Vars, Kmeans, Kmeans_Subdivision, Unterteilungskomprimierung, UnterteilungsKomplettierung, Unterteilung_Rausschreiben, UnterteilungEinlesen, kmeans_signatures, kmeans_top5_host, kmeans_majority5_host, kmeans_finish,
                      relative_score_host,
                      relative_vars_from_counts, device_count, variant_available, launch_count, lnfact_table, score_host, score_bound_host,
                      below_median_host, breakcols_from_spans, contraction_ranges, length_classes, rank_rows, cliquer_from_counts, cliquer_from_hits, HIT_DTYPE, group_score_host, VARIANTS, VARIANT_NAMES,
                      FLAG_NO_PRUNE, FLAG_HOST_FINALIZE, FLAG_GENERAL_BREAK, FLAG_SEED_ONLY, FLAG_SKIP_SEED)
from .msagen import MsaGen
from . import debug

__all__ = ["MSA", "Packed", "RRError", "Einlesen", "Parallel_AllMaxCorrsRechner", "MaxCorrsRausschreiben", "MaxCorrsRausschreiben_bin", "MaxCorrsEinlesen_bin", "MaxCorrsEinlesen",
           "MaxCorrelation", "Cliquer", "CliqueGroup", "CliqueCoverage", "group_reads", "Group_Refinement_Cliques", "coverage_restriction", "Group_Refinement", "Parallel_Group_Refinement", "GroupPrecision", "dropoff_cutoff_host", "Relative_Vars", "Kmeans", "Kmeans_Subdivision", "Unterteilungskomprimierung", "UnterteilungsKomplettierung", "Unterteilung_Rausschreiben", "UnterteilungEinlesen", "kmeans_signatures", "kmeans_top5_host", "kmeans_majority5_host", "kmeans_finish",
           "relative_score_host",
           "relative_vars_from_counts", "device_count", "variant_available", "launch_count", "lnfact_table", "score_host", "score_bound_host",
           "below_median_host", "breakcols_from_spans", "contraction_ranges", "length_classes", "rank_rows", "cliquer_from_counts", "cliquer_from_hits", "HIT_DTYPE", "group_score_host", "MsaGen", "debug", "VARIANTS", "VARIANT_NAMES",
           "FLAG_NO_PRUNE", "FLAG_HOST_FINALIZE", "FLAG_GENERAL_BREAK", "FLAG_SEED_ONLY", "FLAG_SKIP_SEED"]
