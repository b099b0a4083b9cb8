/* rr_maxcorr.h -- C ABI of the B200-native MaxCorrelation scan.
 *
 * Drop-in boundary for the hot path of /root/reference/MaxCorrelation.c.  The reference
 * has no FFI of its own: the path sits behind (1) the process boundary
 *     ./MaxCorrelation <MSA> [-c cov] [-p n]  ->  MaxCorrsOf_<MSA>      (main, 916-1026)
 * which bin/MaxCorrelation (csrc/rr_cli.c) reproduces on top of this header, and (2) the
 * function boundary inside the program, which the entry points below replace one to one:
 *
 *   reference (MaxCorrelation.c)                      this ABI
 *   ------------------------------------------------  ------------------------------------
 *   Einlesen(path, von, bis)            270-335       rr_msa_read / rr_msa_from_text
 *     (fgets rows, keep rule, char codes)
 *   Einlesen, packing part              339-385       rr_pack            (device side)
 *     (Groups, LocalCoverage, Groupsizearray, Coverage)
 *   Parallel_AllMaxCorrsRechner(NTHREADS, 0,          rr_maxcorr_run     (all GPUs, merged)
 *     siglength, mincov, signumber, cutoff) 839-908   rr_scan            (one GPU, one part)
 *     -> HilfsMaxCorrsRechner           745-837
 *     -> PositiveSignificance           421-434       fused epilogue, csrc/rr_score.h
 *     -> Schnitt                        114-125       count kernels (bitset / tcgen05)
 *   MaxCorrsRausschreiben(M, name)      516-532       rr_maxcorr_write
 *
 * Conventions: plain pointers and sizes only; every function returns RR_OK (0) or a
 * negative RR_E_* code and never calls exit(); rr_last_error() gives a message for the
 * calling thread.  Handles are not thread-safe; distinct handles may be used from
 * distinct threads.  There is no CPU implementation of the scan behind this ABI: without
 * a CUDA device rr_pack / rr_scan / rr_maxcorr_run fail with RR_E_NODEV.
 */
#ifndef RR_MAXCORR_H
#define RR_MAXCORR_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RR_OK 0
#define RR_E_IO (-1)    /* cannot open/read/write a file ("MA is missing.", MaxCorrelation.c:284) */
#define RR_E_NOMEM (-2) /* host or device allocation failed (Guarded_Malloc, 41-51) */
#define RR_E_CUDA (-3)  /* a CUDA call or kernel failed */
#define RR_E_ARG (-4)   /* invalid argument */
#define RR_E_NODEV (-5) /* no usable CUDA device (there is no CPU fallback) */

/* which count kernel computes the read-set intersections (Schnitt, 114-125) */
enum {
    RR_VARIANT_AUTO = 0,   /* the fastest one by measurement: RR_VARIANT_UMMA_MXF4 */
    RR_VARIANT_BITSET = 1, /* shared-memory staged u32 bitsets, AND + POPC */
    RR_VARIANT_UMMA = 2,   /* int8 0/1 operands, tcgen05.mma kind::i8, int32 accumulators in TMEM */
    RR_VARIANT_UMMA_F4 = 3,/* the same GEMM with the 0/1 operands stored as packed 4-bit e2m1 (half the HBM/L2
                              bytes; TMA unpacks them into shared memory), tcgen05.mma kind::f8f6f4 at the 8-bit
                              rate, fp32 accumulators in TMEM (exact: counts < 2^24) */
    RR_VARIANT_UMMA_MXF4 = 4 /* the packed e2m1 operands fed to tcgen05.mma kind::mxf4.block_scale with unit (2^0)
                              block scales: 4-bit operands stay packed in shared memory, twice the 8-bit MMA rate */
};

/* flags */
#define RR_FLAG_NO_PRUNE 1u      /* evaluate the exact score of every pair test (no bound-based skipping) */
#define RR_FLAG_HOST_FINALIZE 2u /* rr_maxcorr_run only: re-evaluate each group's winning pair with the host libm so
                                    that the "%f" text is byte-identical to the reference's (device exp/log10 differ
                                    from glibc by <= 2 ulp).  rr_scan rejects it: callers that merge several rr_scan
                                    results themselves call rr_scan_finalize after their merge */
#define RR_FLAG_GENERAL_BREAK 4u /* force the general first-break computation (MaxCorrelation.c:807-810)
                                    even when every row is one contiguous span */

#define RR_FLAG_SEED_ONLY 8u     /* run only the seeding pass (every 64th row tile): multi-GPU runs
                                    exchange the resulting maxima as thresholds before the full pass.  With the
                                    tcgen05 variants the seeded values are rigorous LOWER bounds of scores that
                                    pairs of the seed tiles attain (within ~1e-3 of them), entries without a
                                    partner: good for thresholds, not results */
#define RR_FLAG_SKIP_SEED 16u    /* keep the running maxima already on the device (previous RR_FLAG_SEED_ONLY
                                    scan and/or rr_scan_set_thresholds) and go straight to the full pass */

typedef struct rr_msa rr_msa;       /* host: the kept rows of an MSA */
typedef struct rr_packed rr_packed; /* device: one GPU's packed copy of an MSA */

typedef struct rr_scan_opts {
    int mincov;      /* the -c value (default 30, MaxCorrelation.c:925) */
    int variant;     /* RR_VARIANT_* */
    unsigned flags;  /* RR_FLAG_* */
    int part_index;  /* this scan covers part part_index of part_count pair-balanced */
    int part_count;  /*   contiguous ranges of row sites (1 -> the whole MSA) */
} rr_scan_opts;

typedef struct rr_scan_stats {
    int64_t pair_tests;   /* PositiveSignificance calls the reference would make (820) */
    int64_t exact_evals;  /* pairs whose exact score was evaluated */
    int64_t bound_evals;  /* pairs that went through the queued bound tier (tcgen05 variants: tier 2) */
    int64_t work_units;   /* tile pairs processed */
    int64_t executed_ops; /* tensor-core MACs*2 (UMMA variants) or 32-bit AND+POPC word ops (bitset) executed */
    int variant;          /* variant actually used */
    int rows, cols;       /* R, N */
    int row_sites;        /* sites with at least one admissible row group */
    int general_break;    /* 1 if the general first-break path was used */
    float h2d_ms, pack_ms, prepare_ms, kernel_ms, fetch_ms, finalize_ms;
} rr_scan_stats;

/* ---- host: reading the MSA (Einlesen, reading part) -------------------------------- */
/* rr_msa_read applies the row rules of 291/299 to the file and keeps it memory-mapped: nothing is copied on the
 * host, rr_pack gathers the kept rows straight into its upload ring.  The file must stay unchanged until
 * rr_msa_free (or the first rr_msa_cells call, which materialises the matrix and drops the mapping).
 * rr_msa_from_text does the same for text in caller-owned memory and copies the kept rows out. */
int rr_msa_read(const char *path, rr_msa **out);
/* The reader of RepeatResolver.c (Einlesen 293-429), in front of Group_Refinement, Relative_Vars and Kmeans: the columns
 * von .. bis of the reads that carry a symbol (anything but ' ') at both ends of the window (330).  bis is lowered to the
 * last column of any shorter line, from that line on (328), and the lowered value ends the window that is kept (398).
 * ausgelassen[capacity] (may be NULL with capacity 0) receives 1 per line that was kept and -1 per line left out (332, 367);
 * *n_lines = lines of the file, also when they exceed the capacity (RR_E_ARG then).  RR_E_IO for a missing file ("MA is
 * missing.", 318), a last line without newline (326: the reference exits) or a line too short to hold column von. */
int rr_msa_read_window(const char *path, int von, int bis, rr_msa **out, int8_t *ausgelassen, int64_t capacity, int64_t *n_lines);
int rr_msa_from_text(const char *text, size_t nbytes, rr_msa **out);
/* cells[rows][cols]; codes != 0: values 0..5 as in Signatures (304-329); else raw characters */
int rr_msa_from_cells(const uint8_t *cells, int rows, int cols, int codes, rr_msa **out);
/* allocate an uninitialised rows x cols cell matrix (page-locked when a device exists) to be
 * filled in place through rr_msa_cells() */
int rr_msa_alloc(int rows, int cols, int codes, rr_msa **out);
int rr_msa_rows(const rr_msa *msa);
int rr_msa_cols(const rr_msa *msa);
/* [rows][cols] cell matrix owned by the handle (NULL on allocation failure, see rr_last_error) */
uint8_t *rr_msa_cells(rr_msa *msa);
void rr_msa_free(rr_msa *msa);

/* ---- device ------------------------------------------------------------------------ */
int rr_device_count(void);
/* 1 if the count-kernel variant is compiled into this library */
int rr_variant_available(int variant);
/* copy the cells to `device` and pack them there (bitsets, int8 operands, group sizes,
 * coverage, read spans) */
int rr_pack(const rr_msa *msa, int device, rr_packed **out);
void rr_packed_free(rr_packed *pk);
/* The same in phases, so that n GPUs share the upload and the packing (each GPU still ends up with the whole packed
 * MSA; Einlesen's GrAdd loop 366-384 is the part that is split):
 *   rr_pack_rows        upload the rows [row_lo, row_hi) only and find their covered spans
 *   rr_pack_slice_spans start / end / covered cells of those rows (host arrays of row_hi - row_lo entries) - the
 *                       caller gathers them from all slices
 *   rr_pack_set_spans   the spans of ALL rows (host arrays of rows entries, in row order): fixes the row order and
 *                       packs this slice's rows into full-size group / coverage bitsets, zero for all other rows
 *   rr_pack_bits_device device pointer and size of that buffer ([5 cols][W] group words, then [cols][W] coverage
 *                       words, 32-bit); the caller ORs the buffers of all slices into it (the slices are disjoint, so
 *                       an integer SUM all-reduce - e.g. NCCL on this pointer - is an OR)
 *   rr_pack_finish      group sizes, coverage and result buffers from the merged bitsets; the handle is then what
 *                       rr_pack returns
 * rr_maxcorr_run does this between the GPUs of one process with peer copies; repeatresolver_b200/dist.py between
 * one-process-per-GPU ranks with NCCL. */
int rr_pack_rows(const rr_msa *msa, int device, int row_lo, int row_hi, rr_packed **out);
int rr_pack_slice_spans(rr_packed *pk, int32_t *start, int32_t *end, int32_t *count);
int rr_pack_set_spans(rr_packed *pk, const int32_t *start, const int32_t *end, const int32_t *count);
int rr_pack_bits_device(rr_packed *pk, void **d_bits, size_t *bytes);
int rr_pack_finish(rr_packed *pk);
/* run the scan on the packed MSA; results stay on the device */
int rr_scan(rr_packed *pk, const rr_scan_opts *opts, rr_scan_stats *stats);
/* copy the last scan's result to the host: maxcorr[5*cols] (line g of MaxCorrsOf_*,
 * g = 5*site + {A,C,G,T,gap}); argmax[5*cols] = partner group id or -1 (may be NULL).  After a scan of the whole MSA
 * every group with a maximum has its partner.  After a scan of one part (part_count > 1) an entry without a partner
 * (argmax -1, value > 0) is a threshold this part was given or derived - a lower bound of the group's maximum, attained
 * or exceeded in another part - and disappears in the element-wise max over the parts */
int rr_scan_fetch(rr_packed *pk, double *maxcorr, int32_t *argmax);
/* RR_FLAG_HOST_FINALIZE as a call: maxcorr[g] of every group with a partner (argmax[g] >= 0) is re-evaluated with
 * the host libm from the pair's device-side counts; maxcorr / argmax as returned by rr_scan_fetch or merged from
 * several parts (element-wise max, ties to the smaller partner) */
int rr_scan_finalize(rr_packed *pk, double *maxcorr, const int32_t *argmax);
/* raise the running maxima on the device to at least thr[5*cols] (e.g. the max over all GPUs' seeding passes);
 * a value taken from thr carries no partner: it only prunes, and loses every tie against a real pair */
int rr_scan_set_thresholds(rr_packed *pk, const double *thr);
/* the same exchange without the host: d_values / d_thr are DEVICE pointers (same device and process as the
 * handle, e.g. a torch tensor's data_ptr) to 5*cols doubles; both calls return after their work has completed */
int rr_scan_values_device(rr_packed *pk, double *d_values);
int rr_scan_set_thresholds_device(rr_packed *pk, const double *d_thr);
/* the four intersection counts {schnitt, gr1, gr2, cov} (423-426) of n group pairs, from the
 * device bitsets: out[4*n] */
int rr_pair_counts(rr_packed *pk, int64_t n, const int32_t *gi, const int32_t *gj, int32_t *out);
/* Groupsizearray (385) and Coverage (364-383) as computed on the device */
int rr_packed_sizes(rr_packed *pk, int32_t *gsize /*[5*cols]*/, int32_t *coverage /*[cols]*/);

/* ---- the whole path ------------------------------------------------------------------ */
/* Parallel_AllMaxCorrsRechner replacement: pack on n_gpus devices, scan one part per GPU,
 * merge by element-wise max (882-891).  maxcorr_out[5*cols], argmax_out[5*cols] or NULL. */
int rr_maxcorr_run(const rr_msa *msa, int mincov, int n_gpus, int variant, unsigned flags,
                   double *maxcorr_out, int32_t *argmax_out, rr_scan_stats *stats);
/* MaxCorrsRausschreiben: count lines "%f\n" */
int rr_maxcorr_write(const char *path, const double *maxcorr, int64_t count);
int rr_argmax_write(const char *path, const int32_t *argmax, int64_t count);
/* Binary side format, "MaxCorrsBinOf_<MSA>": magic "RRMAXC01", int64 count, int64 flags (bit 0:
 * partners stored), count doubles, count int32; little endian.  argmax may be NULL.  The text file remains the contract
 * with the unmodified consumers (RepeatResolver.c:609-646, TransposonAssessment.py:51-58). */
int rr_maxcorr_write_bin(const char *path, const double *maxcorr, const int32_t *argmax, int64_t count);
/* The window MaxCorrsEinlesen(file, von, bis) keeps (RepeatResolver.c:631: columns von..bis inclusive), clipped to the
 * file; outputs [5 * (bis - von + 1)] or NULL, *n_out = values delivered; as_text != 0 rounds the values as the "%f"
 * text file would, so a consumer decides exactly as it would on MaxCorrsOf_*. */
int rr_maxcorr_read_bin(const char *path, int von, int bis, int as_text, double *maxcorr_out, int32_t *argmax_out, int64_t *n_out);
/* MaxCorrsEinlesen itself (RepeatResolver.c:609-646) on the text file MaxCorrsOf_<MSA>: the values of the lines i with
 * von <= i / 5 <= bis in file order; a line is what fgets(.., 100, ..) delivers and its value what sscanf("%lf") reads (0.0
 * where there is no number; the reference leaves that slot uninitialised).  maxcorr_out[capacity] may be NULL to count;
 * *n_out = values in the window (RR_E_ARG if they exceed the capacity); RR_E_IO if the file cannot be opened (the reference
 * returns NULL, 619). */
int rr_maxcorr_read_text(const char *path, int von, int bis, double *maxcorr_out, int64_t capacity, int64_t *n_out);

/* ---- host-side pieces exposed for tests and for RR_FLAG_HOST_FINALIZE ------------------ */
double rr_lnfact(unsigned int n); /* gsl_sf_lnfact */
void rr_lnfact_table(double *out, size_t count);
/* PositiveSignificance (421-434) on counts, host libm */
double rr_score_host(uint32_t schnitt, uint32_t gr1, uint32_t gr2, uint32_t cov, int32_t sizei, int32_t sizej);
/* the pruning bounds of csrc/rr_score.h, host build */
double rr_score_bound_host(uint32_t schnitt, uint32_t gr1, uint32_t gr2, uint32_t cov);
int rr_below_median_host(uint32_t schnitt, uint32_t gr1, uint32_t gr2, uint32_t cov);
/* first-break columns (807-810) for rows that are single spans: start[r], end[r] inclusive */
int rr_breakcols_from_spans(const int32_t *start, const int32_t *end, int rows, int cols, int mincov,
                            int32_t *breakcol /*[cols]*/);

/* ---- next scope row: Cliquer, the inner step of Group_Refinement ----------------------
 * RepeatResolver.c:1179-1240: the up to maxclique-1 groups of columns [anfang, ende) that correlate best with a query
 * group (Group_PositiveSignificance 472-488 > greedy, intersection > mincov/4), best first, ties in group order
 * (TheBestUpdater 1156-1176).  Per query: members[maxclique+1], members[0] = the query group, unused slots -1 (the
 * reference's -1 terminator, 1230); scores[maxclique], scores[0] = 100 (1229), unused slots 0; n_members counts the
 * query group itself.  Scores are evaluated (finally) with the host libm on exact integer counts, so members and
 * scores are bit-identical to the reference's.
 *
 * rr_cliquer_batch is the product path: all queries of a Group_Refinement pass (the groups with MaxCorrs > cutoff,
 * 1647-1649) in one call; counts, filter, score bound and score run on the device (csrc/rr_cliquer.cu), the host
 * re-evaluates and orders the few hits per query that decide its clique.  rr_cliquer is the plain single-query
 * implementation (device counts of every candidate through rr_pair_counts, every score on the host) the tests hold
 * it against.  Queries are independent: several GPUs take disjoint slices of the query list. */
typedef struct rr_cliquer_stats {
    int64_t pairs;        /* (query, candidate group) pairs tested = Schnitt calls of 1213 */
    int64_t candidates;   /* pairs above mincov/4 whose score bound exceeds greedy */
    int64_t hits;         /* candidates whose device score exceeds greedy */
    int64_t host_evals;   /* scores re-evaluated with the host libm */
    float kernel_ms;      /* CUDA-event time of the two kernels, summed over launches */
    int launches;
    int retries;          /* launches repeated with fewer queries because the candidate list overflowed */
} rr_cliquer_stats;
int rr_cliquer_batch(rr_packed *pk, int64_t n_queries, const int32_t *query_groups, int anfang, int ende, int mincov,
                     int maxclique, double greedy, int32_t *members /*[n_queries][maxclique+1]*/,
                     double *scores /*[n_queries][maxclique]*/, int32_t *n_members /*[n_queries]*/,
                     rr_cliquer_stats *stats /* may be NULL */);
/* CliqueGroup / CliqueCoverage (RepeatResolver.c:976-1008, 1064-1096), the step that follows Cliquer in Group_Refinement
 * (1662-1664), for a batch of cliques: members[q * stride + m], m < n_members[q] (at most 100, the reference's own limit
 * 986), are the clique's groups - the query itself first, as Cliquer returns it.
 *   groups[q]   = the reads contained in MORE than cutoffs[q] of the member groups
 *   coverage[q] = the reads covered at more than cutoffs[q] of the members' sites
 * each as sc = rows / 64 + 1 words (RepeatResolver.c:59) in the reference's layout: read r (row of the MSA pk was made of)
 * is bit r % 64 of word r / 64 (GrAdd 211-217).  Either output may be NULL.  A negative cutoff selects every read, as
 * the reference's loop does.  Device work on the packed bitsets: 32 reads per instruction with bit-sliced counters
 * (csrc/rr_cliquer.cu), every member word read once. */
int rr_clique_groups(rr_packed *pk, int64_t n_cliques, const int32_t *members, int stride, const int32_t *n_members,
                     const int32_t *cutoffs, uint64_t *groups /*[n_cliques][sc]*/, uint64_t *coverage /*[n_cliques][sc]*/);
/* Group_Refinement as a whole (RepeatResolver.c:1634-1693; Parallel_Group_Refinement 1770-1821 computes the same, group by
 * group, on threads - note that its int argument array truncates cutoff and greedy, 1793-1794: callers that replace the
 * parallel form pass (double)(int)cutoff and (double)(int)greedy).  For every group i with maxcorrs[i] > cutoff, in
 * ascending order (slot q):
 *   query_groups[q] = i
 *   cliques[q][maxclique + 1] = Cliquer(anfang, ende, mincov, maxclique, greedy, i)                              (1649)
 *   sizes[q]     = Sizes[i]: the members before the first entry <= 0 - group 0 ends the count like the -1 does     (1650)
 *   where sizes[q] > 5:
 *     cutoffs[q]  = Cutoffs[i] = Dropoff_Cutoff(i, 0) (1460-1522): the k in [1, Sizes - 1) that minimises
 *                   (n[k-1] - n[k+1]) / min(rows - n[k], n[k]), n[k] = the reads contained in more than k of the first
 *                   Sizes members; the first minimum wins, 1 if no k is admissible.  (The reference evaluates BestCutoff
 *                   530-548 and KorrMaxCutoff 1393-1457 first; both results are overwritten at 1662, neither has a side
 *                   effect, so they are not evaluated here.)
 *     drop_off[q] = Drop_Off[i], the minimum itself (1e6 if no k is admissible)
 *     c_groups[q][sc], c_coverage[q][sc] = CliqueGroup / CliqueCoverage of the clique at that cutoff             (1663-1665)
 *   else cutoffs[q] = 0, drop_off[q] = 1000.0 (1641-1645), c_groups[q] / c_coverage[q] all zero (the reference leaves NULL)
 *     and maxcorrs[i] = 0.0                                                                                       (1685)
 * capacity = slots the caller provides; *n_queries = groups above the cutoff, also when that exceeds the capacity (then
 * RR_E_ARG; maxcorrs is untouched and only query_groups has been written to).  c_groups / c_coverage may be NULL.  maxclique <= 100 (the reference's array of
 * 100 cutoff groups, 1463).  The reference's two "Group Precision" printouts per refined group (1678-1679) are diagnostics
 * of the caller's (repeatresolver_b200.GroupPrecision computes the same two numbers from a bitset).
 * Device work: rr_cliquer_batch, the member counts n[k] of all refined cliques in one kernel (bit-sliced counters, one
 * comparison per k on 32 reads at a time), rr_clique_groups; the cutoff rule itself runs on the host in IEEE double on
 * exact integers, so Cutoffs and Drop_Off are bit-identical to the reference's. */
int rr_group_refinement(rr_packed *pk, double *maxcorrs /* [5 * cols], in/out */, double cutoff, int anfang, int ende, int mincov,
                        int maxclique, double greedy, int64_t capacity, int32_t *query_groups /*[capacity]*/,
                        int32_t *cliques /*[capacity][maxclique+1]*/, int32_t *sizes /*[capacity]*/, int32_t *cutoffs /*[capacity]*/,
                        double *drop_off /*[capacity]*/, uint64_t *c_groups /*[capacity][sc]*/, uint64_t *c_coverage /*[capacity][sc]*/,
                        int64_t *n_queries, rr_cliquer_stats *stats /* may be NULL */);
/* the cutoff rule alone (tests): sizes[k] = reads in more than k of the `size` members, k < size; returns the cutoff */
int rr_dropoff_cutoff_host(const uint32_t *sizes, int size, int signumber, int c, double *drop_off);
int rr_cliquer(rr_packed *pk, int query_group, int anfang, int ende, int mincov, int maxclique, double greedy,
               int32_t *members, double *scores, int *n_members);
/* the host half on given counts (tests): groups[k] ascending candidate ids, counts[4k..4k+3] =
 * {|Gk & Gq|, |Gk & Cq|, |Gq & Ck|, |Ck & Cq|}, sizes[k] = |Gk| */
int rr_cliquer_from_counts(int query_group, int64_t n, const int32_t *groups, const int32_t *counts, const int32_t *sizes,
                           int size_query, int mincov, int maxclique, double greedy, int32_t *members, double *scores,
                           int *n_members);
double rr_group_score_host(uint32_t schnitt, uint32_t gr1, uint32_t gr2, uint32_t cov, int32_t sizei, int32_t sizej);
/* the host half of rr_cliquer_batch on a given hit list (tests): hit_records[n_hits] of 32 bytes each,
 * {int32 slot (index into query_groups), group, s, gr1, gr2, cov; double device score}, in any order; gsize[n_groups] */
int rr_cliquer_from_hits(int64_t n_queries, const int32_t *query_groups, int64_t n_hits, const void *hit_records,
                         const int32_t *gsize, int64_t n_groups, int mincov, int maxclique, double greedy,
                         int32_t *members, double *scores, int32_t *n_members);

/* ---- scope row 8f-3: Relative_Vars (RepeatResolver.c:2424-2493) ----------------------------------------------------
 * Which groups vary inside part u_no of a read partition: groups with MaxCorrs > cutoff that hold at least mingroup
 * reads of the part and have a partner at least 100 group ids away whose two-sided hypergeometric score restricted
 * to the part's reads (Relative_Group_Significance 506-523, CumHypGeo_Log 490-504) exceeds cutoff.
 * unterteilung: [rows of msa] part number of every read; maxcorrs: [5 * cols]; vars: [5 * cols + 1], receives the
 * ascending group ids followed by -1 (the reference's terminator, 2483).  mingroup >= 1 and cutoff >= 0 are required.
 * rr_relative_vars packs the part's rows as an MSA of their own on `device` (rr_pack), which turns the triple
 * intersections into pair intersections; counts, score bound, exact score and marks run in one tiled kernel
 * (csrc/rr_relvars.cu); pairs whose device score lies within 1e-9 of the cutoff are decided with the host libm, so the
 * selection is identical to the reference's. */
int rr_relative_vars(const rr_msa *msa, int device, const int32_t *unterteilung, int u_no, const double *maxcorrs,
                     double cutoff, int mingroup, int32_t *vars, int *n_vars, int64_t *pairs_tested /* may be NULL */);
/* the same on the packed copy of the whole MSA already on the device (e.g. after rr_scan): nothing is packed again, the
 * part is applied as a mask inside the kernel; unterteilung indexes the rows of the MSA pk was made of */
int rr_relative_vars_packed(rr_packed *pk, const int32_t *unterteilung, int u_no, const double *maxcorrs, double cutoff,
                            int mingroup, int32_t *vars, int *n_vars, int64_t *pairs_tested /* may be NULL */);
/* the host half on given counts (tests): gsize_u[g] = |G_g & U|, cov_u = |U|; with S == NULL only the selection
 * (sel_out, n_sel) is returned, else S[n_sel][n_sel] holds |G_sel[a] & G_sel[b] & U| */
int rr_relative_vars_from_counts(int64_t n_groups, const double *maxcorrs, const int32_t *gsize_u, int cov_u, double cutoff,
                                 int mingroup, const int32_t *S, int32_t *sel_out, int *n_sel, int32_t *vars, int *n_vars);
double rr_relative_score_host(uint32_t schnitt, uint32_t gr1, uint32_t gr2, uint32_t cov);

/* ---- scope row 8f-4: Kmeans (RepeatResolver.c:2604-2821) ------------------------------------------------------------
 * Splits part u_no of the read partition by the reads' signatures over the groups `vars` (the output of Relative_Vars):
 * unterteilung[rows of msa] is updated in place exactly as the reference does (the part's reads get cluster + max + 1,
 * 2814-2815), *n_clusters = the reference's return value (non-empty clusters).  All arithmetic is integer: bit-exact.
 * The part's rows are copied to `device` as they are; signatures (2626-2650), the two read x read sweeps (2662-2725), the
 * centroids and the scores the dissolution of small clusters looks up (2735-2745) are computed there (csrc/rr_kmeans.cu);
 * the dissolution itself, sequential by definition, runs on the host. */
int rr_kmeans(const rr_msa *msa, int device, int32_t *unterteilung, int u_no, const int32_t *vars, int n_vars, int mingroup,
              int *n_clusters);
/* host pieces (tests): the part's reads and - as a second implementation the device's are held against - their signatures
 * ([anzahl][n_vars/64+1] 64-bit words; sig_out may be NULL to get the count), the dissolution of small clusters given the
 * device's first assignment (scores computed as it goes), and the two integer rules the kernels share with the host
 * (csrc/rr_kmeans.h) */
int rr_kmeans_signatures(const rr_msa *msa, const int32_t *unterteilung, int u_no, const int32_t *vars, int n_vars,
                         int32_t *reads_out /*[rows]*/, int *anzahl_out, uint64_t *sig_out);
int rr_kmeans_finish(int anzahl, int scv, const uint64_t *sig, const uint64_t *cen, const int32_t *cluster_in, int mingroup,
                     int32_t *cluster_out, int *n_clusters);
int rr_kmeans_top5_host(int anzahl, int scv, const uint64_t *sig, int i, int32_t *best_j /*[5]*/);
uint64_t rr_kmeans_majority5_host(uint64_t a, uint64_t b, uint64_t c, uint64_t d, uint64_t e);

/* the exact contraction ranges of the scan plan (csrc/rr_plan.h), for tests: rows in rank order with spans
 * start[r]..end[r] (inclusive); class c = ranks [class_start[c], class_start[c + 1]), each class sorted by start; every site
 * is taken as a row site, ti / tj sites per row / column block, kunit rows per contraction unit.  A row of class c can
 * only contribute to (row block rb, column block cb) if its rank lies in
 * [k_lo[c * n_colblocks + cb] * kunit, k_hi[c * n_rowblocks + rb] * kunit).  k_lo: [n_classes * ceil(cols / tj)],
 * k_hi: [n_classes * n_rowblocks] with n_rowblocks = ceil(max(cols - 20, 0) / ti) returned through *n_rowblocks. */
int rr_contraction_ranges(const int32_t *start, const int32_t *end, int rows, int cols, int n_classes, const int32_t *class_start,
                          int ti, int tj, int kunit, int32_t *k_lo, int32_t *k_hi, int *n_rowblocks);
/* the length classes rr_pack sorts the rows into (rows sorted by span length, cut at fixed fractions rounded down to whole
 * 256-row blocks; one class below 1024 rows): returns the number of classes, class_start[that + 1] (room for 9) */
int rr_length_classes(int rows, int32_t *class_start);

/* ---- measurement helpers ------------------------------------------------------------------ */
/* CUDA events on the stream every kernel of this handle is launched on */
int rr_timer_start(rr_packed *pk);
int rr_timer_stop(rr_packed *pk, float *elapsed_ms);
/* number of kernels this library has launched in this process so far */
int64_t rr_launch_count(void);

const char *rr_last_error(void);
const char *rr_version(void);

#ifdef __cplusplus
}
#endif
#endif /* RR_MAXCORR_H */
