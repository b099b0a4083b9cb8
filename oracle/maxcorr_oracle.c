/* TEST INFRASTRUCTURE ONLY (oracle/).  Not part of the product: only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * link or execute this file.  The product (repeatresolver_b200/) never does.
 *
 * CPU restatement, in plain C, of the MaxCorrelation all-pairs scan of
 * /root/reference/MaxCorrelation.c (the `-p >= 1` path), written from the
 * specification in SURVEY.md Appendix A.  Each function cites the reference lines
 * it follows.  Differences from the reference, all deliberate:
 *   - no static capacity limits (Max_Var_Anzahl / Max_Sig_Anzahl, lines 18-19);
 *   - instrumented: counts pair tests and records the arg-max partner
 *     (smallest partner group index attaining the max);
 *   - F_beta(.,.,1.0) (lines 396-411) is evaluated in closed form
 *     2s/(|Gi|+|Gj|): its operands are exact integers held in doubles, so the one
 *     division is bit-identical to the reference's three extra bitset passes;
 *   - a handle instead of process globals.
 * The significance itself goes through gsl_cdf_hypergeometric_Q of
 * oracle/gsl_shim.c, exactly as MaxCorrelation.c:415 goes through GSL
 * (GSL version unpinned by the reference -> parity unpinned at that boundary,
 * see gsl_shim.c).  This restatement is validated against the UNMODIFIED
 * reference compiled into oracle/_ref (tests/test_oracle.py, and the
 * committed fixtures under tests/golden/ made by oracle/gen_golden.py).
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <stdint.h>
#include <math.h>
#include <pthread.h>
#include "gsl/gsl_cdf.h"

typedef struct rr_oracle {
    int R;            /* signumber: rows kept                (MaxCorrelation.c:335) */
    int N;            /* siglength: columns                  (MaxCorrelation.c:291) */
    int sc;           /* words per bitset = R/64+1           (MaxCorrelation.c:339) */
    uint64_t *groups; /* [5N][sc]  Groups                    (MaxCorrelation.c:348-351) */
    uint64_t *cover;  /* [N][sc]   LocalCoverage             (MaxCorrelation.c:354-357) */
    int *gsize;       /* [5N]      Groupsizearray            (MaxCorrelation.c:385) */
    int *coverage;    /* [N]       Coverage                  (MaxCorrelation.c:364-383) */
} rr_oracle;

/* MaxCorrelation.c:304-329: character -> code */
static inline int rr_code_of(unsigned char c)
{
    switch (c) {
    case 'a': case 'A': return 0;
    case 'c': case 'C': return 1;
    case 'g': case 'G': return 2;
    case 't': case 'T': return 3;
    case '-': case '_': return 4;
    default: return 5;
    }
}

/* MaxCorrelation.c:114-125 (Schnitt) with popcount_3 (80-86) replaced by the builtin */
static inline int rr_isect(const uint64_t *a, const uint64_t *b, int sc)
{
    int z, n = 0;
    for (z = 0; z < sc; z++) n += __builtin_popcountll(a[z] & b[z]);
    return n;
}

void rr_oracle_free(rr_oracle *o)
{
    if (!o) return;
    free(o->groups); free(o->cover); free(o->gsize); free(o->coverage); free(o);
}

/* MaxCorrelation.c:339-385: bitsets, coverage and group sizes from a code matrix
 * codes[R][N] (row-major, values 0..5). */
rr_oracle *rr_oracle_from_codes(const uint8_t *codes, int R, int N)
{
    rr_oracle *o = (rr_oracle *)calloc(1, sizeof(*o));
    int i, r;
    o->R = R; o->N = N; o->sc = R / 64 + 1;
    o->groups = (uint64_t *)calloc((size_t)5 * N * o->sc + 1, 8);
    o->cover = (uint64_t *)calloc((size_t)N * o->sc + 1, 8);
    o->gsize = (int *)calloc((size_t)5 * N + 1, sizeof(int));
    o->coverage = (int *)calloc((size_t)N + 1, sizeof(int));
    for (r = 0; r < R; r++) {
        const uint8_t *row = codes + (size_t)r * N;
        uint64_t bit = 1ull << (r % 64);
        int w = r / 64;
        for (i = 0; i < N; i++) {
            int c = row[i];
            if (c < 5) {
                o->groups[((size_t)5 * i + c) * o->sc + w] |= bit; /* GrAdd, 189-195, 374 */
                o->cover[(size_t)i * o->sc + w] |= bit;             /* 380 */
                o->coverage[i]++;                                    /* 381 */
                o->gsize[5 * i + c]++;                               /* 385 */
            }
        }
    }
    return o;
}

/* MaxCorrelation.c:270-335 (Einlesen, reading part): first line fixes siglength =
 * strlen-1; a later line is kept iff strlen-1 == siglength; others are skipped. */
rr_oracle *rr_oracle_load(const char *path)
{
    FILE *f = fopen(path, "r");
    char *line = NULL;
    size_t cap = 0;
    ssize_t len;
    int N = -1, R = 0, rcap = 0, i;
    uint8_t *codes = NULL;
    rr_oracle *o;
    if (!f) return NULL; /* "MA is missing." exit(1) at line 284 */
    while ((len = getline(&line, &cap, f)) >= 0) {
        size_t sl = strlen(line); /* fgets/strlen semantics: stops at an embedded NUL */
        if (N < 0) N = (int)sl - 1;
        if ((int)sl - 1 != N) continue;
        if (R == rcap) {
            rcap = rcap ? rcap * 2 : 256;
            codes = (uint8_t *)realloc(codes, (size_t)rcap * (N > 0 ? N : 1));
        }
        for (i = 0; i < N; i++) codes[(size_t)R * N + i] = (uint8_t)rr_code_of((unsigned char)line[i]);
        R++;
    }
    free(line);
    fclose(f);
    if (N < 0) N = 0;
    o = rr_oracle_from_codes(codes ? codes : (const uint8_t *)"", R, N);
    free(codes);
    return o;
}

int rr_oracle_R(const rr_oracle *o) { return o->R; }
int rr_oracle_N(const rr_oracle *o) { return o->N; }
const int *rr_oracle_gsize(const rr_oracle *o) { return o->gsize; }
const int *rr_oracle_coverage(const rr_oracle *o) { return o->coverage; }

/* The four intersections of PositiveSignificance, MaxCorrelation.c:423-426.
 * out = {schnitt, gr1, gr2, cov} for group ids i, j (id = 5*site+group). */
void rr_oracle_counts(const rr_oracle *o, int i, int j, int out[4])
{
    int sc = o->sc;
    const uint64_t *gi = o->groups + (size_t)i * sc, *gj = o->groups + (size_t)j * sc;
    const uint64_t *ci = o->cover + (size_t)(i / 5) * sc, *cj = o->cover + (size_t)(j / 5) * sc;
    out[0] = rr_isect(gi, gj, sc);
    out[3] = rr_isect(ci, cj, sc);
    out[1] = rr_isect(gi, cj, sc);
    out[2] = rr_isect(gj, ci, sc);
}

/* Schnitt (114-125) for every pair of two group lists: out[a * nc + b] = |G_rows[a] & G_cols[b]| */
void rr_oracle_count_matrix(const rr_oracle *o, int nr, const int32_t *rows, int nc, const int32_t *cols, int32_t *out)
{
    int a, b, sc = o->sc;
    for (a = 0; a < nr; a++)
        for (b = 0; b < nc; b++)
            out[(size_t)a * nc + b] = rr_isect(o->groups + (size_t)rows[a] * sc, o->groups + (size_t)cols[b] * sc, sc);
}

/* MaxCorrelation.c:413-434 (PositiveCumHypGeo_Log + PositiveSignificance tail)
 * on already-computed counts; sizei/sizej = Groupsizearray of the two groups. */
double rr_oracle_score(unsigned int schnitt, unsigned int gr1, unsigned int gr2, unsigned int cov,
                       int sizei, int sizej)
{
    double Z;
    if (gr1 == 0 || gr2 == 0) return 0.0;                      /* 428 */
    if ((int)schnitt < 1) return 0.0;                          /* 430 */
    Z = gsl_cdf_hypergeometric_Q(schnitt - 1, gr2, cov - gr2, gr1); /* 415 */
    Z = -1.0 * log10(Z);                                       /* 416 */
    if (isinf(Z) || Z > 99) Z = 99.0;                          /* 417 */
    if (isinf(Z) || Z > 98.0) {                                /* 432 */
        /* F_beta(Gi,Gj,1.0), 396-411: (1+1)*s / ((1+1*1)*s + |Gi\Gj| + |Gj\Gi|) */
        double s = (double)schnitt;
        double g1n2 = (double)(sizei - (int)schnitt);
        double g2n1 = (double)(sizej - (int)schnitt);
        double F = (1.0 + 1.0) * s;
        if (F < 0.0001) F = 0.0;
        else F /= ((1 + 1.0 * 1.0) * s + (1.0 * 1.0 * g1n2) + g2n1);
        Z = 98.0 + F;
    }
    return Z;
}

typedef struct {
    const rr_oracle *o;
    int mincov, nthreads, thread;
    int count_only;  /* walk the loops of 796-817 and count the calls of 820 without making them */
    double *M;       /* [5N] private maxima          (768-770) */
    int32_t *arg;    /* [5N] private arg-max partner (A9) */
    int64_t pairs;   /* number of PositiveSignificance calls (820) */
} rr_scan_job;

static inline void rr_upd(double *M, int32_t *arg, int g, double Z, int partner)
{
    if (Z > M[g]) { M[g] = Z; arg[g] = partner; }                /* 822-823: strict > */
    else if (Z == M[g] && Z > 0.0 && partner < arg[g]) arg[g] = partner; /* A9 tie rule */
}

/* MaxCorrelation.c:745-837 (HilfsMaxCorrsRechner): rows ii with ii % nthreads == thread */
static void *rr_scan_thread(void *x)
{
    rr_scan_job *job = (rr_scan_job *)x;
    const rr_oracle *o = job->o;
    const int N = o->N, R = o->R, sc = o->sc, mincov = job->mincov;
    const int q = mincov / 4;
    int ii, jj, k, kk;
    for (ii = 0; ii < N; ii++) {
        int baseno;
        if (ii % job->nthreads != job->thread) continue;                           /* 796 */
        baseno = o->gsize[5 * ii] + o->gsize[5 * ii + 1] + o->gsize[5 * ii + 2] + o->gsize[5 * ii + 3]; /* 798 */
        for (k = 0; k < 5; k++) {
            int i = 5 * ii + k;
            if (!(o->gsize[i] > q && o->gsize[i] < R && baseno > o->coverage[ii] / 2)) continue; /* 802 */
            for (jj = ii + 20; jj < N; jj++) {                                     /* 804 */
                if (rr_isect(o->cover + (size_t)ii * sc, o->cover + (size_t)jj * sc, sc) < mincov)
                    break;                                                          /* 807-810 */
                for (kk = 0; kk < 5; kk++) {
                    int j = 5 * jj + kk, c[4];
                    double Z;
                    if (!(o->gsize[j] > q && o->gsize[j] < R)) continue;            /* 817 */
                    if (job->count_only) { job->pairs++; continue; }
                    rr_oracle_counts(o, i, j, c);
                    Z = rr_oracle_score(c[0], c[1], c[2], c[3], o->gsize[i], o->gsize[j]); /* 820 */
                    job->pairs++;
                    rr_upd(job->M, job->arg, i, Z, j);
                    rr_upd(job->M, job->arg, j, Z, i);
                }
            }
        }
    }
    return NULL;
}

/* MaxCorrelation.c:839-908 (Parallel_AllMaxCorrsRechner) generalised to a row sample:
 * the scan covers the rows ii with (ii % modulus) in [res_lo, res_hi); one pthread per
 * residue, merged by element-wise max (882-891).  modulus = res_hi - res_lo = nthreads
 * reproduces the reference run `-p nthreads`.  M [5N], arg [5N] (may be NULL),
 * returns the number of pair tests. */
int64_t rr_oracle_scan(const rr_oracle *o, int mincov, int modulus, int res_lo, int res_hi,
                       double *M, int32_t *arg)
{
    int nt = res_hi - res_lo, t, g;
    const int G = 5 * o->N;
    pthread_t *th = (pthread_t *)calloc(nt, sizeof(pthread_t));
    rr_scan_job *jobs = (rr_scan_job *)calloc(nt, sizeof(rr_scan_job));
    int64_t pairs = 0;
    for (t = 0; t < nt; t++) {
        jobs[t].o = o; jobs[t].mincov = mincov; jobs[t].nthreads = modulus; jobs[t].thread = res_lo + t;
        jobs[t].M = (double *)calloc(G + 1, sizeof(double));
        jobs[t].arg = (int32_t *)malloc((G + 1) * sizeof(int32_t));
        for (g = 0; g < G; g++) jobs[t].arg[g] = -1;
        pthread_create(&th[t], NULL, rr_scan_thread, &jobs[t]);
    }
    for (t = 0; t < nt; t++) pthread_join(th[t], NULL);
    for (g = 0; g < G; g++) { M[g] = 0.0; if (arg) arg[g] = -1; }
    for (t = 0; t < nt; t++) {
        for (g = 0; g < G; g++) {
            double Z = jobs[t].M[g];
            int p = jobs[t].arg[g];
            if (Z > M[g]) { M[g] = Z; if (arg) arg[g] = p; }
            else if (arg && Z == M[g] && Z > 0.0 && p >= 0 && p < arg[g]) arg[g] = p;
        }
        pairs += jobs[t].pairs;
        free(jobs[t].M); free(jobs[t].arg);
    }
    free(jobs); free(th);
    return pairs;
}

/* The number of PositiveSignificance calls (820) the scan of the rows ii with (ii % modulus) in [res_lo, res_hi) makes:
 * the same loops, filters and first-break test as rr_scan_thread, without the calls themselves (bench.py: the pair-test
 * count of a CPU sample, next to the product's own count). */
int64_t rr_oracle_count_pairs(const rr_oracle *o, int mincov, int modulus, int res_lo, int res_hi)
{
    int nt = res_hi - res_lo, t;
    pthread_t *th = (pthread_t *)calloc(nt, sizeof(pthread_t));
    rr_scan_job *jobs = (rr_scan_job *)calloc(nt, sizeof(rr_scan_job));
    int64_t pairs = 0;
    for (t = 0; t < nt; t++) {
        jobs[t].o = o; jobs[t].mincov = mincov; jobs[t].nthreads = modulus; jobs[t].thread = res_lo + t;
        jobs[t].count_only = 1;
        pthread_create(&th[t], NULL, rr_scan_thread, &jobs[t]);
    }
    for (t = 0; t < nt; t++) { pthread_join(th[t], NULL); pairs += jobs[t].pairs; }
    free(jobs); free(th);
    return pairs;
}

/* ---- the clique of one query group, /root/reference/RepeatResolver.c ----------
 * Restated for the round that builds the GPU path of Group_Refinement; validated against the unmodified
 * RepeatResolver.c behind oracle/ref_cliquer_driver.c (tests/test_oracle_cliquer.py). */

/* Group_PositiveSignificance, RepeatResolver.c:472-488, on already-computed counts: like PositiveSignificance but
 * without the schnitt < 1 brake (the caller only asks for schnitt > mincov/4) and saturating at 97.90 + F1. */
double rr_oracle_group_score(unsigned int schnitt, unsigned int gr1, unsigned int gr2, unsigned int cov, int sizei,
                             int sizej)
{
    double Z;
    if (gr1 == 0 || gr2 == 0) return 0.0;                                /* 482 */
    Z = gsl_cdf_hypergeometric_Q(schnitt - 1, gr2, cov - gr2, gr1);      /* 451 */
    Z = -1.0 * log10(Z);                                                 /* 452 */
    if (isinf(Z) || Z > 99) Z = 99.0;                                    /* 453 */
    if (isinf(Z) || Z > 98.0) {                                          /* 486: 97.90 + F_beta(G1, G2, 1.0), 432-447 */
        double s = (double)schnitt, F = (1.0 + 1.0) * s;
        if (F < 0.0001) F = 0.0;
        else F /= ((1 + 1.0 * 1.0) * s + (1.0 * 1.0 * (double)(sizei - (int)schnitt)) + (double)(sizej - (int)schnitt));
        Z = 97.90 + F;
    }
    return Z;
}

/* TheBestUpdater, RepeatResolver.c:1156-1176: insertion into the descending list, behind equal values; slot 0 (the
 * query group itself) is never displaced */
static void rr_best_updater(int32_t *clique, double *best, int maxclique, int i, double Z)
{
    int ii, j;
    if (best[maxclique - 1] >= Z) return;                                /* 1158 */
    ii = maxclique - 1;
    while (best[ii] < Z && ii > 0) ii--;                                 /* 1162-1165 */
    ii++;
    for (j = maxclique - 1; j > ii; j--) { best[j] = best[j - 1]; clique[j] = clique[j - 1]; }
    best[ii] = Z;
    clique[ii] = i;
}

/* Cliquer, RepeatResolver.c:1179-1240: the up to maxclique-1 groups of columns [anfang, ende) that correlate best
 * with group a (score > greedy, intersection > mincov/4), best first.  clique: [maxclique+1], clique[0] = a, unused
 * slots and clique[maxclique] = -1; best: [maxclique] scores, best[0] = 100 (1229).  Returns the number of members
 * including a.  The reference leaves the unused slots of its malloc'ed array uninitialised until its trimming loop
 * (1231-1232) overwrites them with -1; this restatement starts from -1 and does not step below slot 1 there. */
int rr_oracle_cliquer(const rr_oracle *o, int anfang, int ende, int mincov, int maxclique, double greedy, int a,
                      int32_t *clique, double *best)
{
    const int sc = o->sc;
    const uint64_t *ga = o->groups + (size_t)a * sc, *ca = o->cover + (size_t)(a / 5) * sc;
    int ii, k, j, n;
    for (j = 0; j <= maxclique; j++) clique[j] = -1;
    for (j = 0; j < maxclique; j++) best[j] = 0.0;                       /* 1198 */
    clique[0] = a;                                                       /* 1197 */
    if (ende > o->N) ende = o->N;
    for (ii = anfang < 0 ? 0 : anfang; ii < ende; ii++)                  /* 1205 */
        for (k = 0; k < 5; k++) {
            const int i = ii * 5 + k;
            const uint64_t *gi = o->groups + (size_t)i * sc, *ci = o->cover + (size_t)ii * sc;
            int schnitt;
            if (i == a) continue;                                        /* 1210 */
            schnitt = rr_isect(gi, ga, sc);                              /* 1213 */
            if (schnitt > mincov / 4) {                                  /* 1215 */
                const double Z = rr_oracle_group_score((unsigned)schnitt, (unsigned)rr_isect(gi, ca, sc),
                                                       (unsigned)rr_isect(ga, ci, sc), (unsigned)rr_isect(ci, ca, sc),
                                                       o->gsize[i], o->gsize[a]);   /* 1217, argument order of 472 */
                if (Z > greedy) rr_best_updater(clique, best, maxclique, i, Z);       /* 1218-1220 */
            }
        }
    best[0] = 100.0;                                                     /* 1229 */
    clique[maxclique] = -1;                                              /* 1230 */
    j = maxclique - 1;
    while (j >= 1 && (best[j] < greedy || clique[j] == clique[j - 1])) { clique[j] = -1; j--; }   /* 1231-1232 */
    for (n = 0; n < maxclique && clique[n] >= 0; n++) {}
    return n;
}

/* ---- Relative_Vars, /root/reference/RepeatResolver.c:2424-2493 ----------------
 * Which groups vary inside one part of the current read partition: restated before any GPU path for it exists,
 * validated against the unmodified RepeatResolver.c behind oracle/ref_relvars_driver.c (tests/test_oracle_relvars.py). */

/* Triple_Schnitt, RepeatResolver.c:150-161 */
static inline int rr_isect3(const uint64_t *a, const uint64_t *b, const uint64_t *c, int sc)
{
    int z, n = 0;
    for (z = 0; z < sc; z++) n += __builtin_popcountll(a[z] & b[z] & c[z]);
    return n;
}

/* Relative_Group_Significance (506-523) on counts, with CumHypGeo_Log (490-504): the two-sided test - the smaller of
 * the lower tail P[X <= schnitt] and the upper tail P[X >= schnitt] of Hypergeom(pop = cov, successes = gr2, draws =
 * gr1), as -log10, capped at 99.  schnitt = |G1 & G2 & U|, gr1 = |G1 & U|, gr2 = |G2 & U|, cov = |U|. */
double rr_oracle_relative_score(unsigned int schnitt, unsigned int gr1, unsigned int gr2, unsigned int cov)
{
    double posP, posQ, Z;
    if (gr1 == 0 || gr2 == 0) return 0.0;                                /* 517 */
    posP = gsl_cdf_hypergeometric_P(schnitt, gr2, cov - gr2, gr1);       /* 492 */
    posQ = gsl_cdf_hypergeometric_Q(schnitt - 1, gr2, cov - gr2, gr1);   /* 493: schnitt - 1 wraps for 0, Q = 0 then */
    if (posP < posQ || schnitt == 0) {                                   /* 495 */
        Z = -1.0 * log10(posP);
        if (isinf(Z) || Z > 99) Z = 99.0;
    } else {
        Z = -1.0 * log10(posQ);                                          /* 501 */
        if (isinf(Z) || Z > 99) Z = 99.0;
    }
    if (isinf(Z) || Z > 99.0) Z = 99.0;                                  /* 521 */
    return Z;
}

/* Relative_Vars, 2424-2493.  unterteilung: [R] part number of every read; maxcorrs: [5N]; vars: [5N+1], receives the
 * selected groups in ascending order followed by -1; returns their number.
 *   selected = MaxCorrs > cutoff (2432) and |U & G| >= mingroup (2449)
 *   kept     = has a partner j >= i + 100 or i >= j + 100 among the selected with score(Gj, Gi, U) > cutoff (2461-2475) */
int rr_oracle_relative_vars(const rr_oracle *o, const int32_t *unterteilung, int u_no, const double *maxcorrs,
                            double cutoff, int mingroup, int32_t *vars)
{
    const int sc = o->sc, G = 5 * o->N;
    uint64_t *U = (uint64_t *)calloc((size_t)sc, sizeof(uint64_t));
    int *sel = (int *)malloc(sizeof(int) * (size_t)(G > 0 ? G : 1));
    int *gu = (int *)malloc(sizeof(int) * (size_t)(G > 0 ? G : 1));
    int i, j, count = 0, cov = 0;
    for (i = 0; i < G; i++) sel[i] = maxcorrs[i] > cutoff ? 1 : 0;                       /* 2430-2434 */
    for (i = 0; i < o->R; i++)
        if (unterteilung[i] == u_no) { U[i / 64] |= (uint64_t)1 << (i % 64); cov++; }    /* 2438 */
    for (i = 0; i < G; i++) {
        gu[i] = 0;
        if (sel[i]) {
            gu[i] = rr_isect(U, o->groups + (size_t)i * sc, sc);
            if (gu[i] < mingroup) sel[i] = 0;                                           /* 2449 */
        }
    }
    for (i = 0; i < G; i++) {
        if (!sel[i]) continue;
        for (j = i + 100; j < G; j++) {                                                 /* 2461 */
            if (!sel[j]) continue;
            {   /* Relative_Group_Significance(Groups[j], Groups[i], U_Group): Group1 = j, Group2 = i (2465) */
                const int s = rr_isect3(o->groups + (size_t)j * sc, o->groups + (size_t)i * sc, U, sc);
                const double Z = rr_oracle_relative_score((unsigned)s, (unsigned)gu[j], (unsigned)gu[i], (unsigned)cov);
                if (Z > cutoff) { sel[i] = 2; sel[j] = 2; }                              /* 2467-2471 */
            }
        }
    }
    for (i = 0; i < G; i++)
        if (sel[i] == 2) vars[count++] = i;                                             /* 2485 */
    vars[count] = -1;                                                                   /* 2483 */
    free(U); free(sel); free(gu);
    return count;
}

/* ---- Kmeans, /root/reference/RepeatResolver.c:2604-2821 ------------------------
 * Splits part u_no of the read partition by the reads' signatures over the selected groups (the output of
 * Relative_Vars).  Restated before any GPU path for it exists; validated against the unmodified RepeatResolver.c behind
 * oracle/ref_kmeans_driver.c (tests/test_oracle_kmeans.py).  The dominant cost is the two read x read similarity
 * sweeps (GrMatch 163-175: sc*64 minus the Hamming distance of two varzahl-bit signatures). */
static inline int rr_match(const uint64_t *a, const uint64_t *b, int scv)
{
    int z, d = 0;
    for (z = 0; z < scv; z++) d += __builtin_popcountll(a[z] ^ b[z]);
    return scv * 64 - d;                                                 /* 174 */
}

/* unterteilung: [R] in/out; vars: varzahl selected group ids.  Returns the number of non-empty clusters (2790-2797);
 * the reads of the part get the new part numbers cluster + max(unterteilung) + 1 (2815-2817). */
int rr_oracle_kmeans(const rr_oracle *o, int32_t *unterteilung, int u_no, const int32_t *vars, int varzahl, int mingroup)
{
    const int scv = varzahl / 64 + 1;                                    /* 2626 */
    int *I = (int *)malloc(sizeof(int) * (size_t)(o->R + 1));
    int anzahl = 0, i, j, k, l, min, aufgeteilt = 0, max_u = 0;
    uint64_t *sig, *cen;
    int *cluster, *size;
    for (i = 0; i < o->R; i++)
        if (unterteilung[i] == u_no) I[anzahl++] = i;                    /* 2616-2623 */
    sig = (uint64_t *)calloc((size_t)(anzahl > 0 ? anzahl : 1) * scv, sizeof(uint64_t));
    cen = (uint64_t *)calloc((size_t)(anzahl > 0 ? anzahl : 1) * scv, sizeof(uint64_t));
    cluster = (int *)calloc((size_t)anzahl + 1, sizeof(int));
    size = (int *)calloc((size_t)anzahl + 1, sizeof(int));
#define RR_HAS(g, r) ((o->groups[(size_t)(g) * o->sc + (r) / 64] >> ((r) % 64)) & 1u)
    for (i = 0; i < anzahl; i++)                                         /* 2634-2642: signature of read I[i] */
        for (j = 0; j < varzahl; j++)
            if (RR_HAS(vars[j], I[i])) sig[(size_t)i * scv + j / 64] |= (uint64_t)1 << (j % 64);
    /* 2659-2706: centroid of read i = majority vote (3 of 5) of the five reads kept by the replace-the-smallest
     * rule below; the slots start as (score 0, read 0) and read i itself takes part */
    for (i = 0; i < anzahl; i++) {
        int bs[5] = {0, 0, 0, 0, 0}, bj[5] = {0, 0, 0, 0, 0};
        for (j = 0; j < anzahl; j++) {
            const int score = rr_match(sig + (size_t)j * scv, sig + (size_t)i * scv, scv);
            for (k = 0; k < 5; k++)                                      /* 2670-2685: exchange sort, ascending */
                for (l = k + 1; l < 5; l++)
                    if (bs[l] < bs[k]) {
                        int t = bs[l]; bs[l] = bs[k]; bs[k] = t;
                        t = bj[l]; bj[l] = bj[k]; bj[k] = t;
                    }
            if (score > bs[0]) { bs[0] = score; bj[0] = j; }             /* 2686-2691 */
        }
        for (j = 0; j < varzahl; j++) {                                  /* 2697-2705 */
            int votes = 0;
            for (k = 0; k < 5; k++) votes += (int)RR_HAS(vars[j], I[bj[k]]);
            if (votes > 2) cen[(size_t)i * scv + j / 64] |= (uint64_t)1 << (j % 64);
        }
    }
    /* 2709-2725: every read joins the centroid (of another read) that matches it best; first best wins */
    for (i = 0; i < anzahl; i++) {
        int best = 0, best_j = 0;
        for (j = 0; j < anzahl; j++) {
            const int score = rr_match(cen + (size_t)j * scv, sig + (size_t)i * scv, scv);
            if (score > best && i != j) { best = score; best_j = j; }
        }
        cluster[i] = best_j;
        size[best_j]++;
    }
    /* 2728-2757: clusters of at most `min` reads are dissolved into clusters of at least `min`, in read order, with
     * the sizes updated as it goes */
    for (min = 2; min < mingroup; min++)
        for (i = 0; i < anzahl; i++)
            if (size[cluster[i]] <= min) {
                int best = 0, best_j = 0;
                for (j = 0; j < anzahl; j++)
                    if (size[j] >= min && cluster[i] != j) {
                        const int score = rr_match(cen + (size_t)j * scv, sig + (size_t)i * scv, scv);
                        if (score > best && i != j) { best = score; best_j = j; }
                    }
                size[cluster[i]]--;
                cluster[i] = best_j;
                size[best_j]++;
            }
#undef RR_HAS
    for (i = 0; i < anzahl; i++)
        if (size[i] > 0) aufgeteilt++;                                   /* 2792-2794 */
    for (i = 0; i < o->R; i++)
        if (unterteilung[i] > max_u) max_u = unterteilung[i];            /* 2814 */
    for (i = 0; i < anzahl; i++) unterteilung[I[i]] = cluster[i] + max_u + 1;   /* 2815 */
    free(I); free(sig); free(cen); free(cluster); free(size);
    return aufgeteilt;
}

/* MaxCorrelation.c:516-532 (MaxCorrsRausschreiben) */
int rr_oracle_write(const char *path, const double *M, int G)
{
    FILE *f = fopen(path, "w");
    int i;
    if (!f) return -1;
    for (i = 0; i < G; i++) fprintf(f, "%f\n", M[i]);
    fclose(f);
    return 0;
}

#ifdef RR_ORACLE_MAIN
/* usage: maxcorr_oracle MSA [-c cov] [-p threads]  -> MaxCorrsOf_<MSA> (same argv rules
 * as MaxCorrelation.c:935-973) */
int main(int argc, char **argv)
{
    int cov = 30, par = 1, i;
    char name[4096];
    rr_oracle *o;
    double *M;
    int32_t *arg;
    int64_t P;
    if (argc < 2) { printf("Usage: ./maxcorr_oracle MSApath <options>\n"); return 0; }
    for (i = 2; i < argc; i++) {
        if (argv[i][0] == '-' && argv[i][1] == 'p' && i + 1 < argc) par = (int)strtol(argv[i + 1], NULL, 10);
        if (argv[i][0] == '-' && argv[i][1] == 'c' && i + 1 < argc) cov = (int)strtol(argv[i + 1], NULL, 10);
    }
    if (par < 1) par = 1;
    o = rr_oracle_load(argv[1]);
    if (!o) { printf("MA is missing.\n"); return 1; }
    M = (double *)calloc((size_t)5 * o->N + 1, sizeof(double));
    arg = (int32_t *)calloc((size_t)5 * o->N + 1, sizeof(int32_t));
    P = rr_oracle_scan(o, cov, par, 0, par, M, arg);
    snprintf(name, sizeof name, "MaxCorrsOf_%s", argv[1]);
    rr_oracle_write(name, M, 5 * o->N);
    printf("pair tests: %lld\n", (long long)P);
    return 0;
}
#endif
