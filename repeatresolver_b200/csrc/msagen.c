/* Synthetic refined-MSA generator (benchmark / test input only; not on the hot path).
 *
 * Produces what the upstream pipeline (DataSimulator.py -> ReadCutter -> InitialAligner
 * -> PW_ReAligner) would hand to MaxCorrelation if the alignment were perfect: an
 * `MSAreal`-style matrix (PW_ReAligner.c:1556-1598: chars "ACGT- ", ' ' outside a read's
 * span).  PW_ReAligner on the config-2 read set takes days and DataSimulator.py is
 * Python 2, so the shapes BASELINE.json names are synthesised here with the reference's
 * own distributions:
 *   - copy families: Tree / Distributed / EquiDistant edit semantics of
 *     DataSimulator.py:29-49, 72-90, 93-115 (edits tracked against template columns);
 *   - reads: length histogram x1000 + U[0,1000), uniform start on copy+2 flanks,
 *     sampled per copy until repeat coverage >= c, clipped to the repeat
 *     (DataSimulator.py:126-160, 222-225);
 *   - PacBio errors: keep 95.2 % / substitute 1.4 % / delete 3.4 %, then geometric
 *     insertions with p = 0.103139 (DataSimulator.py:12-27);
 *   - insertion columns = per-gap maximum run over all reads.
 * Deterministic for a given parameter block (counter-based RNG per read), independent
 * of the thread count.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <stdio.h>
#include <math.h>
#include <pthread.h>
#include "../../include/rr_msagen.h"

typedef struct { uint64_t s; } rng_t;
static inline uint64_t rng_next(rng_t *r)
{
    uint64_t z = (r->s += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
static inline double rng_u01(rng_t *r) { return (double)(rng_next(r) >> 11) * (1.0 / 9007199254740992.0); }
static inline int rng_int(rng_t *r, int n) { return (int)(rng_u01(r) * n); }
static inline rng_t rng_seed(uint64_t seed, uint64_t stream)
{
    rng_t r; r.s = seed * 0xD1342543DE82EF95ull + stream * 0x2545F4914F6CDD1Dull + 0x1234567ull;
    rng_next(&r); rng_next(&r);
    return r;
}

/* DataSimulator.py:122-123 */
static const int kLengthsHisto[40] = {0, 323, 427, 411, 355, 353, 358, 321, 293, 321, 281, 275, 241, 239, 226,
                                      185, 177, 162, 126, 117, 126, 108, 88, 83, 61, 52, 51, 29, 16, 7, 3, 1, 1,
                                      0, 0, 0, 0, 0, 0, 0};

struct rr_msagen {
    rr_msagen_params p;
    int L;              /* template length */
    int ncopies;
    uint8_t *cbase;     /* [ncopies][L]: 0..3 base, 4 = deleted in this copy */
    uint8_t *cins;      /* [ncopies][L]: 0 = none, 1..4 = inserted base+1 after position p */
    int R;              /* reads kept */
    int *rcopy, *ra, *rb; /* per read: copy, template span [a,b) */
    uint8_t *maxrun;    /* [L] insertion columns after template position p */
    int64_t *colstart;  /* [L+1] first column of template position p */
    int N;
    int nthreads;
};

static int sample_length(rng_t *r)
{
    static int total = 0;
    int t, len = -1;
    double rand, prob = 0.0;
    if (!total) for (t = 0; t < 40; t++) total += kLengthsHisto[t];
    rand = rng_u01(r);
    while (prob < rand && len < 39) { len++; prob += (double)kLengthsHisto[len] / (double)total; }
    return len * 1000 + rng_int(r, 1000);
}

/* One SNP-like edit on a copy at template position pos (sub / del / ins in thirds,
 * DataSimulator.py:104-111). */
static void edit_copy(uint8_t *cb, uint8_t *ci, int pos, rng_t *r)
{
    double et = rng_u01(r);
    int rand3 = rng_int(r, 3);
    if (et <= 1.0 / 3.0) {
        if (cb[pos] < 4) cb[pos] = (uint8_t)((cb[pos] + 1 + rand3) & 3);
    } else if (et <= 2.0 / 3.0) {
        cb[pos] = 4;
    } else {
        ci[pos] = (uint8_t)(1 + rng_int(r, 4));
    }
}

static void make_copies(rr_msagen *g)
{
    const int L = g->L, n = g->ncopies;
    rng_t r = rng_seed(g->p.seed, 0xC0FFEEull);
    uint8_t *tmpl = (uint8_t *)malloc(L);
    int i, c, t;
    for (i = 0; i < L; i++) tmpl[i] = (uint8_t)rng_int(&r, 4);
    g->cbase = (uint8_t *)malloc((size_t)n * L);
    g->cins = (uint8_t *)calloc((size_t)n * L, 1);
    if (g->p.type == RR_MSAGEN_TREE) {
        /* DataSimulator.py:93-115: log2(n)+1 doubling levels, each child gets
         * int(diff/2*L) fresh edits on top of its parent. */
        int snps = (int)(g->p.diff / 2.0 * L);
        int levels = (int)(log((double)n) / log(2.0)) + 1;
        int cur = 1, lvl;
        uint8_t *ab = (uint8_t *)malloc((size_t)2 * L), *ai = (uint8_t *)calloc((size_t)2 * L, 1);
        size_t capc = 1;
        memcpy(ab, tmpl, L);
        for (lvl = 0; lvl < levels; lvl++) {
            int next = cur * 2;
            uint8_t *nb = (uint8_t *)malloc((size_t)next * L), *ni = (uint8_t *)malloc((size_t)next * L);
            for (c = 0; c < cur; c++)
                for (t = 0; t < 2; t++) {
                    uint8_t *cb = nb + (size_t)(2 * c + t) * L, *ci = ni + (size_t)(2 * c + t) * L;
                    int e;
                    memcpy(cb, ab + (size_t)c * L, L);
                    memcpy(ci, ai + (size_t)c * L, L);
                    for (e = 0; e < snps; e++) edit_copy(cb, ci, rng_int(&r, L > snps ? L - snps : L), &r);
                }
            free(ab); free(ai);
            ab = nb; ai = ni; cur = next;
            (void)capc;
        }
        for (c = 0; c < n; c++) {
            memcpy(g->cbase + (size_t)c * L, ab + (size_t)(c % cur) * L, L);
            memcpy(g->cins + (size_t)c * L, ai + (size_t)(c % cur) * L, L);
        }
        free(ab); free(ai);
    } else if (g->p.type == RR_MSAGEN_EQUIDISTANT) {
        /* DataSimulator.py:72-90: every copy gets int(diff/2*L) independent edits. */
        int snps = (int)(g->p.diff / 2.0 * L);
        for (c = 0; c < n; c++) {
            uint8_t *cb = g->cbase + (size_t)c * L, *ci = g->cins + (size_t)c * L;
            int e;
            memcpy(cb, tmpl, L);
            for (e = 0; e < snps; e++) edit_copy(cb, ci, rng_int(&r, L), &r);
        }
    } else {
        /* DataSimulator.py:29-49 (Distributed): int(L*diff*3) positions, each edit applied to
         * a random-size random subset of the copies. */
        int snps = (int)(L * g->p.diff * 3), e;
        int *perm = (int *)malloc(sizeof(int) * n);
        for (c = 0; c < n; c++) { memcpy(g->cbase + (size_t)c * L, tmpl, L); perm[c] = c; }
        for (e = 0; e < snps; e++) {
            int pos = 10 + rng_int(&r, L > 20 ? L - 20 : 1), k, x;
            double et;
            int ib = rng_int(&r, 4);
            if (pos >= L) pos = L - 1;
            for (k = n - 1; k > 0; k--) { int j = rng_int(&r, k + 1), tt = perm[k]; perm[k] = perm[j]; perm[j] = tt; }
            k = rng_int(&r, n);
            et = rng_u01(&r);
            for (x = 0; x < k; x++) {
                uint8_t *cb = g->cbase + (size_t)perm[x] * L, *ci = g->cins + (size_t)perm[x] * L;
                if (et <= 1.0 / 3.0) { if (cb[pos] < 4) cb[pos] = (uint8_t)((tmpl[pos] + 1 + k % 3) & 3); }
                else if (et <= 2.0 / 3.0) cb[pos] = 4;
                else ci[pos] = (uint8_t)(1 + ib);
            }
        }
        free(perm);
    }
    free(tmpl);
}

/* DataSimulator.py:126-160 per copy; keeps reads whose repeat overlap >= min_overlap. */
static void sample_reads(rr_msagen *g)
{
    const int L = g->L, flank = g->p.flank, glen = L + 2 * flank;
    int cap = 1024, c;
    g->rcopy = (int *)malloc(sizeof(int) * cap);
    g->ra = (int *)malloc(sizeof(int) * cap);
    g->rb = (int *)malloc(sizeof(int) * cap);
    g->R = 0;
    for (c = 0; c < g->ncopies; c++) {
        rng_t r = rng_seed(g->p.seed, 0x5EAD0000ull + (uint64_t)c);
        double covsum = 0.0;
        long guard = 0;
        while (covsum / (double)L < (double)g->p.coverage && guard++ < 100000000L) {
            int len = sample_length(&r), start, a, b;
            if (len >= glen) len = glen - 1;
            start = rng_int(&r, glen - len);
            a = (start > flank ? start : flank) - flank;
            b = (start + len < flank + L ? start + len : flank + L) - flank;
            covsum += (double)(b - a); /* may be negative for flank-only reads, as in the reference */
            if (b - a < g->p.min_overlap) continue;
            if (g->p.max_reads > 0 && g->R >= g->p.max_reads) continue;
            if (g->R == cap) {
                cap *= 2;
                g->rcopy = (int *)realloc(g->rcopy, sizeof(int) * cap);
                g->ra = (int *)realloc(g->ra, sizeof(int) * cap);
                g->rb = (int *)realloc(g->rb, sizeof(int) * cap);
            }
            g->rcopy[g->R] = c; g->ra[g->R] = a; g->rb[g->R] = b; g->R++;
        }
    }
}

/* Walk one read through the error model.  If out == NULL only the per-gap insertion run
 * maxima are accumulated into maxrun (pass 1); otherwise the row is emitted (pass 2).
 * Both passes consume the identical RNG stream. */
static void emit_read(const rr_msagen *g, int read, uint8_t *maxrun, uint8_t *out, const uint8_t *sym)
{
    const int L = g->L, a = g->ra[read], b = g->rb[read];
    const uint8_t *cb = g->cbase + (size_t)g->rcopy[read] * L, *ci = g->cins + (size_t)g->rcopy[read] * L;
    rng_t r = rng_seed(g->p.seed, 0x7EAD000000ull + (uint64_t)read);
    int p;
    for (p = a; p < b; p++) {
        int code, run = 0, k;
        uint8_t ins[16];
        if (cb[p] == 4) code = 4;
        else {
            double u = rng_u01(&r);
            if (u < 0.837 + 0.115) code = cb[p];
            else if (u < 0.837 + 0.115 + 0.014) code = (cb[p] + 1 + rng_int(&r, 3)) & 3;
            else code = 4;
        }
        if (p + 1 < b) { /* no insertion columns after the last base of the span */
            if (ci[p]) ins[run++] = (uint8_t)(ci[p] - 1);
            if (cb[p] != 4)
                while (run < 15 && rng_u01(&r) < 0.103139) ins[run++] = (uint8_t)rng_int(&r, 4);
        }
        if (!out) {
            if (run > maxrun[p]) maxrun[p] = (uint8_t)run;
        } else {
            int64_t c0 = g->colstart[p];
            int mr = g->maxrun[p];
            out[c0] = sym[code];
            if (p + 1 < b) for (k = 0; k < mr; k++) out[c0 + 1 + k] = k < run ? sym[ins[k]] : sym[4];
        }
    }
}

typedef struct { rr_msagen *g; int t, nt; uint8_t *maxrun; uint8_t *out; size_t stride; const uint8_t *sym; uint8_t fill; } job_t;

static void *pass1_thread(void *x)
{
    job_t *j = (job_t *)x;
    int r;
    for (r = j->t; r < j->g->R; r += j->nt) emit_read(j->g, r, j->maxrun, NULL, NULL);
    return NULL;
}
static void *pass2_thread(void *x)
{
    job_t *j = (job_t *)x;
    int r;
    for (r = j->t; r < j->g->R; r += j->nt) {
        uint8_t *row = j->out + (size_t)r * j->stride;
        int64_t c, c0 = j->g->colstart[j->g->ra[r]], c1 = j->g->colstart[j->g->rb[r] - 1];
        memset(row, j->fill, j->g->N);
        emit_read(j->g, r, NULL, row, j->sym);
        /* leading / trailing gaps of a row become "not covered" (PW_ReAligner.c:459-645) */
        for (c = c0; c <= c1 && row[c] == j->sym[4]; c++) row[c] = j->fill;
        for (c = c1; c >= c0 && row[c] == j->sym[4]; c--) row[c] = j->fill;
        if (j->stride > (size_t)j->g->N) row[j->g->N] = '\n';
    }
    return NULL;
}

rr_msagen *rr_msagen_create(const rr_msagen_params *p)
{
    rr_msagen *g = (rr_msagen *)calloc(1, sizeof(*g));
    int nt, t, i;
    pthread_t th[64];
    job_t jobs[64];
    g->p = *p;
    g->L = p->repeat_len; g->ncopies = p->copies;
    nt = p->threads > 0 ? p->threads : 8;
    if (nt > 64) nt = 64;
    g->nthreads = nt;
    make_copies(g);
    sample_reads(g);
    g->maxrun = (uint8_t *)calloc(g->L + 1, 1);
    for (t = 0; t < nt; t++) {
        jobs[t].g = g; jobs[t].t = t; jobs[t].nt = nt;
        jobs[t].maxrun = (uint8_t *)calloc(g->L + 1, 1);
        pthread_create(&th[t], NULL, pass1_thread, &jobs[t]);
    }
    for (t = 0; t < nt; t++) {
        pthread_join(th[t], NULL);
        for (i = 0; i < g->L; i++) if (jobs[t].maxrun[i] > g->maxrun[i]) g->maxrun[i] = jobs[t].maxrun[i];
        free(jobs[t].maxrun);
    }
    g->colstart = (int64_t *)malloc(sizeof(int64_t) * (g->L + 1));
    g->colstart[0] = 0;
    for (i = 0; i < g->L; i++) g->colstart[i + 1] = g->colstart[i] + 1 + g->maxrun[i];
    /* the last template position carries no insertion columns */
    g->N = (int)(g->colstart[g->L] - g->maxrun[g->L - 1]);
    return g;
}

void rr_msagen_free(rr_msagen *g)
{
    if (!g) return;
    free(g->cbase); free(g->cins); free(g->rcopy); free(g->ra); free(g->rb);
    free(g->maxrun); free(g->colstart); free(g);
}

int rr_msagen_rows(const rr_msagen *g) { return g->R; }
int rr_msagen_cols(const rr_msagen *g) { return g->N; }
const int *rr_msagen_read_copy(const rr_msagen *g) { return g->rcopy; }

static void run_pass2(rr_msagen *g, uint8_t *out, size_t stride, const uint8_t *sym, uint8_t fill)
{
    int nt = g->nthreads, t;
    pthread_t th[64];
    job_t jobs[64];
    for (t = 0; t < nt; t++) {
        jobs[t].g = g; jobs[t].t = t; jobs[t].nt = nt; jobs[t].out = out; jobs[t].stride = stride;
        jobs[t].sym = sym; jobs[t].fill = fill; jobs[t].maxrun = NULL;
        pthread_create(&th[t], NULL, pass2_thread, &jobs[t]);
    }
    for (t = 0; t < nt; t++) pthread_join(th[t], NULL);
}

/* codes[R][N], values 0..5 (the reference's Signatures coding, MaxCorrelation.c:304-329) */
void rr_msagen_fill_codes(rr_msagen *g, uint8_t *codes)
{
    static const uint8_t sym[5] = {0, 1, 2, 3, 4};
    run_pass2(g, codes, (size_t)g->N, sym, 5);
}

/* text[R][N+1]: "ACGT- " rows, each terminated by '\n' (PW_ReAligner.c:1556-1598) */
void rr_msagen_fill_text(rr_msagen *g, char *text)
{
    static const uint8_t sym[5] = {'A', 'C', 'G', 'T', '-'};
    run_pass2(g, (uint8_t *)text, (size_t)g->N + 1, sym, ' ');
}

int rr_msagen_write(rr_msagen *g, const char *path)
{
    size_t bytes = (size_t)g->R * ((size_t)g->N + 1);
    char *text = (char *)malloc(bytes ? bytes : 1);
    FILE *f;
    if (!text) return -2;
    rr_msagen_fill_text(g, text);
    f = fopen(path, "w");
    if (!f) { free(text); return -1; }
    fwrite(text, 1, bytes, f);
    fclose(f);
    free(text);
    return 0;
}

#ifdef RR_MSAGEN_MAIN
int main(int argc, char **argv)
{
    rr_msagen_params p = {RR_MSAGEN_TREE, 100, 40, 30000, 0.01, 1001, 10000, 500, 0, 8};
    const char *out = "MSAreal";
    rr_msagen *g;
    int i;
    for (i = 1; i + 1 < argc; i += 2) {
        if (!strcmp(argv[i], "-t")) p.type = !strcmp(argv[i + 1], "Tree") ? RR_MSAGEN_TREE : !strcmp(argv[i + 1], "Distributed") ? RR_MSAGEN_DISTRIBUTED : RR_MSAGEN_EQUIDISTANT;
        else if (!strcmp(argv[i], "-n")) p.copies = atoi(argv[i + 1]);
        else if (!strcmp(argv[i], "-c")) p.coverage = atoi(argv[i + 1]);
        else if (!strcmp(argv[i], "-l")) p.repeat_len = atoi(argv[i + 1]);
        else if (!strcmp(argv[i], "-d")) p.diff = atof(argv[i + 1]) / 100.0;
        else if (!strcmp(argv[i], "-s")) p.seed = strtoull(argv[i + 1], NULL, 10);
        else if (!strcmp(argv[i], "-m")) p.max_reads = atoi(argv[i + 1]);
        else if (!strcmp(argv[i], "-o")) out = argv[i + 1];
    }
    g = rr_msagen_create(&p);
    fprintf(stderr, "rows %d cols %d\n", rr_msagen_rows(g), rr_msagen_cols(g));
    i = rr_msagen_write(g, out);
    rr_msagen_free(g);
    return i ? 1 : 0;
}
#endif
