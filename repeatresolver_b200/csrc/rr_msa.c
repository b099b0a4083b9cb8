/* rr_msa.c -- host side of the MSA input: the reading half of Einlesen
 * (/root/reference/MaxCorrelation.c:270-335) and the small sequential pieces that stay
 * on the host (first-break sweep over read spans, MaxCorrsOf_* writer).
 *
 * Row rule kept from the reference: the first line fixes siglength = strlen(line)-1
 * (291, the line still carrying its '\n'); a later line is kept iff strlen(line)-1 ==
 * siglength (299), so a last line without '\n' is dropped unless it is one character
 * longer, in which case its last character is cut.  strlen semantics: a NUL inside a
 * line ends it.  Not kept: the static limits Max_Var_Anzahl / Max_Sig_Anzahl (18-19) and
 * fgets' chunking of lines longer than 149 997 characters (287).
 * The characters themselves are classified on the device (rr_pack.cu).
 */
#define _GNU_SOURCE
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <stdarg.h>
#include <errno.h>
#include <fcntl.h>
#include <unistd.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <pthread.h>
#include "rr_host.h"

static __thread char rr_errbuf[512];

void rr_set_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(rr_errbuf, sizeof rr_errbuf, fmt, ap);
    va_end(ap);
}
const char *rr_last_error(void) { return rr_errbuf; }
const char *rr_version(void) { return "repeatresolver_b200 0.1 (sm_100a)"; }

int rr_msa_alloc(int rows, int cols, int codes, rr_msa **out)
{
    rr_msa *m;
    size_t bytes = (size_t)(rows > 0 ? rows : 0) * (size_t)(cols > 0 ? cols : 0);
    if (!out || rows < 0 || cols < 0) { rr_set_error("rr_msa_alloc: bad arguments"); return RR_E_ARG; }
    m = (rr_msa *)calloc(1, sizeof(*m));
    if (!m) return RR_E_NOMEM;
    m->rows = rows; m->cols = cols; m->codes = codes ? 1 : 0;
    m->cells = (uint8_t *)rr_host_alloc(bytes ? bytes : 1, &m->pinned);
    if (!m->cells) { free(m); rr_set_error("out of host memory (%zu bytes)", bytes); return RR_E_NOMEM; }
    *out = m;
    return RR_OK;
}

void rr_msa_free(rr_msa *m)
{
    if (!m) return;
    rr_host_free(m->cells, m->pinned);
    free(m);
}

int rr_msa_rows(const rr_msa *m) { return m ? m->rows : 0; }
int rr_msa_cols(const rr_msa *m) { return m ? m->cols : 0; }
uint8_t *rr_msa_cells(rr_msa *m) { return m ? m->cells : NULL; }

int rr_msa_from_cells(const uint8_t *cells, int rows, int cols, int codes, rr_msa **out)
{
    int rc = rr_msa_alloc(rows, cols, codes, out);
    if (rc) return rc;
    if (rows > 0 && cols > 0) memcpy((*out)->cells, cells, (size_t)rows * cols);
    return RR_OK;
}

/* one line of the text: [p, p+raw) includes the '\n' if there is one; returns strlen() of
 * what fgets would have delivered */
static inline size_t line_strlen(const char *p, size_t raw)
{
    const char *z = (const char *)memchr(p, 0, raw);
    return z ? (size_t)(z - p) : raw;
}

/* copy the kept rows into the cell matrix with a few threads (the copy of a 1-2 GB MSA is memory-bound and
 * dominates the host side once the scan itself takes a fraction of a second) */
typedef struct { const char **rowp; uint8_t *cells; size_t cols; int r0, r1; } copy_job;
static void *copy_rows_thread(void *x)
{
    copy_job *j = (copy_job *)x;
    int r;
    for (r = j->r0; r < j->r1; r++) memcpy(j->cells + (size_t)r * j->cols, j->rowp[r], j->cols);
    return NULL;
}

int rr_msa_from_text(const char *text, size_t nbytes, rr_msa **out)
{
    size_t pos = 0;
    long cols = -1;
    int rows = 0, cap = 0, rc, t, nt;
    const char **rowp = NULL;
    rr_msa *m = NULL;
    pthread_t th[16];
    copy_job jobs[16];
    if (!out || (!text && nbytes)) { rr_set_error("rr_msa_from_text: bad arguments"); return RR_E_ARG; }
    /* pass 1 (sequential, memchr-bound): line boundaries and the keep rule */
    while (pos < nbytes) {
        const char *p = text + pos;
        const char *nl = (const char *)memchr(p, '\n', nbytes - pos);
        size_t raw = nl ? (size_t)(nl - p) + 1 : nbytes - pos;
        size_t sl;
        pos += raw;
        /* strlen semantics need the NUL scan only when the line length could match */
        if (cols >= 0 && (long)raw - 1 < cols) continue;
        sl = line_strlen(p, raw);
        if (cols < 0) cols = (long)sl - 1;                 /* 291 */
        if ((long)sl - 1 != cols) continue;                /* 299 */
        if (rows == cap) {
            const char **np;
            cap = cap ? cap * 2 : 1024;
            np = (const char **)realloc((void *)rowp, sizeof(char *) * (size_t)cap);
            if (!np) { free((void *)rowp); rr_set_error("out of host memory"); return RR_E_NOMEM; }
            rowp = np;
        }
        rowp[rows++] = p;
    }
    if (cols < 0) cols = 0;
    if (cols > 0x7fffffffL / 8) { free((void *)rowp); rr_set_error("MSA too wide (%ld columns)", cols); return RR_E_ARG; }
    rc = rr_msa_alloc(rows, (int)cols, 0, &m);
    if (rc) { free((void *)rowp); return rc; }
    /* pass 2: parallel copy */
    nt = (size_t)rows * (size_t)cols > (1u << 24) ? 8 : 1;
    if (cols > 0)
        for (t = 0; t < nt; t++) {
            jobs[t].rowp = rowp; jobs[t].cells = m->cells; jobs[t].cols = (size_t)cols;
            jobs[t].r0 = (int)((long long)rows * t / nt); jobs[t].r1 = (int)((long long)rows * (t + 1) / nt);
            if (nt == 1) copy_rows_thread(&jobs[t]);
            else if (pthread_create(&th[t], NULL, copy_rows_thread, &jobs[t]) != 0) { copy_rows_thread(&jobs[t]); th[t] = 0; }
        }
    if (cols > 0 && nt > 1)
        for (t = 0; t < nt; t++)
            if (th[t]) pthread_join(th[t], NULL);
    free((void *)rowp);
    *out = m;
    return RR_OK;
}

int rr_msa_read(const char *path, rr_msa **out)
{
    int fd, rc;
    struct stat st;
    void *map;
    if (!path || !out) { rr_set_error("rr_msa_read: bad arguments"); return RR_E_ARG; }
    fd = open(path, O_RDONLY);
    if (fd < 0) { rr_set_error("MA is missing. (%s: %s)", path, strerror(errno)); return RR_E_IO; }
    if (fstat(fd, &st) != 0) { close(fd); rr_set_error("fstat %s: %s", path, strerror(errno)); return RR_E_IO; }
    if (st.st_size == 0) { close(fd); return rr_msa_alloc(0, 0, 0, out); }
    map = mmap(NULL, (size_t)st.st_size, PROT_READ, MAP_PRIVATE, fd, 0);
    close(fd);
    if (map == MAP_FAILED) { rr_set_error("mmap %s: %s", path, strerror(errno)); return RR_E_IO; }
    madvise(map, (size_t)st.st_size, MADV_SEQUENTIAL);
    rc = rr_msa_from_text((const char *)map, (size_t)st.st_size, out);
    munmap(map, (size_t)st.st_size);
    return rc;
}

/* MaxCorrsRausschreiben (516-532): one "%f\n" per group */
int rr_maxcorr_write(const char *path, const double *maxcorr, int64_t count)
{
    FILE *f;
    int64_t i;
    if (!path || (!maxcorr && count)) return RR_E_ARG;
    f = fopen(path, "w");
    if (!f) { rr_set_error("cannot write %s: %s", path, strerror(errno)); return RR_E_IO; }
    for (i = 0; i < count; i++) fprintf(f, "%f\n", maxcorr[i]);
    if (fclose(f) != 0) { rr_set_error("write %s: %s", path, strerror(errno)); return RR_E_IO; }
    return RR_OK;
}

int rr_argmax_write(const char *path, const int32_t *argmax, int64_t count)
{
    FILE *f;
    int64_t i;
    if (!path || (!argmax && count)) return RR_E_ARG;
    f = fopen(path, "w");
    if (!f) { rr_set_error("cannot write %s: %s", path, strerror(errno)); return RR_E_IO; }
    for (i = 0; i < count; i++) fprintf(f, "%d\n", argmax[i]);
    if (fclose(f) != 0) return RR_E_IO;
    return RR_OK;
}

/* First-break columns for rows that are single spans.  With contiguous spans the shared
 * coverage |C[ii] & C[jj]| = #{r : start_r <= ii, end_r >= jj} never increases with jj, so
 * the reference's "stop at the first jj with shared coverage < mincov" (807-810) is
 *     break(ii) = max(ii+20, e*(ii)+1),
 * e*(ii) = the mincov-th largest end among rows with start <= ii (-1 if fewer).  One sweep
 * with a min-heap of the mincov largest ends. */
static void heap_sift_down(int32_t *h, int n, int i)
{
    for (;;) {
        int l = 2 * i + 1, r = l + 1, s = i;
        int32_t t;
        if (l < n && h[l] < h[s]) s = l;
        if (r < n && h[r] < h[s]) s = r;
        if (s == i) return;
        t = h[i]; h[i] = h[s]; h[s] = t; i = s;
    }
}
static void heap_sift_up(int32_t *h, int i)
{
    while (i > 0) {
        int p = (i - 1) / 2;
        int32_t t;
        if (h[p] <= h[i]) return;
        t = h[i]; h[i] = h[p]; h[p] = t; i = p;
    }
}
static int cmp_span_start(const void *a, const void *b)
{
    const int32_t *x = (const int32_t *)a, *y = (const int32_t *)b;
    if (x[0] != y[0]) return x[0] < y[0] ? -1 : 1;
    return (x[1] > y[1]) - (x[1] < y[1]);
}

int rr_breakcols_from_spans(const int32_t *start, const int32_t *end, int rows, int cols, int mincov,
                            int32_t *breakcol)
{
    int32_t *sp, *heap;
    int r, ii, next = 0, hn = 0, nsp = 0;
    if (rows < 0 || cols < 0) { rr_set_error("rr_breakcols_from_spans: bad arguments"); return RR_E_ARG; }
    if (cols == 0) return RR_OK;
    if (!breakcol || (rows && (!start || !end))) { rr_set_error("rr_breakcols_from_spans: null pointer"); return RR_E_ARG; }
    if (mincov <= 0) { /* shared coverage < mincov never holds */
        for (ii = 0; ii < cols; ii++) breakcol[ii] = cols > ii + 20 ? cols : ii + 20;
        return RR_OK;
    }
    sp = (int32_t *)malloc(sizeof(int32_t) * 2 * (size_t)(rows ? rows : 1));
    heap = (int32_t *)malloc(sizeof(int32_t) * (size_t)mincov);
    if (!sp || !heap) { free(sp); free(heap); return RR_E_NOMEM; }
    for (r = 0; r < rows; r++)
        if (end[r] >= start[r] && start[r] >= 0) { sp[2 * nsp] = start[r]; sp[2 * nsp + 1] = end[r]; nsp++; }
    qsort(sp, (size_t)nsp, 2 * sizeof(int32_t), cmp_span_start);
    for (ii = 0; ii < cols; ii++) {
        int32_t estar, b;
        while (next < nsp && sp[2 * next] <= ii) {
            int32_t e = sp[2 * next + 1];
            if (hn < mincov) { heap[hn] = e; heap_sift_up(heap, hn); hn++; }
            else if (e > heap[0]) { heap[0] = e; heap_sift_down(heap, hn, 0); }
            next++;
        }
        estar = hn == mincov ? heap[0] : -1;
        b = estar + 1;
        if (b < ii + 20) b = ii + 20;
        breakcol[ii] = b;
    }
    free(sp); free(heap);
    return RR_OK;
}
