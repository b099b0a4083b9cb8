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
#include <time.h>
#include "rr_host.h"

static __thread char rr_errbuf[512];

void rr_set_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(rr_errbuf, sizeof rr_errbuf, fmt, ap);
    va_end(ap);
}
void rr_trace_mark(const char *tag)
{
    static int on = -1;
    static double t0 = 0.0, last = 0.0;
    struct timespec ts;
    double now;
    if (on < 0) on = getenv("RR_TRACE") ? 1 : 0;
    if (!on) return;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    now = ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6;
    if (t0 == 0.0) t0 = last = now;
    fprintf(stderr, "[rr trace] %-28s +%9.3f ms  (t = %9.3f ms)\n", tag, now - last, now - t0);
    last = now;
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
    if (m->cells) rr_host_free(m->cells, m->pinned);
    if (m->map) munmap((void *)m->map, m->map_len);
    free(m->rowoff);
    free(m);
}

int rr_msa_rows(const rr_msa *m) { return m ? m->rows : 0; }
int rr_msa_cols(const rr_msa *m) { return m ? m->cols : 0; }

/* copy rows [r0, r1) into a [.][cols] matrix with a few threads (memory-bound) */
typedef struct { const rr_msa *m; uint8_t *dst; int r0, r1; } copy_job;
static void *copy_rows_thread(void *x)
{
    copy_job *j = (copy_job *)x;
    int r;
    for (r = j->r0; r < j->r1; r++) memcpy(j->dst + (size_t)(r - j->r0) * j->m->cols, rr_msa_row(j->m, r), (size_t)j->m->cols);
    return NULL;
}
void rr_msa_gather_rows(const rr_msa *m, int r0, int r1, uint8_t *dst, int threads)
{
    pthread_t th[16];
    copy_job jobs[16];
    int t, nt = threads;
    if (r1 <= r0 || m->cols <= 0) return;
    if (nt > 16) nt = 16;
    if ((size_t)(r1 - r0) * (size_t)m->cols < (1u << 22) || nt < 1) nt = 1;
    for (t = 0; t < nt; t++) {
        jobs[t].m = m;
        jobs[t].r0 = r0 + (int)((long long)(r1 - r0) * t / nt);
        jobs[t].r1 = r0 + (int)((long long)(r1 - r0) * (t + 1) / nt);
        jobs[t].dst = dst + (size_t)(jobs[t].r0 - r0) * m->cols;
        th[t] = 0;
        if (nt == 1 || pthread_create(&th[t], NULL, copy_rows_thread, &jobs[t]) != 0) { copy_rows_thread(&jobs[t]); th[t] = 0; }
    }
    for (t = 0; t < nt; t++)
        if (th[t]) pthread_join(th[t], NULL);
}

uint8_t *rr_msa_cells(rr_msa *m)
{
    if (!m) return NULL;
    if (!m->cells && m->map) { /* text-backed: materialise once (pageable memory: rr_pack stages it through its ring) */
        size_t bytes = (size_t)m->rows * (size_t)m->cols;
        uint8_t *c = (uint8_t *)malloc(bytes ? bytes : 1);
        if (!c) { rr_set_error("out of host memory (%zu bytes)", bytes); return NULL; }
        rr_msa_gather_rows(m, 0, m->rows, c, 8);
        m->cells = c; m->pinned = 0;
        munmap((void *)m->map, m->map_len);
        m->map = NULL; m->map_len = 0;
        free(m->rowoff); m->rowoff = NULL;
    }
    return m->cells;
}

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

/* Line index of a text MSA with the reference's row rules (291, 299): the first line fixes cols = strlen - 1, every
 * line of that length is a row.  Two parallel passes over byte ranges / line ranges (a 1-2 GB MSA is memchr-bound):
 * newline positions, then the keep test (the NUL scan of strlen semantics only where the raw length could match). */
typedef struct {
    const char *text; size_t lo, hi;      /* pass 1: byte range */
    int64_t *starts; size_t n, cap; int oom;
    const int64_t *line; size_t l0, l1, nbytes; long cols; uint8_t *keep;  /* pass 2: line range */
} scan_job;
static void *scan_newlines_thread(void *x)
{
    scan_job *j = (scan_job *)x;
    size_t pos = j->lo;
    while (pos < j->hi) {
        const char *nl = (const char *)memchr(j->text + pos, '\n', j->hi - pos);
        if (!nl) break;
        pos = (size_t)(nl - j->text) + 1;
        if (j->n == j->cap) {
            size_t nc = j->cap ? j->cap * 2 : 4096;
            int64_t *np = (int64_t *)realloc(j->starts, nc * sizeof(int64_t));
            if (!np) { j->oom = 1; return NULL; }
            j->starts = np; j->cap = nc;
        }
        j->starts[j->n++] = (int64_t)pos;  /* a line starts after every '\n' */
    }
    return NULL;
}
static void *scan_keep_thread(void *x)
{
    scan_job *j = (scan_job *)x;
    size_t l;
    for (l = j->l0; l < j->l1; l++) {
        size_t raw = (size_t)(j->line[l + 1] - j->line[l]);
        j->keep[l] = (long)raw - 1 >= j->cols && (long)line_strlen(j->text + j->line[l], raw) - 1 == j->cols;
    }
    return NULL;
}
static int index_lines(const char *text, size_t nbytes, int *rows_out, long *cols_out, int64_t **rowoff_out)
{
    enum { MAXT = 16 };
    pthread_t th[MAXT];
    scan_job jobs[MAXT];
    int t, nt = nbytes > (1u << 24) ? 8 : 1, rows = 0;
    size_t nlines = 1, l, k;
    int64_t *line, *rowoff;
    uint8_t *keep;
    long cols;
    memset(jobs, 0, sizeof jobs);
    for (t = 0; t < nt; t++) {
        jobs[t].text = text; jobs[t].lo = nbytes / nt * t; jobs[t].hi = t == nt - 1 ? nbytes : nbytes / nt * (t + 1);
        th[t] = 0;
        if (nt == 1 || pthread_create(&th[t], NULL, scan_newlines_thread, &jobs[t]) != 0) { scan_newlines_thread(&jobs[t]); th[t] = 0; }
    }
    for (t = 0; t < nt; t++) {
        if (th[t]) pthread_join(th[t], NULL);
        nlines += jobs[t].n;
    }
    line = (int64_t *)malloc(sizeof(int64_t) * (nlines + 1));
    for (t = 0, k = 1; t < nt; t++) {
        if (line && !jobs[t].oom) { memcpy(line + k, jobs[t].starts, sizeof(int64_t) * jobs[t].n); k += jobs[t].n; }
        else if (line) { free(line); line = NULL; }
        free(jobs[t].starts);
    }
    if (!line) { rr_set_error("out of host memory"); return RR_E_NOMEM; }
    line[0] = 0;
    if ((size_t)line[nlines - 1] == nbytes) nlines--;  /* the text ends with '\n': no line after it */
    line[nlines] = (int64_t)nbytes;
    cols = nlines ? (long)line_strlen(text, (size_t)(line[1] - line[0])) - 1 : -1;   /* 291 */
    keep = (uint8_t *)malloc(nlines ? nlines : 1);
    if (!keep) { free(line); rr_set_error("out of host memory"); return RR_E_NOMEM; }
    for (t = 0; t < nt; t++) {
        jobs[t].line = line; jobs[t].l0 = nlines * t / nt; jobs[t].l1 = nlines * (t + 1) / nt;
        jobs[t].cols = cols; jobs[t].keep = keep; jobs[t].nbytes = nbytes;
        th[t] = 0;
        if (nt == 1 || pthread_create(&th[t], NULL, scan_keep_thread, &jobs[t]) != 0) { scan_keep_thread(&jobs[t]); th[t] = 0; }
    }
    for (t = 0; t < nt; t++)
        if (th[t]) pthread_join(th[t], NULL);
    for (l = 0; l < nlines; l++) rows += keep[l];
    rowoff = (int64_t *)malloc(sizeof(int64_t) * (size_t)(rows ? rows : 1));
    if (!rowoff) { free(line); free(keep); rr_set_error("out of host memory"); return RR_E_NOMEM; }
    for (l = 0, k = 0; l < nlines; l++)
        if (keep[l]) rowoff[k++] = line[l];                /* 299 */
    free(line); free(keep);
    *rows_out = rows; *cols_out = cols < 0 ? 0 : cols; *rowoff_out = rowoff;
    return RR_OK;
}

/* text in caller-owned memory: the rows are copied out (the text may go away) */
int rr_msa_from_text(const char *text, size_t nbytes, rr_msa **out)
{
    long cols = 0;
    int rows = 0, rc;
    int64_t *rowoff = NULL;
    rr_msa *m = NULL, view;
    if (!out || (!text && nbytes)) { rr_set_error("rr_msa_from_text: bad arguments"); return RR_E_ARG; }
    if ((rc = index_lines(text, nbytes, &rows, &cols, &rowoff))) return rc;
    if (cols > 0x7fffffffL / 8) { free(rowoff); rr_set_error("MSA too wide (%ld columns)", cols); return RR_E_ARG; }
    rc = rr_msa_alloc(rows, (int)cols, 0, &m);
    if (rc) { free(rowoff); return rc; }
    memset(&view, 0, sizeof view);
    view.rows = rows; view.cols = (int)cols; view.map = text; view.rowoff = rowoff;
    rr_msa_gather_rows(&view, 0, rows, m->cells, 8);
    free(rowoff);
    *out = m;
    return RR_OK;
}

/* Einlesen's file handling (270-335): the file stays mapped, the rows are indexed, nothing is copied */
int rr_msa_read(const char *path, rr_msa **out)
{
    int fd, rc, rows = 0;
    long cols = 0;
    struct stat st;
    void *map;
    int64_t *rowoff = NULL;
    rr_msa *m;
    if (!path || !out) { rr_set_error("rr_msa_read: bad arguments"); return RR_E_ARG; }
    rr_trace_mark("read: start");
    fd = open(path, O_RDONLY);
    if (fd < 0) { rr_set_error("MA is missing. (%s: %s)", path, strerror(errno)); return RR_E_IO; }
    if (fstat(fd, &st) != 0) { close(fd); rr_set_error("fstat %s: %s", path, strerror(errno)); return RR_E_IO; }
    if (st.st_size == 0) { close(fd); return rr_msa_alloc(0, 0, 0, out); }
    map = mmap(NULL, (size_t)st.st_size, PROT_READ, MAP_PRIVATE, fd, 0);
    close(fd);
    if (map == MAP_FAILED) { rr_set_error("mmap %s: %s", path, strerror(errno)); return RR_E_IO; }
    if ((size_t)st.st_size > (1u << 26)) rr_cuda_warmup_begin();  /* the context comes up while the text is indexed */
    rc = index_lines((const char *)map, (size_t)st.st_size, &rows, &cols, &rowoff);
    rr_trace_mark("read: lines indexed");
    if (!rc && cols > 0x7fffffffL / 8) { rc = RR_E_ARG; rr_set_error("MSA too wide (%ld columns)", cols); }
    m = rc ? NULL : (rr_msa *)calloc(1, sizeof(*m));
    if (!m) {
        if (!rc) { rc = RR_E_NOMEM; rr_set_error("out of host memory"); }
        free(rowoff); munmap(map, (size_t)st.st_size);
        return rc;
    }
    m->rows = rows; m->cols = (int)cols; m->map = (const char *)map; m->map_len = (size_t)st.st_size; m->rowoff = rowoff;
    *out = m;
    return RR_OK;
}

/* Einlesen of RepeatResolver.c (293-429), the reader in front of the rows 8f-2 .. 8f-4: the columns von .. bis of the reads
 * that carry a symbol - anything but ' ' - at BOTH ends of the window (330); bis is lowered to the last column of a shorter
 * line when one is met and stays lowered for the lines that follow (328), and the window that is finally kept is
 * von .. the lowered bis (398).  ausgelassen[line] = 1 for a read that was kept, -1 for one left out (332, 367), as
 * UnterteilungsKomplettierung needs it.  A last line without '\n' is an error (326: the reference exits), and so is a line too
 * short to hold column von, or von > bis (the reference reads stale buffer contents there).  The cells are kept as raw
 * characters (classified on the device like every other MSA). */
int rr_msa_read_window(const char *path, int von, int bis, rr_msa **out, int8_t *ausgelassen, int64_t capacity, int64_t *n_lines)
{
    FILE *f;
    char *text = NULL;
    size_t len = 0, cap = 0, pos, nl = 0, kept = 0, k;
    unsigned char *keep = NULL;
    size_t *lstart = NULL;
    long fsize;
    int rc = RR_OK, cols;
    rr_msa *m = NULL;
    if (!path || !out || !n_lines || von < 0 || bis < von || capacity < 0 || (capacity && !ausgelassen)) {
        rr_set_error("rr_msa_read_window: bad arguments");
        return RR_E_ARG;
    }
    *out = NULL; *n_lines = 0;
    f = fopen(path, "rb");
    if (!f) { rr_set_error("MA is missing. (%s: %s)", path, strerror(errno)); return RR_E_IO; }
    if (fseek(f, 0, SEEK_END) != 0 || (fsize = ftell(f)) < 0 || fseek(f, 0, SEEK_SET) != 0) { fclose(f); rr_set_error("cannot size %s", path); return RR_E_IO; }
    cap = (size_t)fsize;
    text = (char *)malloc(cap ? cap : 1);
    if (!text) { fclose(f); rr_set_error("out of host memory (%zu bytes)", cap); return RR_E_NOMEM; }
    len = fread(text, 1, cap, f);
    fclose(f);
    for (pos = 0; pos < len; pos++) nl += text[pos] == '\n';
    if (len && text[len - 1] != '\n') { free(text); rr_set_error("%s: the last line has no newline", path); return RR_E_IO; }
    keep = (unsigned char *)calloc(nl ? nl : 1, 1);
    lstart = (size_t *)malloc(sizeof(size_t) * (nl ? nl : 1));
    if (!keep || !lstart) { rc = RR_E_NOMEM; rr_set_error("out of host memory"); goto done; }
    for (pos = 0, k = 0; pos < len; k++) {                               /* one pass in reading order: bis only ever goes down */
        const char *line = text + pos;
        const char *e = (const char *)memchr(line, '\n', len - pos);
        const size_t l = (size_t)(e - line);                             /* without the newline (327) */
        if ((long)l - 1 < (long)bis) bis = (int)l - 1;                   /* 328 */
        if (bis < von) { rc = RR_E_IO; rr_set_error("%s: line %zu is too short for column %d", path, k + 1, von); goto done; }
        lstart[k] = pos;
        keep[k] = line[von] != ' ' && line[bis] != ' ';                  /* 330 */
        kept += keep[k];
        pos += l + 1;
    }
    cols = nl ? bis + 1 - von : 0;                                       /* 398 */
    if (kept > 0x7fffffffu) { rc = RR_E_ARG; rr_set_error("too many reads"); goto done; }
    if ((rc = rr_msa_alloc((int)kept, cols, 0, &m))) goto done;
    for (k = 0, pos = 0; k < nl; k++)
        if (keep[k]) { memcpy(m->cells + pos * (size_t)cols, text + lstart[k] + von, (size_t)cols); pos++; }
    for (k = 0; k < nl && (int64_t)k < capacity; k++) ausgelassen[k] = keep[k] ? 1 : -1;
    *n_lines = (int64_t)nl;
    *out = m;
    if (ausgelassen && (int64_t)nl > capacity) { rc = RR_E_ARG; rr_set_error("rr_msa_read_window: %zu lines, room for %lld marks", nl, (long long)capacity); rr_msa_free(m); *out = NULL; }
done:
    free(text); free(keep); free(lstart);
    return rc;
}

/* MaxCorrsRausschreiben (516-532): one "%f\n" per group */
/* "%f\n" of one double, exactly as printf rounds it (the exact binary value to six decimals, ties to even), without printf:
 * v = m * 2^e with a 53-bit m, so m * 10^6 fits 73 bits and one shift with an exact remainder test gives the digits.
 * Returns the length written to out (at most 25 bytes), 0 for values that do not fit the integer path (|v| >= 2^52, inf,
 * nan): the caller prints those with fprintf.  The text is MaxCorrsRausschreiben's (MaxCorrelation.c:527-530); 5N lines of it were 50 ms of fprintf. */
static size_t fmt_f6_line(double v, char *out)
{
    uint64_t bits, m, ip, fp;
    unsigned __int128 P, q;
    int ex, sh, n = 0, k;
    char tmp[24];
    memcpy(&bits, &v, 8);
    ex = (int)((bits >> 52) & 0x7ff);
    m = bits & 0xfffffffffffffull;
    if (ex >= 1075) return 0;                                             /* |v| >= 2^52, inf, nan */
    if (ex) m |= 1ull << 52; else ex = 1;                                 /* subnormals: 0.frac * 2^-1022 */
    sh = 1075 - ex;                                                       /* v = m * 2^-sh, sh in 1 .. 1074 */
    P = (unsigned __int128)m * 1000000u;
    if (sh >= 74) q = 0;                                                  /* P < 2^73 <= half of 2^sh */
    else {
        const unsigned __int128 one = 1, rem = P & ((one << sh) - 1), half = one << (sh - 1);
        q = P >> sh;
        if (rem > half || (rem == half && (q & 1))) q++;
    }
    ip = (uint64_t)(q / 1000000u);
    fp = (uint64_t)(q % 1000000u);
    if (bits >> 63) out[n++] = '-';
    k = 0;
    do { tmp[k++] = (char)('0' + ip % 10); ip /= 10; } while (ip);
    while (k) out[n++] = tmp[--k];
    out[n++] = '.';
    for (k = 5; k >= 0; k--) { out[n + k] = (char)('0' + fp % 10); fp /= 10; }
    n += 6;
    out[n++] = '\n';
    return (size_t)n;
}

static size_t fmt_int_line(int32_t v, char *out)
{
    char tmp[12];
    int n = 0, k = 0;
    uint32_t u = v < 0 ? 0u - (uint32_t)v : (uint32_t)v;
    if (v < 0) out[n++] = '-';
    do { tmp[k++] = (char)('0' + u % 10); u /= 10; } while (u);
    while (k) out[n++] = tmp[--k];
    out[n++] = '\n';
    return (size_t)n;
}

/* lines formatted into a block of memory, one fwrite per block */
static int write_lines(const char *path, const double *d, const int32_t *a, int64_t count)
{
    enum { BLOCK = 1 << 15 };
    FILE *f = fopen(path, "w");
    char *buf;
    int64_t i = 0;
    if (!f) { rr_set_error("cannot write %s: %s", path, strerror(errno)); return RR_E_IO; }
    buf = (char *)malloc((size_t)BLOCK * 32);
    if (!buf) { fclose(f); rr_set_error("out of host memory"); return RR_E_NOMEM; }
    while (i < count) {
        const int64_t n = count - i < BLOCK ? count - i : BLOCK;
        size_t len = 0;
        int64_t k;
        if (d)
            for (k = 0; k < n; k++) {
                const size_t l = fmt_f6_line(d[i + k], buf + len);
                if (l) { len += l; continue; }
                if (fwrite(buf, 1, len, f) != len) len = (size_t)-1;      /* what precedes the odd value, then the value */
                else { len = 0; fprintf(f, "%f\n", d[i + k]); }
                if (len) break;
            }
        else
            for (k = 0; k < n; k++) len += fmt_int_line(a[i + k], buf + len);
        if (len == (size_t)-1 || fwrite(buf, 1, len, f) != len) {
            free(buf); fclose(f); rr_set_error("write %s: %s", path, strerror(errno)); return RR_E_IO;
        }
        i += n;
    }
    free(buf);
    if (fclose(f) != 0) { rr_set_error("write %s: %s", path, strerror(errno)); return RR_E_IO; }
    return RR_OK;
}

int rr_maxcorr_write(const char *path, const double *maxcorr, int64_t count)
{
    if (!path || (!maxcorr && count)) return RR_E_ARG;
    return write_lines(path, maxcorr, NULL, count);
}

int rr_argmax_write(const char *path, const int32_t *argmax, int64_t count)
{
    if (!path || (!argmax && count)) return RR_E_ARG;
    return write_lines(path, NULL, argmax, count);
}

/* ---- binary side format of the result ("the MaxCorrsOf_* binary side-format to skip text") ----
 * MaxCorrsBinOf_<MSA>, little endian:  8 bytes magic "RRMAXC01", int64 count (= 5 * siglength), int64 flags (bit 0: the
 * arg-max partners follow), count doubles, then count int32.  The text file stays the contract with the unmodified
 * downstream programs; this one spares a consumer that links the library the 5N sscanf calls of MaxCorrsEinlesen
 * (RepeatResolver.c:609-646) and keeps the values at full precision. */
static const char RR_BIN_MAGIC[8] = {'R', 'R', 'M', 'A', 'X', 'C', '0', '1'};

int rr_maxcorr_write_bin(const char *path, const double *maxcorr, const int32_t *argmax, int64_t count)
{
    FILE *f;
    int64_t head[2];
    if (!path || count < 0 || (!maxcorr && count)) { rr_set_error("rr_maxcorr_write_bin: bad arguments"); return RR_E_ARG; }
    f = fopen(path, "wb");
    if (!f) { rr_set_error("cannot write %s: %s", path, strerror(errno)); return RR_E_IO; }
    head[0] = count; head[1] = argmax ? 1 : 0;
    if (fwrite(RR_BIN_MAGIC, 1, 8, f) != 8 || fwrite(head, sizeof(int64_t), 2, f) != 2 ||
        (count && fwrite(maxcorr, sizeof(double), (size_t)count, f) != (size_t)count) ||
        (count && argmax && fwrite(argmax, sizeof(int32_t), (size_t)count, f) != (size_t)count)) {
        rr_set_error("write %s: %s", path, strerror(errno));
        fclose(f);
        return RR_E_IO;
    }
    if (fclose(f) != 0) { rr_set_error("write %s: %s", path, strerror(errno)); return RR_E_IO; }
    return RR_OK;
}

/* what a "%f" line of MaxCorrsOf_* gives back to sscanf("%lf") (516-532 / 632) */
static double rr_through_text(double v)
{
    char buf[400];
    snprintf(buf, sizeof buf, "%f", v);
    return strtod(buf, NULL);
}

/* The window MaxCorrsEinlesen(inputfile, von, bis) keeps: the groups of columns von..bis inclusive (631: i/5 >= von &&
 * i/5 <= bis), clipped to the file.  maxcorr_out / argmax_out: [5 * (bis - von + 1)] or NULL; *n_out = values delivered.
 * as_text != 0 rounds every value the way the text file would ("%f", six decimals), so that a consumer makes the very
 * decisions it would make on MaxCorrsOf_*.  argmax_out without stored partners is filled with -1. */
int rr_maxcorr_read_bin(const char *path, int von, int bis, int as_text, double *maxcorr_out, int32_t *argmax_out, int64_t *n_out)
{
    FILE *f;
    char magic[8];
    int64_t head[2], lo, hi, n, i;
    if (!path || !n_out) { rr_set_error("rr_maxcorr_read_bin: bad arguments"); return RR_E_ARG; }
    *n_out = 0;
    f = fopen(path, "rb");
    if (!f) { rr_set_error("cannot open %s: %s", path, strerror(errno)); return RR_E_IO; }
    if (fread(magic, 1, 8, f) != 8 || memcmp(magic, RR_BIN_MAGIC, 8) != 0 || fread(head, sizeof(int64_t), 2, f) != 2 || head[0] < 0) {
        rr_set_error("%s is not a MaxCorrsBinOf_ file", path);
        fclose(f);
        return RR_E_IO;
    }
    lo = (int64_t)5 * (von < 0 ? 0 : von);
    hi = bis < 0 ? 0 : (int64_t)5 * ((int64_t)bis + 1);
    if (hi > head[0]) hi = head[0];
    n = hi > lo ? hi - lo : 0;
    if (n && maxcorr_out) {
        if (fseek(f, (long)(24 + lo * (int64_t)sizeof(double)), SEEK_SET) != 0 || fread(maxcorr_out, sizeof(double), (size_t)n, f) != (size_t)n) {
            rr_set_error("%s is truncated", path);
            fclose(f);
            return RR_E_IO;
        }
        if (as_text)
            for (i = 0; i < n; i++) maxcorr_out[i] = rr_through_text(maxcorr_out[i]);
    }
    if (n && argmax_out) {
        if (!(head[1] & 1)) {
            for (i = 0; i < n; i++) argmax_out[i] = -1;
        } else if (fseek(f, (long)(24 + head[0] * (int64_t)sizeof(double) + lo * (int64_t)sizeof(int32_t)), SEEK_SET) != 0 ||
                   fread(argmax_out, sizeof(int32_t), (size_t)n, f) != (size_t)n) {
            rr_set_error("%s is truncated", path);
            fclose(f);
            return RR_E_IO;
        }
    }
    fclose(f);
    *n_out = n;
    return RR_OK;
}

/* MaxCorrsEinlesen (RepeatResolver.c:609-646) on the text file: the values of the lines i with von <= i/5 <= bis, in file
 * order.  A "line" is what fgets(buffer, 100, file) delivers - up to 99 characters, so a longer line counts as several (623) -
 * and its value what sscanf("%lf") reads at its start: leading white space skipped, the longest valid number, strtod's
 * correctly rounded result.  Where the reference would leave the slot of a line without a number uninitialised, 0.0 is
 * stored.  maxcorr_out [capacity] may be NULL to count; *n_out = values in the window, also beyond the capacity (RR_E_ARG
 * then; the reference writes past its siglength * 5 doubles in that case). */
int rr_maxcorr_read_text(const char *path, int von, int bis, double *maxcorr_out, int64_t capacity, int64_t *n_out)
{
    FILE *f;
    char buffer[101];
    int64_t i = 0, j = 0;
    if (!path || !n_out || capacity < 0 || (capacity && !maxcorr_out)) { rr_set_error("rr_maxcorr_read_text: bad arguments"); return RR_E_ARG; }
    *n_out = 0;
    f = fopen(path, "r");
    if (!f) { rr_set_error("cannot open %s: %s", path, strerror(errno)); return RR_E_IO; }   /* the reference returns NULL (619) */
    while (fgets(buffer, 100, f) != NULL) {
        if (i / 5 >= von && i / 5 <= bis) {
            if (j < capacity) {
                char *end;
                const double v = strtod(buffer, &end);
                maxcorr_out[j] = end == buffer ? 0.0 : v;
            }
            j++;
        }
        i++;
    }
    fclose(f);
    *n_out = j;
    if (maxcorr_out && j > capacity) { rr_set_error("rr_maxcorr_read_text: %lld values in the window, room for %lld", (long long)j, (long long)capacity); return RR_E_ARG; }
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
