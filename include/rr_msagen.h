/* rr_msagen.h -- synthetic MSAreal generator (benchmark/test input; C ABI).
 *
 * Not a replacement for a reference interface: the reference generates its inputs with
 * DataSimulator.py (Python 2) + ReadCutter + InitialAligner + PW_ReAligner.  This
 * generator synthesises the perfect-alignment MSA those programs converge to, using
 * DataSimulator.py's distributions (DataSimulator.py:12-27, 29-115, 122-160), so that
 * the shapes named in BASELINE.json can be produced in seconds.  See csrc/msagen.c.
 */
#ifndef RR_MSAGEN_H
#define RR_MSAGEN_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

enum { RR_MSAGEN_TREE = 0, RR_MSAGEN_DISTRIBUTED = 1, RR_MSAGEN_EQUIDISTANT = 2 };

typedef struct rr_msagen_params {
    int type;          /* RR_MSAGEN_*            (DataSimulator.py -t) */
    int copies;        /* copy number            (-n, default 100) */
    int coverage;      /* per-copy coverage      (-c, default 40) */
    int repeat_len;    /* repeat length in bases (-l, default 30000) */
    double diff;       /* copy difference as a fraction (-d 1 -> 0.01) */
    uint64_t seed;
    int flank;         /* unique flank on each side (10000, DataSimulator.py:222-225) */
    int min_overlap;   /* reads overlapping the repeat by fewer bases are dropped */
    int max_reads;     /* 0 = no cap */
    int threads;       /* 0 = 8 */
} rr_msagen_params;

typedef struct rr_msagen rr_msagen;

rr_msagen *rr_msagen_create(const rr_msagen_params *p);
void rr_msagen_free(rr_msagen *g);
int rr_msagen_rows(const rr_msagen *g);
int rr_msagen_cols(const rr_msagen *g);
const int *rr_msagen_read_copy(const rr_msagen *g);          /* [rows] ground-truth copy id */
void rr_msagen_fill_codes(rr_msagen *g, uint8_t *codes);     /* codes[rows][cols], 0..5 */
void rr_msagen_fill_text(rr_msagen *g, char *text);          /* text[rows][cols+1], '\n' ended */
int rr_msagen_write(rr_msagen *g, const char *path);         /* 0 ok */

#ifdef __cplusplus
}
#endif
#endif
