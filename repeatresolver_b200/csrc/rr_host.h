/* rr_host.h -- internal declarations shared by the host-side C sources and the CUDA
 * translation units of librr_maxcorr.so (not installed; the public ABI is
 * include/rr_maxcorr.h). */
#ifndef RR_HOST_H
#define RR_HOST_H
#include <stddef.h>
#include <stdint.h>
#include "../../include/rr_maxcorr.h"
#ifdef __cplusplus
extern "C" {
#endif

struct rr_msa {
    int rows;       /* signumber (MaxCorrelation.c:335) */
    int cols;       /* siglength (MaxCorrelation.c:291) */
    int codes;      /* cells hold 0..5 codes instead of raw characters */
    int pinned;     /* cells came from rr_host_alloc with page-locking */
    uint8_t *cells; /* [rows][cols]; NULL while the MSA is still text-backed (see below) */
    /* text-backed MSA (rr_msa_read): the file stays mapped and row r is the cols bytes at map + rowoff[r]; the rows
     * are gathered straight into the upload ring by rr_pack, so a 1-2 GB MSA is never copied (or page-locked) on the
     * host.  rr_msa_cells() materialises the matrix on demand. */
    const char *map;
    size_t map_len;
    int64_t *rowoff;
};

/* row r of either representation */
static inline const uint8_t *rr_msa_row(const struct rr_msa *m, int r)
{
    return m->cells ? m->cells + (size_t)r * (size_t)m->cols : (const uint8_t *)m->map + m->rowoff[r];
}
/* first-use CUDA initialisation (context creation) on a background thread / wait for it */
void rr_cuda_warmup_begin(void);
void rr_cuda_warmup_end(void);

/* page-locked when a CUDA device is present, plain malloc otherwise (host buffer only) */
void *rr_host_alloc(size_t bytes, int *pinned);
void rr_host_free(void *p, int pinned);

void rr_set_error(const char *fmt, ...);

/* RR_TRACE=1 in the environment: one stderr line per host phase, milliseconds since the first call */
void rr_trace_mark(const char *tag);

#ifdef __cplusplus
}
#endif
#endif
