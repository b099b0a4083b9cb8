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
    uint8_t *cells; /* [rows][cols] */
};

/* page-locked when a CUDA device is present, plain malloc otherwise (host buffer only) */
void *rr_host_alloc(size_t bytes, int *pinned);
void rr_host_free(void *p, int pinned);

void rr_set_error(const char *fmt, ...);

#ifdef __cplusplus
}
#endif
#endif
