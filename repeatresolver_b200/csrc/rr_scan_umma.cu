// rr_scan_umma.cu -- count-kernel variant A (tcgen05 int8) -- placeholder until the kernel lands.
#include "rr_kernels.h"
#include "rr_device.cuh"
#include "rr_plan.h"

struct rr_umma_state { int unused; };
int rr_umma_available(void) { return 0; }
int rr_umma_row_sites(void) { return 24; }
int rr_umma_col_sites(void) { return 48; }
int rr_umma_kblock(void) { return 128; }
void rr_umma_free(rr_umma_state *s) { delete s; }
int rr_umma_scan(rr_umma_state *&, rr_scan_params &, rr_plan &, const uint8_t *, const int32_t *, int, int, cudaStream_t)
{
    rr_set_error("tcgen05 variant not built");
    return RR_E_ARG;
}
