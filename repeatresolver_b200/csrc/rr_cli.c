/* rr_cli.c -- drop-in for the reference program: same command line, same output file.
 *
 *   ./MaxCorrelation <MSApath> [-c <cov, default 30>] [-p <n>] [-f a b]
 *
 * mirrors main() of /root/reference/MaxCorrelation.c:916-1026: the MSA path is argv[1]
 * (924); a flag matches on argv[i][0]=='-' && argv[i][1]==letter and takes the next
 * argument(s) (937-972); -f is parsed and ignored (964-972, its filter is commented out at
 * 299); the result goes to the literal concatenation "MaxCorrsOf_" + argv[1] (991-993) as
 * 5*siglength lines "%f\n" (516-532); exit code 0, or 1 with "MA is missing." when the file
 * cannot be opened (284).  Differences: -p is the number of B200s to use (the reference's
 * thread count has no meaning here; values above the device count are clamped).  -p 0 is
 * REFUSED: in the reference it selects a different routine (AllMaxCorrsRechner, 575-635:
 * no base-count filter, groups with fewer than 5 correlations above the cutoff zeroed, 1010-
 * 1013), which this program does not implement - silently running the -p >= 1 semantics
 * would write a different file.  Only the first -p devices are made visible to CUDA before
 * its first call.  The arg-max partners go to a side file "MaxCorrsArgOf_" + argv[1] (never into
 * MaxCorrsOf_*); --variant bitset|umma, --no-finalize and --bin (also write the binary side
 * file "MaxCorrsBinOf_" + argv[1]) are extra switches.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include "../../include/rr_maxcorr.h"

void rr_trace_mark(const char *tag); /* librr_maxcorr.so internal: RR_TRACE=1 prints host phase times */

int main(int argc, char *argv[])
{
    const char *path;
    int cov = 30, parallel = 1, i, rc, variant = RR_VARIANT_AUTO, ndev, write_bin = 0;
    unsigned flags = RR_FLAG_HOST_FINALIZE;
    rr_msa *msa = NULL;
    rr_scan_stats st;
    double *M;
    int32_t *A;
    char name[4096];
    time_t t0 = time(NULL);
    long G;

    if (argc < 2) { printf("Usage: ./MaxCorrelation MSApath <options>\n"); exit(0); }
    path = argv[1];
    for (i = 2; i < argc; i++) {
        if (argv[i][0] == '-' && argv[i][1] == 'p' && i + 1 < argc) {
            parallel = (int)strtol(argv[i + 1], NULL, 10);
            printf("GPUS: %d\n", parallel);
        }
    }
    for (i = 2; i < argc; i++) {
        if (argv[i][0] == '-' && argv[i][1] == 'c' && i + 1 < argc) {
            cov = (int)strtol(argv[i + 1], NULL, 10);
            printf("Coverage %d\n", cov);
        }
    }
    for (i = 2; i < argc; i++) {
        if (argv[i][0] == '-' && argv[i][1] == 'f' && i + 2 < argc)
            printf("Full coverage from column %ld until %ld.\n", strtol(argv[i + 1], NULL, 10), strtol(argv[i + 2], NULL, 10));
        if (!strcmp(argv[i], "--variant") && i + 1 < argc)
            variant = !strcmp(argv[i + 1], "bitset") ? RR_VARIANT_BITSET : !strcmp(argv[i + 1], "umma") ? RR_VARIANT_UMMA : RR_VARIANT_AUTO;
        if (!strcmp(argv[i], "--bin")) write_bin = 1;
        if (!strcmp(argv[i], "--no-finalize")) flags &= ~RR_FLAG_HOST_FINALIZE;
        if (!strcmp(argv[i], "--no-prune")) flags |= RR_FLAG_NO_PRUNE;
    }
    if (parallel < 1) {
        fprintf(stderr, "\nError in MaxCorrelation\n   -p %d: the serial routine of the reference (AllMaxCorrsRechner, its count filter and "
                        "printing) is not implemented; use -p 1..8 = number of GPUs\n", parallel);
        exit(1);
    }
    {   /* make only the first -p devices visible: driver start-up and context creation scale with the visible devices */
        const char *vis = getenv("CUDA_VISIBLE_DEVICES");
        char buf[512];
        size_t n = 0;
        int k;
        if (vis && *vis) {
            int commas = 0;
            for (n = 0; vis[n] && n + 1 < sizeof buf; n++) {
                if (vis[n] == ',' && ++commas == parallel) break;
                buf[n] = vis[n];
            }
            buf[n] = 0;
        } else {
            for (k = 0; k < parallel && n + 8 < sizeof buf; k++) n += (size_t)snprintf(buf + n, sizeof buf - n, k ? ",%d" : "%d", k);
        }
        setenv("CUDA_VISIBLE_DEVICES", buf, 1);
    }
    rr_trace_mark("cli: start");
    rc = rr_msa_read(path, &msa);
    rr_trace_mark("cli: MSA read");
    if (rc == RR_E_IO) { printf("MA is missing.\n"); exit(1); }
    if (rc) { fprintf(stderr, "\nError in MaxCorrelation\n   %s\n", rr_last_error()); exit(1); }
    printf("There are %d sequences.\n", rr_msa_rows(msa));
    printf("Siglength is %d.\n", rr_msa_cols(msa));
    ndev = rr_device_count();
    if (ndev < 1) { fprintf(stderr, "\nError in MaxCorrelation\n   no CUDA device (there is no CPU path)\n"); exit(1); }
    if (parallel > ndev) parallel = ndev;
    G = 5L * rr_msa_cols(msa);
    M = (double *)calloc((size_t)G + 1, sizeof(double));
    A = (int32_t *)calloc((size_t)G + 1, sizeof(int32_t));
    if (!M || !A) { fprintf(stderr, "\nError in MaxCorrelation\n   Out of memory\n"); exit(1); }
    rc = rr_maxcorr_run(msa, cov, parallel, variant, flags, M, A, &st);
    if (rc) { fprintf(stderr, "\nError in MaxCorrelation\n   %s\n", rr_last_error()); exit(1); }
    snprintf(name, sizeof name, "MaxCorrsOf_%s", path);
    printf("%s\n", name);
    rr_trace_mark("cli: scan done");
    rc = rr_maxcorr_write(name, M, G);
    if (rc) { printf("DateiVerbratei!\n"); exit(1); }
    snprintf(name, sizeof name, "MaxCorrsArgOf_%s", path);
    rr_argmax_write(name, A, G);
    if (write_bin) {   /* full-precision values + partners for consumers that link the library (rr_maxcorr_read_bin) */
        snprintf(name, sizeof name, "MaxCorrsBinOf_%s", path);
        if (rr_maxcorr_write_bin(name, M, A, G)) { printf("DateiVerbratei!\n"); exit(1); }
    }
    rr_trace_mark("cli: files written");
    printf("pair tests %lld, exact evaluations %lld, kernel %.3f ms (%s), pack %.3f ms, h2d %.3f ms\n",
           (long long)st.pair_tests, (long long)st.exact_evals, st.kernel_ms,
           st.variant == RR_VARIANT_UMMA ? "tcgen05 int8" : st.variant == RR_VARIANT_UMMA_F4 ? "tcgen05 e2m1" : st.variant == RR_VARIANT_UMMA_MXF4 ? "tcgen05 mxf4" : "bitset", st.pack_ms, st.h2d_ms);
    printf("Runtime: %lu sec.\n", (unsigned long)(time(NULL) - t0));
    rr_msa_free(msa);
    free(M); free(A);
    exit(0);
}
