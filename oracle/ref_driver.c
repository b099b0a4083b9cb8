/* TEST INFRASTRUCTURE ONLY (oracle/).  Driver around the UNMODIFIED reference
 * /root/reference/MaxCorrelation.c compiled with -Dmain=ref_main (every function and
 * global of that file has external linkage).  It calls the reference's own
 * Einlesen (270-393) and HilfsMaxCorrsRechner (745-837) so that
 *   - the scan can be timed without the parse, and
 *   - an exact 1/k cyclic row sample can be run (ii % NTHREADS == thread, line 796):
 *     NTHREADS = modulus, threads res_lo..res_hi-1.
 * usage: ref_driver MSA cov modulus res_lo res_hi [outfile|-] [reps]
 * prints one line per repetition:  REF R N threads scan_seconds parse_seconds first_residue
 * Repetition r scans the residues res_lo + r * threads .. (the file is parsed once); residues must stay below
 * min(modulus, 128), the size of the reference's HilfsMaxCorr[] (MaxCorrelation.c:36).
 */
#include <stdio.h>
#include <stdlib.h>
#include <time.h>
#include <pthread.h>

extern void Einlesen(char *MApath_p, int von, int bis);
extern void *HilfsMaxCorrsRechner(void *x);
extern double *HilfsMaxCorr[128];
extern int siglength, signumber;

static double now(void)
{
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec + 1e-9 * ts.tv_nsec;
}

int main(int argc, char **argv)
{
    int cov, modulus, lo, hi, nt, t, i, reps = 1, rep;
    pthread_t th[128];
    int *args[128];
    double t0, t1, t2;
    if (argc < 6) { fprintf(stderr, "usage: ref_driver MSA cov modulus res_lo res_hi [out]\n"); return 2; }
    cov = atoi(argv[2]); modulus = atoi(argv[3]); lo = atoi(argv[4]); hi = atoi(argv[5]);
    nt = hi - lo;
    if (argc > 7) reps = atoi(argv[7]);
    if (nt < 1 || reps < 1 || lo < 0 || lo + reps * nt > 128 || lo + reps * nt > modulus) {
        fprintf(stderr, "need 0 <= res_lo and res_lo + reps * threads <= min(modulus, 128)\n");
        return 2;
    }
    t0 = now();
    Einlesen(argv[1], -1, -1);
    t1 = now();
    for (rep = 0; rep < reps; rep++, lo += nt) {
    if (rep > 0) t1 = now();
    for (t = 0; t < nt; t++) {
        /* same 7-int argument block as Parallel_AllMaxCorrsRechner builds (853-860) */
        args[t] = (int *)malloc(sizeof(int) * 7);
        args[t][0] = 0; args[t][1] = siglength; args[t][2] = cov; args[t][3] = signumber;
        args[t][4] = 0; args[t][5] = modulus; args[t][6] = lo + t;
        pthread_create(&th[t], NULL, HilfsMaxCorrsRechner, args[t]);
    }
    for (t = 0; t < nt; t++) pthread_join(th[t], NULL);
    for (i = 0; i < siglength * 5; i++)            /* merge, 882-891 */
        for (t = 1; t < nt; t++)
            if (HilfsMaxCorr[lo + t][i] > HilfsMaxCorr[lo][i]) HilfsMaxCorr[lo][i] = HilfsMaxCorr[lo + t][i];
    t2 = now();
    if (rep == 0 && argc > 6 && argv[6][0] != '-') {
        FILE *f = fopen(argv[6], "w");
        if (!f) return 1;
        for (i = 0; i < siglength * 5; i++) fprintf(f, "%f\n", HilfsMaxCorr[lo][i]);
        fclose(f);
    }
    printf("REF %d %d %d %.6f %.6f %d\n", signumber, siglength, nt, t2 - t1, rep == 0 ? t1 - t0 : 0.0, lo);
    fflush(stdout);
    }
    return 0;
}
