/* TEST INFRASTRUCTURE ONLY (oracle/).  Calls the UNMODIFIED /root/reference/RepeatResolver.c (compiled with
 * -Dmain=ref_rr_main by oracle/Makefile; no source is copied): its reader Einlesen (293-429), MaxCorrsEinlesen (609-646) and
 * Kmeans_Subdivision (3382-3404) - the caller of Relative_Vars and Kmeans for every part of a read partition, between two
 * runs of Unterteilungskomprimierung (1823-1843) - as main calls it at 4065.
 *
 *   ref_subdivision_driver MSA von bis MaxCorrsFile unterteilung.txt cutoff mingroup
 * unterteilung.txt: signumber integers (one per kept read, in reading order).
 * Prints "R N" and "PARTS p p p ..." (the partition after the call). */
#include <stdio.h>
#include <stdlib.h>

extern void Einlesen(char *MApath_p, int von, int bis);
extern double *MaxCorrsEinlesen(char *inputfile, int von, int bis);
extern void Kmeans_Subdivision(int *Unterteilung, double *MaxCorrs, double cutoff, int mingroup);
extern int siglength, signumber;

int main(int argc, char **argv)
{
    int i, *U;
    double *M;
    FILE *f;
    if (argc != 8) { fprintf(stderr, "usage: %s MSA von bis MaxCorrsFile unterteilung cutoff mingroup\n", argv[0]); return 2; }
    Einlesen(argv[1], atoi(argv[2]), atoi(argv[3]));
    M = MaxCorrsEinlesen(argv[4], 0, siglength - 1);
    if (!M) { fprintf(stderr, "cannot read %s\n", argv[4]); return 2; }
    U = (int *)calloc((size_t)signumber + 1, sizeof(int));
    if (!(f = fopen(argv[5], "r"))) { fprintf(stderr, "cannot open %s\n", argv[5]); return 2; }
    for (i = 0; i < signumber; i++) if (fscanf(f, "%d", &U[i]) != 1) { fprintf(stderr, "short unterteilung file\n"); return 2; }
    fclose(f);
    Kmeans_Subdivision(U, M, atof(argv[6]), atoi(argv[7]));
    fflush(stdout);
    printf("%d %d\n", signumber, siglength);
    printf("PARTS");
    for (i = 0; i < signumber; i++) printf(" %d", U[i]);
    printf("\n");
    return 0;
}
