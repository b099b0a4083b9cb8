/* TEST INFRASTRUCTURE ONLY (oracle/).  Calls the UNMODIFIED /root/reference/RepeatResolver.c (compiled with
 * -Dmain=ref_rr_main by oracle/Makefile; no source is copied): its reader of the result file, MaxCorrsEinlesen (609-646).
 *
 *   ref_einlesen_driver MaxCorrsFile von bis siglength count
 * siglength sizes the reference's array (siglength * 5 doubles, 627); the first `count` values are printed as hex floats,
 * one per line ("NULL" if the reference returned NULL). */
#include <stdio.h>
#include <stdlib.h>

extern double *MaxCorrsEinlesen(char *inputfile, int von, int bis);
extern int siglength;

int main(int argc, char **argv)
{
    double *M;
    int i, count;
    if (argc != 6) { fprintf(stderr, "usage: %s MaxCorrsFile von bis siglength count\n", argv[0]); return 2; }
    siglength = atoi(argv[4]);
    count = atoi(argv[5]);
    M = MaxCorrsEinlesen(argv[1], atoi(argv[2]), atoi(argv[3]));
    if (!M) { printf("NULL\n"); return 0; }
    for (i = 0; i < count; i++) printf("%a\n", M[i]);
    return 0;
}
