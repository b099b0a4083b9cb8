/* TEST INFRASTRUCTURE ONLY (oracle/).  Calls the UNMODIFIED /root/reference/RepeatResolver.c (compiled with
 * -Dmain=ref_rr_main by oracle/Makefile; no source is copied): its reader Einlesen (293-429) alone.
 *
 *   ref_window_driver MSA von bis
 * prints "R N sc lines", then "A a a a ..." (Ausgelassen per line of the file), then one line "G w w w ..." per group (5 N
 * lines, the sc words of Groups[i], hex) and one line "C w w w ..." per column (LocalCoverage[i]). */
#include <stdio.h>
#include <stdlib.h>

extern void Einlesen(char *MApath_p, int von, int bis);
extern int siglength, signumber, sc, realsigno;
extern int Ausgelassen[];
extern unsigned long *Groups[], *LocalCoverage[];

int main(int argc, char **argv)
{
    int i, j;
    if (argc != 4) { fprintf(stderr, "usage: %s MSA von bis\n", argv[0]); return 2; }
    Einlesen(argv[1], atoi(argv[2]), atoi(argv[3]));
    fflush(stdout);
    printf("%d %d %d %d\n", signumber, siglength, sc, realsigno);
    printf("A");
    for (i = 0; i < realsigno; i++) printf(" %d", Ausgelassen[i]);
    printf("\n");
    for (i = 0; i < 5 * siglength; i++) {
        printf("G");
        for (j = 0; j < sc; j++) printf(" %lx", Groups[i][j]);
        printf("\n");
    }
    for (i = 0; i < siglength; i++) {
        printf("C");
        for (j = 0; j < sc; j++) printf(" %lx", LocalCoverage[i][j]);
        printf("\n");
    }
    return 0;
}
