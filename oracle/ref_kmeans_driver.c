/* TEST INFRASTRUCTURE ONLY (oracle/).  Calls the UNMODIFIED /root/reference/RepeatResolver.c (compiled with
 * -Dmain=ref_rr_main by oracle/Makefile; no source is copied) for SURVEY.md section 8f row 4: its reader Einlesen
 * (293-429) and Kmeans (2604-2821).
 *
 *   ref_kmeans_driver MSA von bis unterteilung.txt u_no mingroup var [var ...]
 * unterteilung.txt: signumber integers (one per kept read, in reading order).
 * Prints "R N", then "SPLIT n" (the return value) and "PARTS p p p ..." (the partition after the call). */
#include <stdio.h>
#include <stdlib.h>

extern void Einlesen(char *MApath_p, int von, int bis);
extern int Kmeans(int *Unterteilung, int u_no, int *Vars, int mingroup);
extern int siglength, signumber;

int main(int argc, char **argv)
{
    int von, bis, u_no, mingroup, i, n;
    int *U, *vars;
    FILE *f;
    if (argc < 7) { fprintf(stderr, "usage: %s MSA von bis unterteilung u_no mingroup [var ...]\n", argv[0]); return 2; }
    von = atoi(argv[2]); bis = atoi(argv[3]); u_no = atoi(argv[5]); mingroup = atoi(argv[6]);
    Einlesen(argv[1], von, bis);
    printf("%d %d\n", signumber, siglength);
    U = (int *)calloc((size_t)signumber + 1, sizeof(int));
    if (!(f = fopen(argv[4], "r"))) { fprintf(stderr, "cannot open %s\n", argv[4]); return 2; }
    for (i = 0; i < signumber; i++) if (fscanf(f, "%d", &U[i]) != 1) { fprintf(stderr, "short unterteilung file\n"); return 2; }
    fclose(f);
    vars = (int *)malloc(sizeof(int) * (size_t)(argc - 7 + 1));
    for (i = 7; i < argc; i++) vars[i - 7] = atoi(argv[i]);
    vars[argc - 7] = -1;                                   /* the terminator Relative_Vars writes (2483) */
    n = Kmeans(U, u_no, vars, mingroup);
    printf("SPLIT %d\n", n);
    printf("PARTS");
    for (i = 0; i < signumber; i++) printf(" %d", U[i]);
    printf("\n");
    return 0;
}
