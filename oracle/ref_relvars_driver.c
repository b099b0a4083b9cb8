/* TEST INFRASTRUCTURE ONLY (oracle/).  Calls the UNMODIFIED /root/reference/RepeatResolver.c (compiled with
 * -Dmain=ref_rr_main by oracle/Makefile; no source is copied) for SURVEY.md section 8f row 3: its reader Einlesen
 * (293-429), Relative_Vars (2424-2493) and, for explicitly named pairs, Relative_Group_Significance (506-523).
 *
 *   ref_relvars_driver MSA von bis maxcorrs.txt unterteilung.txt u_no cutoff mingroup [i:j ...]
 * maxcorrs.txt: 5*siglength numbers, unterteilung.txt: signumber integers (one per kept read, in reading order).
 * Prints "R N", then "VARS n  v v v ...", then per named pair "PAIR i j score" (%.17g). */
#include <stdio.h>
#include <stdlib.h>

extern void Einlesen(char *MApath_p, int von, int bis);
extern int *Relative_Vars(int *Unterteilung, int u_no, double *MaxCorrs, double cutoff, int mingroup);
extern double Relative_Group_Significance(unsigned long *Group1, unsigned long *Group2, unsigned long *Cov);
extern unsigned long *GrInitialize();
extern void GrAdd(unsigned long *group, int element);
extern unsigned long *Groups[];
extern int siglength, signumber;

int main(int argc, char **argv)
{
    int von, bis, u_no, mingroup, i, k, n = 0;
    double cutoff, *M;
    int *U, *vars;
    unsigned long *ug;
    FILE *f;
    if (argc < 9) { fprintf(stderr, "usage: %s MSA von bis maxcorrs unterteilung u_no cutoff mingroup [i:j ...]\n", argv[0]); return 2; }
    von = atoi(argv[2]); bis = atoi(argv[3]); u_no = atoi(argv[6]); cutoff = atof(argv[7]); mingroup = atoi(argv[8]);
    Einlesen(argv[1], von, bis);
    printf("%d %d\n", signumber, siglength);
    M = (double *)calloc((size_t)siglength * 5 + 1, sizeof(double));
    U = (int *)calloc((size_t)signumber + 1, sizeof(int));
    if (!(f = fopen(argv[4], "r"))) { fprintf(stderr, "cannot open %s\n", argv[4]); return 2; }
    for (i = 0; i < siglength * 5; i++) if (fscanf(f, "%lf", &M[i]) != 1) { fprintf(stderr, "short maxcorrs file\n"); return 2; }
    fclose(f);
    if (!(f = fopen(argv[5], "r"))) { fprintf(stderr, "cannot open %s\n", argv[5]); return 2; }
    for (i = 0; i < signumber; i++) if (fscanf(f, "%d", &U[i]) != 1) { fprintf(stderr, "short unterteilung file\n"); return 2; }
    fclose(f);
    vars = Relative_Vars(U, u_no, M, cutoff, mingroup);
    while (vars[n] != -1) n++;
    printf("VARS %d ", n);
    for (i = 0; i < n; i++) printf(" %d", vars[i]);
    printf("\n");
    ug = GrInitialize();
    for (i = 0; i < signumber; i++) if (U[i] == u_no) GrAdd(ug, i);
    for (k = 9; k < argc; k++) {
        int a, b;
        if (sscanf(argv[k], "%d:%d", &a, &b) != 2) continue;
        printf("PAIR %d %d %.17g\n", a, b, Relative_Group_Significance(Groups[b], Groups[a], ug));   /* argument order of 2465 */
    }
    return 0;
}
