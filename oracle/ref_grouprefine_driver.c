/* TEST INFRASTRUCTURE ONLY (oracle/).  Calls the UNMODIFIED /root/reference/RepeatResolver.c (compiled with
 * -Dmain=ref_rr_main by oracle/Makefile; no source is copied) for the whole of SURVEY.md section 8f row 2: its reader
 * Einlesen (293-429), MaxCorrsEinlesen (609-646) and Group_Refinement (1634-1693) - Cliquer, the three cutoff rules
 * (BestCutoff 530-548, KorrMaxCutoff 1393-1457, Dropoff_Cutoff 1460-1522; only the last one's result survives, 1660-1662),
 * CliqueGroup, CliqueCoverage and the two GroupPrecision printouts (1098-1131) - or, with threads > 0,
 * Parallel_Group_Refinement (1770-1821).
 *
 *   ref_grouprefine_driver MSA von bis MaxCorrsFile cutoff mincov maxclique greedy threads
 * The reference prints "Group Precision maj / min" twice per refined group while it runs; afterwards this driver prints
 * "R N sc" and, per group i whose MaxCorrs exceeded the cutoff, one line
 *   "GR i size cutoff dropoff(%a) maxcorr_after(%a) | members ... | group words (hex) | coverage words (hex)"
 * (the word lists are empty for groups that were not refined, where the reference leaves NULL). */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

extern void Einlesen(char *MApath_p, int von, int bis);
extern double *MaxCorrsEinlesen(char *inputfile, int von, int bis);
extern void Group_Refinement(double *MaxCorrs, double cutoff, int anfang, int ende, int mincov, int maxclique, double greedy);
extern void Parallel_Group_Refinement(double *MaxCorrs, double cutoff, int anfang, int ende, int mincov, int maxclique,
                                      double greedy, int NTHREADS);
extern int siglength, signumber, sc;
extern int Sizes[], Cutoffs[], *Cliques[];
extern double Drop_Off[];
extern unsigned long *C_Groups[], *C_Coverage[];

int main(int argc, char **argv)
{
    int von, bis, mincov, maxclique, threads, i, j;
    double cutoff, greedy, *M, *M0;
    if (argc != 10) { fprintf(stderr, "usage: %s MSA von bis MaxCorrsFile cutoff mincov maxclique greedy threads\n", argv[0]); return 2; }
    von = atoi(argv[2]); bis = atoi(argv[3]); cutoff = atof(argv[5]); mincov = atoi(argv[6]); maxclique = atoi(argv[7]);
    greedy = atof(argv[8]); threads = atoi(argv[9]);
    Einlesen(argv[1], von, bis);
    M = MaxCorrsEinlesen(argv[4], 0, siglength - 1);      /* the file holds the window's values */
    if (!M) { fprintf(stderr, "cannot read %s\n", argv[4]); return 1; }
    M0 = malloc(sizeof(double) * 5 * siglength);
    memcpy(M0, M, sizeof(double) * 5 * siglength);
    if (threads > 0) Parallel_Group_Refinement(M, cutoff, 0, siglength, mincov, maxclique, greedy, threads);
    else Group_Refinement(M, cutoff, 0, siglength, mincov, maxclique, greedy);
    fflush(stdout);
    printf("%d %d %d\n", signumber, siglength, sc);
    for (i = 0; i < 5 * siglength; i++) {
        /* the parallel form keeps cutoff and greedy in an int array (1793-1794) */
        if (!(M0[i] > (threads > 0 ? (double)(int)cutoff : cutoff))) continue;
        printf("GR %d %d %d %a %a |", i, Sizes[i], Cutoffs[i], Drop_Off[i], M[i]);
        for (j = 0; j <= maxclique; j++) printf(" %d", Cliques[i][j]);
        printf(" |");
        if (C_Groups[i]) for (j = 0; j < sc; j++) printf(" %lx", C_Groups[i][j]);
        printf(" |");
        if (C_Coverage[i]) for (j = 0; j < sc; j++) printf(" %lx", C_Coverage[i][j]);
        printf("\n");
    }
    return 0;
}
