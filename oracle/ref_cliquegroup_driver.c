/* TEST INFRASTRUCTURE ONLY (oracle/).  Calls the UNMODIFIED /root/reference/RepeatResolver.c (compiled with
 * -Dmain=ref_rr_main by oracle/Makefile; no source is copied) for the second half of SURVEY.md section 8f row 2: its
 * reader Einlesen (293-429), Cliquer (1179-1240) and, on the clique Cliquer returns, CliqueGroup (976-1008) and
 * CliqueCoverage (1064-1096) as Group_Refinement calls them (1662-1664).
 *
 *   ref_cliquegroup_driver MSA von bis mincov maxclique greedy c,c,... a [a ...]
 * prints "R N sc" and then, per query group a and cutoff c, one line
 *   "a c n  m m m ... | g g g ... | v v v ..."   (members, then the sc words of the group and of the coverage, hex). */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

extern void Einlesen(char *MApath_p, int von, int bis);
extern int *Cliquer(int anfang, int ende, int mincov, int maxclique, double greedy, int a);
extern unsigned long *CliqueGroup(int *Clique, int c);
extern unsigned long *CliqueCoverage(int *Clique, int c);
extern int siglength, signumber, sc;

int main(int argc, char **argv)
{
    int von, bis, mincov, maxclique, k, cuts[32], ncut = 0;
    double greedy;
    char *tok;
    if (argc < 9) { fprintf(stderr, "usage: %s MSA von bis mincov maxclique greedy c,c,... a [a ...]\n", argv[0]); return 2; }
    von = atoi(argv[2]); bis = atoi(argv[3]); mincov = atoi(argv[4]); maxclique = atoi(argv[5]); greedy = atof(argv[6]);
    for (tok = strtok(argv[7], ","); tok && ncut < 32; tok = strtok(NULL, ",")) cuts[ncut++] = atoi(tok);
    Einlesen(argv[1], von, bis);
    printf("%d %d %d\n", signumber, siglength, sc);
    for (k = 8; k < argc; k++) {
        const int a = atoi(argv[k]);
        int *c = Cliquer(0, siglength, mincov, maxclique, greedy, a), n = 0, j, t;
        while (n < 100 && c[n] >= 0) n++;                /* as CliqueGroup sizes it (986-993) */
        for (t = 0; t < ncut; t++) {
            unsigned long *g = CliqueGroup(c, cuts[t]), *v = CliqueCoverage(c, cuts[t]);
            printf("%d %d %d ", a, cuts[t], n);
            for (j = 0; j < n; j++) printf(" %d", c[j]);
            printf(" |");
            for (j = 0; j < sc; j++) printf(" %lx", g[j]);
            printf(" |");
            for (j = 0; j < sc; j++) printf(" %lx", v[j]);
            printf("\n");
            free(g); free(v);
        }
        free(c);
    }
    return 0;
}
