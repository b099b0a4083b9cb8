/* TEST INFRASTRUCTURE ONLY (oracle/).  Calls the UNMODIFIED /root/reference/RepeatResolver.c (compiled with
 * -Dmain=ref_rr_main by oracle/Makefile; no source is copied) for SURVEY.md section 8f row 2: its reader Einlesen
 * (293-429: rows with a symbol at both ends of the column window [von, bis]) and Cliquer (1179-1240), plus
 * Group_PositiveSignificance (472-488) for the scores of the members.
 *
 *   ref_cliquer_driver MSA von bis mincov maxclique greedy a [a ...]
 * prints "R N" and then, per query group a, one line "a n  member:score ..." (scores as %.17g). */
#include <stdio.h>
#include <stdlib.h>

extern void Einlesen(char *MApath_p, int von, int bis);
extern int *Cliquer(int anfang, int ende, int mincov, int maxclique, double greedy, int a);
extern double Group_PositiveSignificance(unsigned long *Group1, unsigned long *Group2, unsigned long *Cov1,
                                         unsigned long *Cov2);
extern unsigned long *Groups[];
extern unsigned long *LocalCoverage[];
extern int siglength, signumber;

int main(int argc, char **argv)
{
    int von, bis, mincov, maxclique, k;
    double greedy;
    if (argc < 8) { fprintf(stderr, "usage: %s MSA von bis mincov maxclique greedy a [a ...]\n", argv[0]); return 2; }
    von = atoi(argv[2]); bis = atoi(argv[3]); mincov = atoi(argv[4]); maxclique = atoi(argv[5]); greedy = atof(argv[6]);
    Einlesen(argv[1], von, bis);
    printf("%d %d\n", signumber, siglength);
    for (k = 7; k < argc; k++) {
        const int a = atoi(argv[k]);
        int *c = Cliquer(0, siglength, mincov, maxclique, greedy, a), n = 0, j;
        while (n < maxclique && c[n] >= 0) n++;
        printf("%d %d", a, n);
        for (j = 1; j < n; j++)
            printf(" %d:%.17g", c[j],
                   Group_PositiveSignificance(Groups[c[j]], Groups[a], LocalCoverage[c[j] / 5], LocalCoverage[a / 5]));
        printf("\n");
        free(c);
    }
    return 0;
}
