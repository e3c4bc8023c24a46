/* sdmd_sdmin.c — SDEnergyMin (ToolFunctions.c:325-491) with a linear-time neighbour search.
 *
 * The reference builds every atom's list of "surrounding atoms" (all n with |r_n - r_i|^2 <= 144 A^2 that are neither bonded nor
 * constrained to i, in ascending n) by testing all N^2 pairs.  Here the atoms are binned once into a uniform grid of 12 A cells over
 * their bounding box (the reference applies no periodic image in this function, so neither does the grid); an atom's candidates are
 * the atoms of the 27 cells around it, tested with the reference's own expression and taken in ascending n, so the lists — and with
 * them the relaxation, which walks them in order and draws one RandomValue per violated entry — are the reference's, bit for bit.
 * Everything after the list construction is the reference's loop restated (same arithmetic order, same draws, same progress lines,
 * same MiniEnergy*.gro dumps).
 *
 * Build: the reference's ToolFunctions.c is compiled with -DSDEnergyMin=sdmd_reference_SDEnergyMin, so this SDEnergyMin is the
 * one Prepare() calls.  Small systems keep the reference's own code (the O(N^2) search is seconds up to ~10^4 atoms);
 * SDMD_FAST_SDMIN=1 / 0 forces the choice (tests force the grid search on small systems and compare with the reference). */
#include "DMD.h"

void sdmd_reference_SDEnergyMin(long int stepNum, struct AtomListStr *thisList);

struct sdmin_inter {
    int connect[512];
    double coordinate[4];
    double hiLimit[512];
    double loLimit[512];
    double length[512];
};

static int cmp_int(const void *a, const void *b) {
    int x = *(const int *)a, y = *(const int *)b;
    return (x > y) - (x < y);
}

void SDEnergyMin(long int stepNum, struct AtomListStr *thisList) {
    const char *env = getenv("SDMD_FAST_SDMIN");
    int fast = env ? atoi(env) : (atomnum > 8192);
    if (!fast) { sdmd_reference_SDEnergyMin(stepNum, thisList); return; }

    long step = 0;
    long totalCN = INFINIT * INFINIT;
    double factor = 0.1;
    double diff;
    double energy;
    double direction[4];
    double length2;
    char directory[100];
    struct sdmin_inter *interaction;
    struct FileStr outputFile;

    interaction = (struct sdmin_inter *)calloc(atomnum + 1, sizeof(struct sdmin_inter));
    sprintf(directory, "%s/MiniEnergy%s.gro", datadir, REMDInfo.REMD_ExtraName);
    outputFile.file = fopen(directory, "w");

    /* ---- uniform grid of 12 A cells over the bounding box (counting sort of the atoms by cell) ---- */
    const double R = 12.0;
    double lo[4] = {0, INFINIT, INFINIT, INFINIT}, hi[4] = {0, -INFINIT, -INFINIT, -INFINIT};
    for (int i = 1; i <= atomnum; ++i)
        for (int k = 1; k <= 3; k++) {
            double c = thisList->ptr[i]->dynamic->coordinate[k];
            if (c < lo[k]) lo[k] = c;
            if (c > hi[k]) hi[k] = c;
        }
    int gn[4] = {0, 1, 1, 1};
    for (int k = 1; k <= 3; k++) {
        double span = hi[k] - lo[k];
        gn[k] = (int)(span / R) + 1;
        if (gn[k] > 1024) gn[k] = 1024;       /* (a cell may then be wider than 12 A: still a superset of the neighbourhood) */
    }
    double gw[4];
    for (int k = 1; k <= 3; k++) gw[k] = (hi[k] - lo[k]) / gn[k] > R ? (hi[k] - lo[k]) / gn[k] : R;
    long ncell = (long)gn[1] * gn[2] * gn[3];
    int *cell_of = (int *)calloc(atomnum + 1, sizeof(int));
    int *cell_start = (int *)calloc(ncell + 1, sizeof(int));
    int *cell_atoms = (int *)calloc(atomnum + 1, sizeof(int));
    for (int i = 1; i <= atomnum; ++i) {
        int g[4];
        for (int k = 1; k <= 3; k++) {
            g[k] = (int)((thisList->ptr[i]->dynamic->coordinate[k] - lo[k]) / gw[k]);
            if (g[k] < 0) g[k] = 0;
            if (g[k] >= gn[k]) g[k] = gn[k] - 1;
        }
        cell_of[i] = (g[3] * gn[2] + g[2]) * gn[1] + g[1];
        cell_start[cell_of[i] + 1]++;
    }
    for (long c = 0; c < ncell; c++) cell_start[c + 1] += cell_start[c];
    {
        int *fill = (int *)calloc(ncell + 1, sizeof(int));
        for (int i = 1; i <= atomnum; ++i) cell_atoms[cell_start[cell_of[i]] + fill[cell_of[i]]++] = i;   /* ascending i inside a cell */
        free(fill);
    }
    int cand_cap = 4096;
    int *cand = (int *)malloc(sizeof(int) * cand_cap);

    for (int i = 1; i <= atomnum; ++i) {

        TRANSFER_VECTOR(interaction[i].coordinate, thisList->ptr[i]->dynamic->coordinate);

        /* bond information, ToolFunctions.c:353-362 */
        int m = 0;
        struct ConstraintStr *thisBond = thisList->ptr[i]->property->bond;
        while (thisBond != NULL) {
            m++;
            interaction[i].connect[m] = thisBond->connection;
            interaction[i].hiLimit[m] = sqrt(thisBond->dmax);
            interaction[i].loLimit[m] = sqrt(thisBond->dmin);
            interaction[i].length[m]  = 0.5 * (interaction[i].hiLimit[m] + interaction[i].loLimit[m]);
            thisBond = thisBond->next;
        }

        /* constraint information, :364-373 */
        struct ConstraintStr *thisConstr = thisList->ptr[i]->property->constr;
        while (thisConstr != NULL) {
            m++;
            interaction[i].connect[m] = thisConstr->connection;
            interaction[i].hiLimit[m] = sqrt(thisConstr->dmax);
            interaction[i].loLimit[m] = sqrt(thisConstr->dmin);
            interaction[i].length[m]  = 0.5 * (interaction[i].hiLimit[m] + interaction[i].loLimit[m]);
            thisConstr = thisConstr->next;
        }

        /* the surrounding atoms, :375-404: candidates from the 27 grid cells around i, in ascending n */
        int nc = 0;
        {
            int c = cell_of[i];
            int g1 = c % gn[1], g2 = (c / gn[1]) % gn[2], g3 = c / (gn[1] * gn[2]);
            for (int d3 = -1; d3 <= 1; d3++)
                for (int d2 = -1; d2 <= 1; d2++)
                    for (int d1 = -1; d1 <= 1; d1++) {
                        int h1 = g1 + d1, h2 = g2 + d2, h3 = g3 + d3;
                        if (h1 < 0 || h2 < 0 || h3 < 0 || h1 >= gn[1] || h2 >= gn[2] || h3 >= gn[3]) continue;
                        int cc = (h3 * gn[2] + h2) * gn[1] + h1;
                        for (int s = cell_start[cc]; s < cell_start[cc + 1]; s++) {
                            if (nc == cand_cap) { cand_cap *= 2; cand = (int *)realloc(cand, sizeof(int) * cand_cap); }
                            cand[nc++] = cell_atoms[s];
                        }
                    }
            qsort(cand, nc, sizeof(int), cmp_int);
        }
        for (int q = 0; q < nc; q++) {
            int n = cand[q];
            direction[1] = thisList->ptr[n]->dynamic->coordinate[1] - thisList->ptr[i]->dynamic->coordinate[1];
            direction[2] = thisList->ptr[n]->dynamic->coordinate[2] - thisList->ptr[i]->dynamic->coordinate[2];
            direction[3] = thisList->ptr[n]->dynamic->coordinate[3] - thisList->ptr[i]->dynamic->coordinate[3];

            length2 = DOT_PROD(direction, direction);
            if (length2 <= 144 && //radius of 12 A
                !(connectionMap[i][n] & BOND_CONNECT) &&
                !(connectionMap[i][n] & CONSTRAINT_CONNECT)) {
                if (i != n && !FindElemInList(interaction[i].connect, n, -1, -1, 0)) {
                    m++;
                    if (m > 510) break;
                    interaction[i].connect[m] = n;
                    interaction[i].hiLimit[m] = INFINIT;

                    int type1 = thisList->ptr[i]->property->typeofAtom;
                    int type2 = thisList->ptr[n]->property->typeofAtom;

                    if (type1 > type2) {
                        int t = type2;
                        type2 = type1;
                        type1 = t;
                    }

                    interaction[i].loLimit[m] = sqrt(potentialPairCollision[type1][type2].dmin);
                    interaction[i].length[m]  = interaction[i].loLimit[m] + 0.001;
                }
            }
        }

        if (m > 510) {
            printf("!!ERROR!! too many interacting atoms surround atom #%i, which has %i. code needs to be optimized! %s:%i\n", i, m, __FILE__, __LINE__);
            exit(EXIT_FAILURE);
        }
    }
    free(cand); free(cell_atoms); free(cell_start); free(cell_of);

    /* ---- the relaxation, ToolFunctions.c:412-484 ---- */
    printf("\nMinimizing the potential energy:\n");
    while (totalCN > 0) {
        for (int i = 1; i <= atomnum; i ++) {
            int num = 1;
            double change[4] = {0};
            while (interaction[i].connect[num] != 0) {
                direction[1] = interaction[interaction[i].connect[num]].coordinate[1] - interaction[i].coordinate[1];
                direction[2] = interaction[interaction[i].connect[num]].coordinate[2] - interaction[i].coordinate[2];
                direction[3] = interaction[interaction[i].connect[num]].coordinate[3] - interaction[i].coordinate[3];

                length2 = DOT_PROD(direction, direction);
                length2 = sqrt(length2);

                if (length2 >= interaction[i].hiLimit[num] || length2 <= interaction[i].loLimit[num]) {
                    diff = interaction[i].length[num] - length2;
                    energy = RandomValue(0, 1) * factor * diff;

                    change[1] -= energy * direction[1] / interaction[i].length[num];
                    change[2] -= energy * direction[2] / interaction[i].length[num];
                    change[3] -= energy * direction[3] / interaction[i].length[num];
                }
                num ++;
            }

            interaction[i].coordinate[1] += change[1];
            interaction[i].coordinate[2] += change[2];
            interaction[i].coordinate[3] += change[3];
        }

        /* recheck, :442-478 */
        totalCN = 0;
        for (int i = 1; i <= atomnum; i ++) {
            int num = 1;
            while (interaction[i].connect[num] != 0) {
                if (interaction[i].connect[num] < i) {
                    num ++;
                    continue;
                }

                direction[1] = interaction[interaction[i].connect[num]].coordinate[1] - interaction[i].coordinate[1];
                direction[2] = interaction[interaction[i].connect[num]].coordinate[2] - interaction[i].coordinate[2];
                direction[3] = interaction[interaction[i].connect[num]].coordinate[3] - interaction[i].coordinate[3];

                length2 = DOT_PROD(direction, direction);
                length2 = sqrt(length2);

                if (length2 >= interaction[i].hiLimit[num] || length2 <= interaction[i].loLimit[num]) {
                    if (step == stepNum) {
                        printf("atoms between %8i (%s [%4i%3s]) and %8i (%s [%4i%3s]), cur dis = %.4lf, min = %.4lf, max = %.4lf\n",
                               i,
                               atom[i].property->nameofAA,
                               atom[i].property->sequence.aminoacidNum,
                               atom[i].property->name,
                               interaction[i].connect[num],
                               atom[interaction[i].connect[num]].property->nameofAA,
                               atom[interaction[i].connect[num]].property->sequence.aminoacidNum,
                               atom[interaction[i].connect[num]].property->name,
                               length2,
                               interaction[i].loLimit[num],
                               interaction[i].hiLimit[num]);
                    }
                    totalCN ++;
                }
                num ++;
            }
            TRANSFER_VECTOR(thisList->ptr[i]->dynamic->coordinate, interaction[i].coordinate);
        }

        printf("step = %5li, total conflicts = %li\n", ++step, totalCN);
        if (step % 1000 == 0) SaveGRO(&outputFile);
        if (step % 1000 == 0 && factor > 0.05) factor -= 0.01;
        if (step > stepNum) break;
    }

    SaveGRO(&outputFile);

    free(interaction);
    fclose(outputFile.file);
    printf("Done!\n\n");
}
