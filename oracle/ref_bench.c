/*
 * ref_bench.c — TEST INFRASTRUCTURE / CPU baseline.  The UNMODIFIED sDMD reference (-DGEL build, objects compiled by
 * oracle/Makefile from the sources where they lie) timed on the hot path at sizes its own driver cannot start:
 *
 *     PBC("full") + LinkList("full") + Predict(all atoms) + CreateCBT()          (Prepare(), SGThread.c:44-62)
 *
 * What stops the stock binary above 46 339 atoms is one allocation: the dense connectionMap,
 * INT_2CALLOC(connectionMap, atomnum + 1, atomnum + 1) (Initialization.c:976, DMD.h:69-72), whose element count is computed
 * in `int` and overflows.  A bond-free GEL box never writes the map and reads only zeros from it (TimePrediction.c:95,109,149),
 * so this harness links with -Wl,--wrap=calloc and answers that ONE request (an element count that cannot be a real request:
 * negative as an int, i.e. > 2^63 as a size_t) with a window of untouched, never-written, read-only anonymous memory as large
 * as the map would have been, (N+1)^2 ints = 4 TB of ADDRESS SPACE at 10^6 atoms (MAP_NORESERVE; every page that is read maps
 * the kernel's shared zero page — the huge zero page where transparent huge pages are on, so the page tables stay small), plus
 * 8 GB on either side in case the compiler wraps the row offsets NN*size2 to 32 bits.  Every connectionMap[i][j] the reference
 * evaluates is then a zero — SURVEY 8c's "aliased rows", without touching a line of the reference.
 *
 * usage: ref_bench <passes> <slots.bin|-> [<state.bin>] -- <sDMD args...>
 *   prints one line per pass: "pass <k> seconds <t>" (the reference keeps no evaluation counter; callers take the pair count
 *   from the oracle port, which is pinned to the reference) and, if a file is given, writes every atom's event slot after
 *   the LAST pass: natom x { double time, potential; int32 partner, eventType, subEventType, pad; int64 partnerCounter }.
 *   <state.bin> receives what the first pass starts from, so the same pass can be run elsewhere (parity at 10^6 atoms against
 *   the reference itself): int64 natom; double box[3]; double x[n], y[n], z[n], vx[n], vy[n], vz[n]; double gauss[n] — the
 *   thermostat draws the pass is about to make (Maxwell_Boltzmann_Distribution on glibc rand(), drawn ahead and rewound).
 */
#define _GNU_SOURCE
#include <sys/mman.h>
#include <time.h>
#include "DMD.h"

extern void *__real_calloc(size_t nmemb, size_t size);
void *__wrap_calloc(size_t nmemb, size_t size) {
    if (size == sizeof(int) && nmemb > ((size_t)1 << 62)) {           /* (int)(N+1)*(N+1) overflowed and was sign-extended */
        const size_t half = (size_t)1 << 33;                           /* +- 2^31 ints */
        const size_t rows = ((size_t)atomnum + 1) * ((size_t)atomnum + 1) * sizeof(int);
        const size_t len = (2 * half + rows + (2u << 20)) & ~(((size_t)2 << 20) - 1);
        char *w = mmap(NULL, len, PROT_READ, MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
        if (w == MAP_FAILED) { perror("ref_bench: mmap of the zero window"); exit(EXIT_FAILURE); }
        madvise(w, len, MADV_HUGEPAGE);
        fprintf(stderr, "ref_bench: connectionMap (%d+1)^2 answered with a read-only zero window of %.2f TB of address space\n", atomnum,
                len / 1099511627776.0);
        return w + half;
    }
    return __real_calloc(nmemb, size);
}

/* glibc random() state save / restore (rand() shares random()'s state); -std=c99 hides the POSIX prototypes */
extern char *initstate(unsigned int seed, char *state, size_t n);
extern char *setstate(char *state);
static char rng_scratch[128] __attribute__((aligned(8)));
static void rng_save(char saved[128]) {
    char *sys = initstate(1, rng_scratch, 128);
    memcpy(saved, sys, 128);
    setstate(sys);
}
static void rng_restore(const char saved[128]) {
    static char live[128] __attribute__((aligned(8)));
    memcpy(live, saved, 128);
    setstate(live);
}

static double now(void) {
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec + 1e-9 * ts.tv_nsec;
}

struct slot_rec { double time, potential; int32_t partner, type, sub, pad; int64_t pcounter; };

int main(int argc, const char *argv[]) {
    if (argc < 5) { /* <passes> <slots> -- <at least one sDMD argument> */ fprintf(stderr, "usage: %s <passes> <slots.bin|-> -- <sDMD args>\n", argv[0]); return 2; }
    int passes = atoi(argv[1]);
    const char *slots = argv[2];
    const char *state = (argc > 3 && strcmp(argv[3], "--")) ? argv[3] : NULL;
    int a = 3;
    while (a < argc && strcmp(argv[a], "--")) a++;
    if (a >= argc) return 2;
    argv[a] = argv[0];
    double t0 = now();
    InputData(argc - a, argv + a);
    fprintf(stderr, "ref_bench: InputData %.2f s, %d atoms\n", now() - t0, atomnum);

    atomList.ptr = (struct AtomStr **)__real_calloc(atomnum + 1, sizeof(struct AtomStr *));
    for (int i = 1; i <= atomnum; i++) atomList.ptr[i] = &atom[i];
    int *renewAll = __real_calloc(atomnum + 2, sizeof(int));
    for (int i = 1; i <= atomnum; i++) renewAll[i] = i;
    if (state) {
        FILE *f = fopen(state, "wb");
        if (!f) { perror(state); return 1; }
        int64_t n64 = atomnum;
        double box[3] = {boxDimension[1], boxDimension[2], boxDimension[3]};
        fwrite(&n64, sizeof(n64), 1, f);
        fwrite(box, sizeof(double), 3, f);
        double *col = malloc(sizeof(double) * atomnum);
        for (int d = 1; d <= 3; d++) { for (int i = 1; i <= atomnum; i++) col[i - 1] = atom[i].dynamic->coordinate[d]; fwrite(col, sizeof(double), atomnum, f); }
        for (int d = 1; d <= 3; d++) { for (int i = 1; i <= atomnum; i++) col[i - 1] = atom[i].dynamic->velocity[d]; fwrite(col, sizeof(double), atomnum, f); }
        char rng0[128];
        rng_save(rng0);
        for (int i = 0; i < atomnum; i++) col[i] = Maxwell_Boltzmann_Distribution(0, 1);
        rng_restore(rng0);
        fwrite(col, sizeof(double), atomnum, f);
        free(col);
        fclose(f);
    }
    for (int k = 0; k < passes; k++) {
        if (CBT.node) { free(CBT.node); free(CBT.leaf); free(CBT.time); CBT.node = NULL; CBT.leaf = NULL; CBT.time = NULL; CBT.count = 0; }
        if (celllist) { free(celllist); celllist = NULL; }      /* LinkList("full") allocates a fresh one, Events.c:125 */
        double t1 = now();
        PBC("full", NULL, &atomList);
        LinkList("full", NULL, NULL, &atomList);
        Predict(renewAll, &atomList);
        CreateCBT();
        printf("pass %d seconds %.6f\n", k, now() - t1);
        fflush(stdout);
    }
    if (strcmp(slots, "-")) {
        FILE *f = fopen(slots, "wb");
        if (!f) { perror(slots); return 1; }
        for (int i = 1; i <= atomnum; i++) {
            struct InteractionEventStr *e = &atom[i].dynamic->event;
            struct slot_rec r = {e->time, e->potential, e->partner, e->eventType, e->subEventType, 0, e->partnerCounter};
            fwrite(&r, sizeof(r), 1, f);
        }
        fclose(f);
    }
    return 0;
}
