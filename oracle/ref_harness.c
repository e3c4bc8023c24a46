/*
 * ref_harness.c — TEST INFRASTRUCTURE.  Golden-vector generator that links the UNMODIFIED sDMD
 * reference objects (compiled by oracle/Makefile straight from the reference tree into oracle/_ref/)
 * and records what the reference itself computes on the hot path:
 *
 *   <prefix>_pre.snap    state after PBC("full") + LinkList("full"), before the first Predict
 *                        (SGThread.c:44-45) + the thermostat gaussian stream the run will consume
 *   <prefix>_post.snap   every atom's event slot after Predict(all) (SGThread.c:47-52), the link list
 *                        (Events.c:121-156), CBT.node / CBT.leaf after CreateCBT (CBT.c:16-30), CalPotenE
 *   <prefix>_final.snap  full state after `max_iter` loop iterations + one trace record per iteration
 *
 * The event loop below calls the reference's own AssignJob / UpdateData / TimeForward / HazardCheck /
 * DoEvent / Predict / UpdateCBT in the order of SingleThreadRun (SGThread.c:75-152) — only the file
 * output and progress printing are left out.  `ref_harness ... --printf-log` prints the same lines as a
 * -DDEBUG_PRINTF build so the loop can be diffed against the stock binary.
 *
 * With `--warm N` the harness first runs N loop iterations silently, then takes the "pre" snapshot from
 * that mid-run state (slots, link list and counters included), rebuilds the CBT from the stored times the
 * way a `continue` run does (Initialization.c:360-509 + CreateCBT), and traces from there.
 *
 * usage: ref_harness <out_prefix> <max_iter> <n_gauss> [--printf-log] [--no-sdmin] [--warm N] -- <sDMD args...>
 */
#include "DMD.h"
#include "sdmd_refglue.h"

/* ---- glibc random() state save / restore (rand() shares random()'s state) ---- */
/* -std=c99 hides these POSIX prototypes; declare them rather than change the reference's flags */
extern char *initstate(unsigned int seed, char *state, size_t n);
extern char *setstate(char *state);
static char rng_scratch[128] __attribute__((aligned(8)));
static void rng_save(char saved[128]) {
    char *sys = initstate(1, rng_scratch, 128); /* switches away; returns the live state with position packed in */
    memcpy(saved, sys, 128);
    setstate(sys);                              /* switch back */
}
static void rng_restore(const char saved[128]) {
    static char live[128] __attribute__((aligned(8)));
    memcpy(live, saved, 128);
    setstate(live);
}

static void put_cbt_and_cells(sdmd_snap *s) {
    int ncell = cellnum[1] * cellnum[2] * cellnum[3];
    sdmd_snap_put(s, "celllist", "i4", celllist, (uint64_t)atomnum + ncell + 1);
    int cn[3] = {cellnum[1], cellnum[2], cellnum[3]};
    double cs[3] = {cellsize[1], cellsize[2], cellsize[3]};
    sdmd_snap_put(s, "cellnum", "i4", cn, 3);
    sdmd_snap_put(s, "cellsize", "f8", cs, 3);
    if (CBT.node) {
        sdmd_snap_put(s, "cbt_node", "i4", CBT.node, (uint64_t)atomnum * 2);
        sdmd_snap_put(s, "cbt_leaf", "i4", CBT.leaf, (uint64_t)atomnum + 1);
    }
}

/* SingleThreadRun(), SGThread.c:67-156, without file output / progress printing */
static long run_loop(long max_iter, struct sdmd_trace_rec *trace, int printf_log) {
    int renewList[256] = {0};
    struct AtomListStr *thisList = &atomList;
    struct AtomStr *target, *partner;
    long iter = 0;
    while (currenttime <= timestep && iter < max_iter) {
        AssignJob(&target, &partner, renewList, thisList);
        UpdateData(thisList->timeInr);
        TimeForward(thisList->timeInr);
        currenttime += thisList->timeInr;
        int hazard = HazardCheck(target, partner, thisList->ptr[target->dynamic->HB.neighbor],
                                 (partner != NULL ? thisList->ptr[partner->dynamic->HB.neighbor] : NULL), thisList);
        if (trace) {
            struct sdmd_trace_rec *r = &trace[iter];
            r->frame = frame;
            r->target = target->property->num;
            r->partner = target->dynamic->event.partner;
            r->type = target->dynamic->event.eventType;
            r->sub = target->dynamic->event.subEventType;
            r->dt = thisList->timeInr;
            r->executed = !hazard;
        }
        iter++;
        if (printf_log)
            printf("frame = %4li, target atom = %5i, partner = %5i, %6s:%6s, type = %2i", frame, target->property->num,
                   target->dynamic->event.partner, target->property->name,
                   partner == NULL ? "NULL" : partner->property->name, target->dynamic->event.eventType);
        if (!hazard) {
            if (printf_log) printf(", executed!\n");
            frame++;
            DoEvent(thisList);
            target->dynamic->event.counter++;
            if (renewList[0] == 2 && strcmp(partner->property->extraProperty[1], "_CONS")) partner->dynamic->event.counter++;
            Predict(renewList, thisList);
            UpdateCBT(renewList);
            if (strcmp(wallDyn.mark, "no")) DoWallDyn();                                  /* SGThread.c:129 */
        } else {
            if (printf_log) printf(", denied, self/partner changed!\n");
            renewList[0] = 1;
            renewList[2] = 0;
            renewList[3] = 0;
            Predict(renewList, thisList);
            UpdateCBT(renewList);
            countReCal++;
        }
    }
    return iter;
}

void DoWallDyn(void);
/* moving wall: wallDyn.* and the current box edges next to the view */
static void put_wall_dyn(sdmd_snap *s) {
    int32_t wi[3] = {strcmp(wallDyn.mark, "compress") == 0 ? 1 : strcmp(wallDyn.mark, "expand") == 0 ? 2 : strcmp(wallDyn.mark, "pulse") == 0 ? 3 : 0,
                     wallDyn.sign, wallDyn.touch};
    double wf[10] = {wallDyn.size, wallDyn.rate, wallDyn.step, wallDyn.curTime, wallDyn.origBoxDim[1], wallDyn.origBoxDim[2], wallDyn.origBoxDim[3],
                     boxDimension[1], boxDimension[2], boxDimension[3]};
    sdmd_snap_put(s, "wall_dyn_i4", "i4", wi, 3);
    sdmd_snap_put(s, "wall_dyn_f8", "f8", wf, 10);
}

int main(int argc, const char *argv[]) {
    if (argc < 5) {
        fprintf(stderr, "usage: %s <out_prefix> <max_iter> <n_gauss> [--printf-log] -- <sDMD args>\n", argv[0]);
        return 2;
    }
    const char *prefix = argv[1];
    long max_iter = atol(argv[2]);
    long n_gauss = atol(argv[3]);
    int printf_log = 0, skip_sdmin = 0;
    long warm = 0;
    int a = 4;
    for (; a < argc && strcmp(argv[a], "--"); a++) {
        if (!strcmp(argv[a], "--printf-log")) printf_log = 1;
        if (!strcmp(argv[a], "--no-sdmin")) skip_sdmin = 1;
        if (!strcmp(argv[a], "--warm") && a + 1 < argc) warm = atol(argv[++a]);
    }
    if (a >= argc) { fprintf(stderr, "missing -- before sDMD args\n"); return 2; }
    argv[a] = argv[0];
    InputData(argc - a, argv + a);

    char path[1024];
    struct sdmd_host_view view;

    /* ---- Prepare(), SGThread.c:21-64, minus file output ---- */
    atomList.ptr = (struct AtomStr **)calloc(atomnum + 1, sizeof(struct AtomStr *));
    for (int i = 1; i <= atomnum; i++) atomList.ptr[i] = &atom[i];
#ifndef GEL
    if (!skip_sdmin) SDEnergyMin(20000, &atomList);
#endif
    PBC("full", NULL, &atomList);
    LinkList("full", NULL, NULL, &atomList);

    if (warm > 0) {
        /* run `warm` iterations first, then restart the bookkeeping from that state */
        int *renewAll = calloc(atomnum + 2, sizeof(int));
        for (int i = 1; i <= atomnum; i++) renewAll[i] = i;
        Predict(renewAll, &atomList);
        free(renewAll);
        CreateCBT();
        run_loop(warm, NULL, 0);
        free(CBT.node); free(CBT.leaf); free(CBT.time);
        CBT.node = NULL; CBT.leaf = NULL; CBT.count = 0;
    }

    int rc = sdmd_glue_flatten(&view);
    if (rc) { fprintf(stderr, "flatten failed: %d\n", rc); return 1; }

    /* pre-draw the gaussian stream with the reference's own generator, then rewind */
    char rng0[128];
    rng_save(rng0);
    double *gauss = malloc(sizeof(double) * (n_gauss > 0 ? n_gauss : 1));
    for (long k = 0; k < n_gauss; k++) gauss[k] = Maxwell_Boltzmann_Distribution(0, 1);
    rng_restore(rng0);

    int32_t warm_flag = warm > 0;
    snprintf(path, sizeof(path), "%s_pre.snap", prefix);
    sdmd_snap *s = sdmd_snap_open(path);
    sdmd_snap_put_view(s, &view);
    put_cbt_and_cells(s);
    put_wall_dyn(s);
    sdmd_snap_put(s, "gauss", "f8", gauss, n_gauss);
    sdmd_snap_put(s, "warm", "i4", &warm_flag, 1);
    sdmd_snap_close(s);

    if (!warm) {
        int *renewAll = calloc(atomnum + 2, sizeof(int));
        for (int i = 1; i <= atomnum; i++) renewAll[i] = i;
        Predict(renewAll, &atomList);
        free(renewAll);
    }
    CreateCBT();

    sdmd_glue_gather_dynamic(&view);
    snprintf(path, sizeof(path), "%s_post.snap", prefix);
    s = sdmd_snap_open(path);
    sdmd_snap_put_view(s, &view);
    put_cbt_and_cells(s);
    double epot = CalPotenE(), ekin = CalKinetE();
    sdmd_snap_put(s, "epot", "f8", &epot, 1);
    sdmd_snap_put(s, "ekin", "f8", &ekin, 1);
    sdmd_snap_close(s);

    long frame0 = frame, recal0 = countReCal;
    long base[7] = {collisioneventsum, bondeventsum, HBeventsum, HBNeighboreventsum, thermostateventsum, pbcandcrosseventsum,
                    walleventsum};
    struct sdmd_trace_rec *trace = calloc(max_iter > 0 ? max_iter : 1, sizeof(*trace));
    long iter = run_loop(max_iter, trace, printf_log);

    sdmd_glue_gather_dynamic(&view);
    snprintf(path, sizeof(path), "%s_final.snap", prefix);
    s = sdmd_snap_open(path);
    sdmd_snap_put_view(s, &view);
    put_cbt_and_cells(s);
    put_wall_dyn(s);
    sdmd_snap_put(s, "trace", "u1", trace, (uint64_t)iter * sizeof(*trace));
    int64_t stats[12] = {frame, iter, countReCal - recal0, collisioneventsum - base[0], bondeventsum - base[1],
                         HBeventsum - base[2], HBNeighboreventsum - base[3], thermostateventsum - base[4],
                         pbcandcrosseventsum - base[5], warningsum, frame0, walleventsum - base[6]};
    sdmd_snap_put(s, "stats", "i8", stats, 12);
    epot = CalPotenE(); ekin = CalKinetE();
    sdmd_snap_put(s, "epot", "f8", &epot, 1);
    sdmd_snap_put(s, "ekin", "f8", &ekin, 1);
    sdmd_snap_close(s);

    fprintf(stderr, "ref_harness: natom=%d iterations=%ld frame=%ld denied=%ld t=%.6f epot=%.6f\n", atomnum, iter, frame,
            countReCal, currenttime, epot);
    return 0;
}
