/*
 * sdmd_dropin.c — the reference's driver with the hot path on the GPU.
 *
 * Compiled together with the UNMODIFIED reference sources (see sdmd_b200/host/Makefile): the only change to the
 * reference build is `-DSingleThread=sdmd_reference_SingleThread` on SGThread.c, so that ThreadProcess()
 * (ToolFunctions.c:12-20) reaches the SingleThread() defined here instead.  Everything around the hot path — CLI,
 * parameter.txt / .gro parsing, force-field library, SDEnergyMin, every output file — is the reference's own code.
 *
 * SingleThread() below = Prepare() + SingleThreadRun() (SGThread.c:15-156) with
 *   PBC("full") / LinkList("full") / Predict(all) / CreateCBT        -> sdmd_cuda_pbc_full / cell_build / predict_all / cbt_build
 *   the while-loop (AssignJob ... UpdateCBT, DoEvent included)        -> sdmd_cuda_run(ctx, outputrecord, ...)
 *   SaveData(...) every outputrate                                    -> unchanged, after sdmd_cuda_download + scatter
 */
#include "DMD.h"
#include "sdmd_refglue.h"

static void die(const char *what, int rc) {
    printf("!!ERROR!!: %s failed (%d): %s\n", what, rc, sdmd_cuda_last_error());
    exit(EXIT_FAILURE);
}

static void sync_to_reference(sdmd_ctx *ctx, struct sdmd_host_view *v) {
    int rc = sdmd_cuda_download(ctx, v);
    if (rc) die("sdmd_cuda_download", rc);
    sdmd_glue_scatter_dynamic(v);
    rc = sdmd_cuda_download_celllist(ctx, celllist);
    if (rc) die("sdmd_cuda_download_celllist", rc);
    struct sdmd_run_stats st;
    sdmd_cuda_get_stats(ctx, &st);
    frame = st.frame;
    currenttime = st.currenttime;
    countReCal = st.count_recal;
    collisioneventsum = st.collision; bondeventsum = st.bond; HBeventsum = st.hb; HBNeighboreventsum = st.hb_neighbor;
    thermostateventsum = st.thermostat; pbcandcrosseventsum = st.pcc;
}

void SingleThread(void) {
    /* ---- Prepare(), SGThread.c:21-64 ---- */
    char *extraName = REMDInfo.flag ? REMDInfo.REMD_ExtraName : NULL;
    InitializeFiles(extraName, fileList);
    SaveData(lgf, fileList);
    atomList.ptr = (struct AtomStr **)calloc(atomnum + 1, sizeof(struct AtomStr *));
    for (int i = 1; i <= atomnum; i++) atomList.ptr[i] = &atom[i];
    if (strcmp("new", neworcontinue) != 0) {
        printf("!!ERROR!!: the CUDA drop-in starts \"new\" runs only\n");
        exit(EXIT_FAILURE);
    }
    SaveData(sysInfo, fileList);
#ifndef GEL
    SDEnergyMin(20000, &atomList);
#endif

    struct sdmd_host_view view;
    int rc = sdmd_glue_flatten(&view);
    if (rc) die("sdmd_glue_flatten (feature outside the hot-path scope?)", rc);
    sdmd_ctx *ctx = NULL;
    rc = sdmd_cuda_create(&view, 0, &ctx);
    if (rc) die("sdmd_cuda_create", rc);
    int cn[3];
    double cs[3];
    sdmd_cuda_cell_grid(ctx, cn, cs);
    if (cn[0] != cellnum[1] || cn[1] != cellnum[2] || cn[2] != cellnum[3]) die("cell grid mismatch with CreateCell", -1);
    INT_CALLOC(celllist, (atomnum + cellnum[1] * cellnum[2] * cellnum[3] + 1));

    if ((rc = sdmd_cuda_pbc_full(ctx))) die("pbc_full", rc);
    if ((rc = sdmd_cuda_cell_build(ctx))) die("cell_build", rc);
    if ((rc = sdmd_cuda_rng_fill_libc(ctx, atomnum))) die("rng_fill", rc);   /* ThermostatTime draws of Predict(all) */
    if ((rc = sdmd_cuda_predict_all(ctx))) die("predict_all", rc);
    if ((rc = sdmd_cuda_cbt_build(ctx))) die("cbt_build", rc);

    /* ---- SingleThreadRun(), SGThread.c:67-156 ---- */
    const int64_t chunk = 200000;        /* gaussians generated ahead per refill */
    struct sdmd_run_stats st;
    memset(&st, 0, sizeof(st));
    double gap = 0;
    for (;;) {
        if (sdmd_cuda_rng_available(ctx) < chunk / 2 && (rc = sdmd_cuda_rng_fill_libc(ctx, chunk))) die("rng_fill", rc);
        rc = sdmd_cuda_run(ctx, outputrecord, (int64_t)1 << 40, NULL, 0, &st);
        if (rc) die("sdmd_cuda_run", rc);
        currenttime = st.currenttime;
        double processratio = (currenttime - oldcurrenttime) / (timestep - oldcurrenttime) * 100;
        if (processratio >= gap + 0.01) { printf("Process=%8.2lf%%\r", processratio); fflush(stdout); gap = processratio; }
        if (st.stop_reason == SDMD_STOP_T_UNTIL) {      /* currenttime >= outputrecord, SGThread.c:132-139 */
            sync_to_reference(ctx, &view);
            for (int i = 0; i < lenFileType; i++)
                if (fileList[i].mark) SaveData(i, fileList);
            outputrecord += outputrate;
        } else if (st.stop_reason == SDMD_STOP_TIME_END) {
            break;
        } else if (st.stop_reason != SDMD_STOP_NEED_RNG) {
            die("unexpected stop", st.stop_reason);
        }
    }
    sync_to_reference(ctx, &view);
    warningsum += st.warnings;
    sdmd_cuda_destroy(ctx);
}
