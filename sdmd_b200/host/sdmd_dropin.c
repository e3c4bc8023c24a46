/*
 * sdmd_dropin.c — the reference's driver with the hot path on the GPU.
 *
 * Compiled together with the UNMODIFIED reference sources (see sdmd_b200/host/Makefile).  The only change to the
 * reference build is two renames on the compiler command line: SGThread.c gets -DSingleThread=sdmd_reference_SingleThread
 * and REMD.c gets -DREMD=sdmd_reference_REMD, so that ThreadProcess() (ToolFunctions.c:12-20) reaches the
 * SingleThread() / REMD() defined here.  Everything around the hot path — CLI, parameter.txt / .gro parsing,
 * force-field library, SDEnergyMin, checkpoint reader, every output file, the REMD socket client and the sServer /
 * sREMD programs — is the reference's own code.
 *
 *   Prepare()           SGThread.c:21-64   PBC/LinkList/Predict(all)/CreateCBT -> sdmd_cuda_pbc_full / cell_build /
 *                                          predict_all / cbt_build ("continue": upload_celllist + cbt_build)
 *   SingleThreadRun()   SGThread.c:67-156  the while-loop incl. DoEvent        -> sdmd_cuda_run(ctx, next boundary)
 *   REMDRun()           REMD.c:29-126      + CalPotenE                          -> sdmd_cuda_potential_energy
 *                                          + ReplicaExchange (reference's socket client, unchanged)
 *   SaveData(...)                          unchanged, after sdmd_cuda_download + scatter
 */
#include "DMD.h"
#include "sdmd_refglue.h"

/* the reference's socket client, REMD.c:135-192 (external linkage there, no prototype in DMD.h) */
void ReplicaExchange(double *E, struct REMDTempStr *T, char *serverName, int *port_number);

static struct sdmd_host_view view;
static sdmd_ctx *ctx;
static long base_counts[8];

static void die(const char *what, int rc) {
    printf("!!ERROR!!: %s failed (%d): %s\n", what, rc, sdmd_cuda_last_error());
    exit(EXIT_FAILURE);
}

static void sync_to_reference(void) {
    int rc = sdmd_cuda_download(ctx, &view);
    if (rc) die("sdmd_cuda_download", rc);
    sdmd_glue_scatter_dynamic(&view);
    rc = sdmd_cuda_download_celllist(ctx, celllist);
    if (rc) die("sdmd_cuda_download_celllist", rc);
    struct sdmd_run_stats st;
    sdmd_cuda_get_stats(ctx, &st);
    frame = st.frame;
    currenttime = st.currenttime;
    countReCal = base_counts[0] + st.count_recal;
    collisioneventsum = base_counts[1] + st.collision;
    bondeventsum = base_counts[2] + st.bond;
    HBeventsum = base_counts[3] + st.hb;
    HBNeighboreventsum = base_counts[4] + st.hb_neighbor;
    thermostateventsum = base_counts[5] + st.thermostat;
    pbcandcrosseventsum = base_counts[6] + st.pcc;
    walleventsum = base_counts[7] + st.wall;
    if (strcmp(wallExist, "no")) {                /* moving wall: the box edges WallTime's retry may have changed, wallDyn.touch */
        double box[3];
        int32_t touch = 0;
        if (sdmd_cuda_wall_dyn_state(ctx, box, &touch) == 0 && view.wall_dyn) {
            for (int k = 0; k < 3; k++) boxDimension[k + 1] = box[k];
            wallDyn.touch = touch;
        }
    }
    HBnumformed = (int)st.hb_num_formed;          /* what SaveHB prints (-HBn) and SaveDAT stores, DataSave.c:461-470, 573-574 */
    alphaHBformed = (int)st.alpha_hb_formed;
}

/* Prepare(), SGThread.c:21-64 */
static void prepare_cuda(void) {
    char *extraName = REMDInfo.flag ? REMDInfo.REMD_ExtraName : NULL;
    InitializeFiles(extraName, fileList);
    SaveData(lgf, fileList);
    atomList.ptr = (struct AtomStr **)calloc(atomnum + 1, sizeof(struct AtomStr *));
    for (int i = 1; i <= atomnum; i++) atomList.ptr[i] = &atom[i];
    int fresh = strcmp("new", neworcontinue) == 0;
    if (fresh) {
        SaveData(sysInfo, fileList);
        if (strcmp(wallExist, "smooth") == 0) SaveData(wallInfo, fileList);   /* SGThread.c:37 */
        if (obstObj.mark) SaveData(obstInfo, fileList);                       /* SGThread.c:38 (CG: refused by the glue) */
#ifndef GEL
        SDEnergyMin(20000, &atomList);
#endif
    } else if (strcmp("continue", neworcontinue) == 0) {
        outputrecord = currenttime + outputrate;
    } else {
        printf("!!ERROR!!: the simulation configuration is invalid!\n");
        exit(EXIT_FAILURE);
    }
    base_counts[0] = countReCal; base_counts[1] = collisioneventsum; base_counts[2] = bondeventsum;
    base_counts[3] = HBeventsum; base_counts[4] = HBNeighboreventsum; base_counts[5] = thermostateventsum;
    base_counts[6] = pbcandcrosseventsum; base_counts[7] = walleventsum;

    int rc = sdmd_glue_flatten(&view);
    if (rc) die("sdmd_glue_flatten (feature outside the hot-path scope?)", rc);
    /* one replica per process; the REMD launcher places replica id on GPU id % G (sREMD_cuda.c) */
    const char *dev = getenv("SDMD_CUDA_DEVICE");
    rc = sdmd_cuda_create(&view, dev ? atoi(dev) : 0, &ctx);
    if (rc) die("sdmd_cuda_create", rc);
    int cn[3];
    double cs[3];
    sdmd_cuda_cell_grid(ctx, cn, cs);
    if (cn[0] != cellnum[1] || cn[1] != cellnum[2] || cn[2] != cellnum[3]) die("cell grid mismatch with CreateCell", -1);
    if (fresh) {
        INT_CALLOC(celllist, (atomnum + cellnum[1] * cellnum[2] * cellnum[3] + 1));
        if ((rc = sdmd_cuda_pbc_full(ctx))) die("pbc_full", rc);
        if ((rc = sdmd_cuda_cell_build(ctx))) die("cell_build", rc);
        if ((rc = sdmd_cuda_rng_fill_libc(ctx, atomnum))) die("rng_fill", rc);   /* ThermostatTime draws of Predict(all) */
        if ((rc = sdmd_cuda_predict_all(ctx))) die("predict_all", rc);
    } else {
        /* savedData.dat carries every atom's pending event and the link list (Initialization.c:360-509) */
        if ((rc = sdmd_cuda_upload_celllist(ctx, celllist))) die("upload_celllist", rc);
    }
    if ((rc = sdmd_cuda_cbt_build(ctx))) die("cbt_build", rc);
}

static void save_outputs(void) {
    sync_to_reference();
    for (int i = 0; i < lenFileType; i++)
        if (fileList[i].mark) SaveData(i, fileList);
    outputrecord += outputrate;
}

/* DoWallDyn() where SingleThreadRun calls it (SGThread.c:129): after the executed event that carried currenttime to wallDyn.curTime */
static int wall_dyn_mode(void) {
    return strcmp(wallDyn.mark, "compress") == 0 ? 1 : strcmp(wallDyn.mark, "expand") == 0 ? 2 : strcmp(wallDyn.mark, "pulse") == 0 ? 3 : 0;
}
static void wall_dyn_step(void) {
    struct sdmd_wall_dyn wd;
    memset(&wd, 0, sizeof(wd));
    wd.mode = wall_dyn_mode(); wd.sign = wallDyn.sign; wd.touch = wallDyn.touch;
    wd.size = wallDyn.size; wd.rate = wallDyn.rate; wd.step = wallDyn.step; wd.cur_time = wallDyn.curTime;
    for (int k = 0; k < 3; k++) wd.orig_box[k] = wallDyn.origBoxDim[k + 1];
    int rc = sdmd_cuda_wall_dyn_step(ctx, &wd);
    if (rc) die("sdmd_cuda_wall_dyn_step", rc);
    wallDyn.sign = wd.sign; wallDyn.touch = wd.touch; wallDyn.curTime = wd.cur_time;
    for (int k = 0; k < 3; k++) boxDimension[k + 1] = wd.box[k];
    if (wd.mode == 0) sprintf(wallDyn.mark, "no");
}

/* run to `boundary`; returns the stop reason once it is TIME_END or T_UNTIL */
static int run_to(double boundary, struct sdmd_run_stats *st, double *gap) {
    const int64_t chunk = 200000;        /* gaussians generated ahead per refill */
    for (;;) {
        int rc;
        if (sdmd_cuda_rng_available(ctx) < chunk / 2 && (rc = sdmd_cuda_rng_fill_libc(ctx, chunk))) die("rng_fill", rc);
        rc = sdmd_cuda_run(ctx, boundary, (int64_t)1 << 40, NULL, 0, st);
        if (rc) die("sdmd_cuda_run", rc);
        currenttime = st->currenttime;
        double processratio = (currenttime - oldcurrenttime) / (timestep - oldcurrenttime) * 100;
        if (processratio >= *gap + 0.01) { printf("Process=%8.2lf%%\r", processratio); fflush(stdout); *gap = processratio; }
        if (st->stop_reason == SDMD_STOP_NEED_RNG) continue;
        if (st->stop_reason == SDMD_STOP_T_UNTIL || st->stop_reason == SDMD_STOP_TIME_END) return st->stop_reason;
        die("unexpected stop", st->stop_reason);
    }
}

/* SingleThread(), SGThread.c:15-19 */
void SingleThread(void) {
    prepare_cuda();
    struct sdmd_run_stats st;
    memset(&st, 0, sizeof(st));
    double gap = 0;
    for (;;) {
        /* the next thing the host has to do: a dump (SGThread.c:132-139) or, under a moving wall, DoWallDyn (SGThread.c:129) */
        int dyn = strcmp(wallDyn.mark, "no") != 0;
        double boundary = (dyn && wallDyn.curTime < outputrecord) ? wallDyn.curTime : outputrecord;
        int reason = run_to(boundary, &st, &gap);
        if (dyn && currenttime >= wallDyn.curTime) wall_dyn_step();
        if (reason != SDMD_STOP_T_UNTIL) break;
        if (currenttime >= outputrecord) save_outputs();
    }
    sync_to_reference();
    warningsum += st.warnings;
    sdmd_cuda_destroy(ctx);
}

/* NCCL exchange (sREMD_cuda's default): the replicas of the run form one communicator, rank = replica id.  Rank 0 creates the
 * ncclUniqueId and publishes it through the file the launcher named; the others wait for it. */
static int nccl_mode(void) { const char *e = getenv("SDMD_REMD_NCCL"); return e && e[0] == '1'; }
static void nccl_join(void) {
    const char *sr = getenv("SDMD_REMD_NRANKS"), *sk = getenv("SDMD_REMD_RANK"), *path = getenv("SDMD_REMD_ID_FILE"), *ss = getenv("SDMD_REMD_SEED");
    if (!sr || !sk || !path || !ss) { printf("!!ERROR!!: SDMD_REMD_NCCL needs SDMD_REMD_NRANKS / _RANK / _ID_FILE / _SEED (set by sREMD_cuda)\n"); exit(EXIT_FAILURE); }
    int nranks = atoi(sr), rank = atoi(sk), rc;
    unsigned char id[128];
    if (rank == 0) {
        if ((rc = sdmd_remd_nccl_unique_id(id))) die("sdmd_remd_nccl_unique_id", rc);
        char tmp[1100];
        snprintf(tmp, sizeof(tmp), "%s.tmp", path);
        FILE *f = fopen(tmp, "wb");
        if (!f || fwrite(id, 1, 128, f) != 128) { printf("!!ERROR!!: cannot write %s\n", tmp); exit(EXIT_FAILURE); }
        fclose(f);
        rename(tmp, path);                                   /* atomic: readers never see a partial id */
    } else {
        for (int tries = 0;; tries++) {
            FILE *f = fopen(path, "rb");
            if (f) {
                size_t got = fread(id, 1, 128, f);
                fclose(f);
                if (got == 128) break;
            }
            if (tries > 6000) { printf("!!ERROR!!: no NCCL id from replica 0 in %s\n", path); exit(EXIT_FAILURE); }
            usleep(10000);
        }
    }
    if ((rc = sdmd_remd_nccl_init(ctx, nranks, rank, id))) die("sdmd_remd_nccl_init", rc);
    if ((rc = sdmd_remd_seed_server(ctx, (uint32_t)strtoul(ss, NULL, 10)))) die("sdmd_remd_seed_server", rc);
}

/* REMD(), REMD.c:24-126 */
void REMD(void) {
    prepare_cuda();
    const int nccl = nccl_mode();
    if (nccl) nccl_join();
    int outputRate = REMDInfo.REMD_OutputRate;
    int rate = ((int)(currenttime / outputRate) + 1) * outputRate;      /* REMD.c:32 */
    struct sdmd_run_stats st;
    memset(&st, 0, sizeof(st));
    double gap = 0;
    for (;;) {
        double boundary = outputrecord < rate ? outputrecord : (double)rate;
        int dyn = strcmp(wallDyn.mark, "no") != 0;                       /* moving wall: DoWallDyn, REMD.c:85 */
        if (dyn && wallDyn.curTime < boundary) boundary = wallDyn.curTime;
        int reason = run_to(boundary, &st, &gap);
        if (dyn && currenttime >= wallDyn.curTime) wall_dyn_step();
        if (reason != SDMD_STOP_T_UNTIL) break;
        if (currenttime >= outputrecord) save_outputs();                 /* REMD.c:88-95 */
        if (currenttime >= rate) {                                       /* REMD.c:97-104 */
            double potenEnergy;
            int rc;
            if (nccl) {
                /* CalPotenE + ReplicaExchange + the server's sweep in one call: energy computed on the device into the send
                 * record, ncclAllGather, replicated Server.c:145-192 sweep; the new temperature is set on the device inside */
                struct sdmd_remd_temp t = {REMDInfo.REMD_Temperature.num, 0, REMDInfo.REMD_Temperature.T};
                if ((rc = sdmd_remd_exchange_energy(ctx, &t, &potenEnergy))) die("sdmd_remd_exchange_energy", rc);
                REMDInfo.REMD_Temperature.num = t.num;
                REMDInfo.REMD_Temperature.T = t.T;
                targetTemperature = t.T;
            } else {
                if ((rc = sdmd_cuda_potential_energy(ctx, &potenEnergy))) die("potential_energy", rc);
                ReplicaExchange(&potenEnergy, &REMDInfo.REMD_Temperature, REMDInfo.REMD_ServerName, &REMDInfo.REMD_PortNum);
                targetTemperature = REMDInfo.REMD_Temperature.T;
                sdmd_cuda_set_temperature(ctx, targetTemperature);
            }
            SaveData(RE, fileList);
            rate += outputRate;
        }
    }
    sync_to_reference();
    warningsum += st.warnings;
    if (nccl) {                                                          /* the sentinel of REMD.c:119-122, as a collective */
        int rc = sdmd_remd_finish(ctx);
        if (rc) die("sdmd_remd_finish", rc);
    } else {
        tmpDouble = 0.0;
        REMDInfo.REMD_Temperature.T = 0.0;
        ReplicaExchange(&tmpDouble, &REMDInfo.REMD_Temperature, REMDInfo.REMD_ServerName, &REMDInfo.REMD_PortNum);
    }
    sdmd_cuda_destroy(ctx);
}
