/*
 * sdmd_cuda.h — C-ABI of libsdmd_cuda.so, the B200 (sm_100a) implementation of sDMD's
 * event-prediction / event-scheduling hot path.
 *
 * Plain C: pointers, sizes and PODs only.  Host code (the sDMD C driver, tests, bench) talks to the
 * CUDA side exclusively through the functions declared here.  Every entry point names the reference
 * interface it replaces (paths relative to the sDMD source tree, `src/`).
 *
 * Conventions
 *   - atom ids are 1-based exactly like the reference (`atom[1..atomnum]`); every per-atom array in
 *     `sdmd_host_view` has natom+1 elements and element 0 is unused.
 *   - distances in the step tables are squared (the reference squares them at load time,
 *     Initialization.c:1507,1515,1313-1330).
 *   - all functions return 0 on success or a negative SDMD_ERR_* code; nothing here ever falls back to
 *     a CPU implementation — when no CUDA device is usable the call fails.
 */
#ifndef SDMD_CUDA_H
#define SDMD_CUDA_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SDMD_ABI_VERSION 8

/* reference constants, DMD.h:25-30 */
#define SDMD_INFINIT       1E8
#define SDMD_ZERO          1E-8
#define SDMD_INVALID       (-111)
#define SDMD_NATOMTYPE     32
#define SDMD_EXTEND_RATIO  0.90
#define SDMD_MAX_STEPS     16     /* longest shipped table row has 13 steps */
#define SDMD_NHBCLASS      12     /* potentialPairHB[12][33][33], DMD.h:550 */
#define SDMD_MAX_OBST      4      /* flat obstruction walls / spherical obstacles the device keeps (each) */
#define SDMD_MAX_CHARGE    8      /* charge sources incl. their periodic images */

/* connection-map bits, DMD.h:59-63 */
#define SDMD_BOND_CONNECT        0x01
#define SDMD_CONSTRAINT_CONNECT  0x02
#define SDMD_HB_CONNECT          0x04
#define SDMD_NEIGHBOR_CONNECT    0x08

/* enum EEventType, DMD.h:162-179 (values are part of the log / checkpoint format) */
enum sdmd_event_type {
    SDMD_EV_COLI = 0, SDMD_EV_BOND = 1, SDMD_EV_HB = 2, SDMD_EV_HBNE = 3, SDMD_EV_THER = 4,
    SDMD_EV_WALL = 5, SDMD_EV_OBST = 6, SDMD_EV_TUNL = 7, SDMD_EV_CHRG = 8, SDMD_EV_SPHO = 9,
    SDMD_EV_PCC = 10, SDMD_EV_CGSU = 11, SDMD_EV_LAGV = 12, SDMD_EV_TBD = 13, SDMD_EV_CACL = 14,
    SDMD_EV_INVD = 15
};

/* enum EInteractionType, DMD.h:181-191 */
enum sdmd_sub_event {
    SDMD_SUB_CORECOLLI = 0, SDMD_SUB_WELLBOUNC = 1, SDMD_SUB_WELLCAPTR = 2, SDMD_SUB_WELLESCAP = 3,
    SDMD_SUB_TBDCAPTR = 4, SDMD_SUB_TBDESCAP = 5, SDMD_SUB_TBDHBFORM = 6, SDMD_SUB_TBDHBBRIK = 7,
    SDMD_SUB_INVALID = 8
};

/* enum EHBEventType, DMD.h:193-198 */
enum sdmd_hb_event { SDMD_HB_NOEVENT = 0, SDMD_HB_FORM = 1, SDMD_HB_BREAK = 2, SDMD_HB_SKIP = 3 };

enum sdmd_error {
    SDMD_OK = 0,
    SDMD_ERR_NO_DEVICE   = -1,   /* no usable CUDA device / kernel image (never a CPU fallback) */
    SDMD_ERR_CUDA        = -2,   /* a CUDA runtime call failed; see sdmd_cuda_last_error() */
    SDMD_ERR_BAD_ARG     = -3,
    SDMD_ERR_UNSUPPORTED = -4,   /* feature outside the hot-path scope (rough walls, moving walls, CG, obstacles, tunnels, flow) */
    SDMD_ERR_TABLE       = -5,   /* a pair needed a potential table that does not exist (FindPair exit, ToolFunctions.c:114-117) */
    SDMD_ERR_STATE       = -6,   /* device-side consistency check failed (link list / event type) */
    SDMD_ERR_NCCL        = -7
};

/* One flattened ConstraintStr + its StepPotenStr chain (DMD.h:244-257).
 * acc[k] is StepPotenStr.accumulated (ToolFunctions.c:897-919). */
struct sdmd_step_table {
    double dmin2;
    double dmax2;                     /* 0 for collision tables */
    int32_t nstep;
    int32_t pad_;
    double d2[SDMD_MAX_STEPS];
    double e[SDMD_MAX_STEPS];
    double acc[SDMD_MAX_STEPS];
};

/* Flattened, pointer-free view of the reference's process globals (DMD.h:475-613) that the hot path
 * reads.  The host keeps ownership of everything; sdmd_cuda_create copies what it needs. */
struct sdmd_host_view {
    int32_t abi_version;              /* SDMD_ABI_VERSION */
    int32_t natom;                    /* atomnum */
    double  box[3];                   /* boxDimension[1..3] */
    double  cutoffr;                  /* cutoffr */
    double  thermoF;                  /* Thermo_Freq */
    double  target_temperature;       /* targetTemperature (reduced) */
    double  currenttime;              /* currenttime */
    double  timestep;                 /* timestep = end of run */
    int32_t thermostat_on;            /* strcmp(thermostatType,"no") != 0 */
    int32_t method_ding;              /* strncmp(Methodtype,"Ding",1) == 0 */
    int32_t hydrogen_bond;            /* built with HYDROGEN_BOND (DMD.h:23) */
    int32_t gel_mode;                 /* built with GEL (no BondTime, SGThread.c:170) */
    int64_t frame;                    /* executed events so far */

    /* per atom, natom+1 entries */
    double *x, *y, *z;                /* dynamic->coordinate[1..3] */
    double *vx, *vy, *vz;             /* dynamic->velocity[1..3] */
    double *mass;                     /* property->mass */
    int32_t *type;                    /* property->typeofAtom (mutable: HB form/break) */
    int32_t *aa_type;                 /* property->typeofAA */
    int32_t *protein_num;             /* property->sequence.proteinNum */
    int32_t *aa_num;                  /* property->sequence.aminoacidNum */
    uint8_t *is_cons;                 /* extraProperty[1] == "_CONS" */
    int32_t *hb_bond;                 /* dynamic->HB.bondConnection */
    int32_t *hb_neighbor;             /* dynamic->HB.neighbor */
    uint8_t *hb_role;                 /* dynamic->HB.role ('A','D','S' or 0) */
    uint8_t *hb_itype;                /* dynamic->HB.interactionType */
    int32_t *hbn_partner;             /* dynamic->HBNeighbor.neighborPartner, 5 ints per atom */
    int32_t *hbn_partner_num;         /* dynamic->HBNeighbor.partnerNum */
    int32_t *hbn_status;              /* dynamic->HBNeighbor.neighborStatus */
    double  *ev_time;                 /* dynamic->event.* */
    double  *ev_potential;
    int32_t *ev_partner;
    int64_t *ev_counter;
    int64_t *ev_pcounter;
    uint8_t *ev_type;
    uint8_t *ev_sub;
    int32_t *cell_idx;                /* dynamic->cellIndex[1..3], 3 ints per atom; may be NULL before the first cell build */

    /* covalent / auxiliary bonds, property->bond in list order; CSR over atoms: bond_off[natom+2] */
    int32_t *bond_off;
    int32_t *bond_j;
    double  *bond_dmin2, *bond_dmax2;
    /* CONSTRn pairs, property->constr in list order; con_table indexes `tables` */
    int32_t *con_off;
    int32_t *con_j;
    int32_t *con_table;
    /* static bits of connectionMap (BOND / CONSTRAINT), CSR, row order ascending j */
    int32_t *cm_off;
    int32_t *cm_j;
    int32_t *cm_flags;

    /* potential tables */
    int32_t ntable;
    struct sdmd_step_table *tables;
    int32_t *coll_index;              /* [33*33] potentialPairCollision after RightPair resolution (ToolFunctions.c:278-292); -1 = none */
    int32_t *hb_index;                /* [12*33*33] potentialPairHB; -1 = empty entry */
    double  hb_barrier[3];            /* HBPotential.BB_v, SS_v, BS_v (Events.c:1097-1119) */

    /* smooth confining wall (WallExist smooth; ReadWall Initialization.c:1907-1935, WallTime TimePrediction.c:635-737,
     * WallEvent Events.c:423-499, PBC Events.c:210-216).  wall_kind = SDMD_WALL_NONE when wallExist is "no". */
    int32_t wall_kind;                /* enum sdmd_wall_kind */
    int32_t wall_type;                /* wall[0].property->typeofAtom: row of the collision table the wall interacts through */
    double  wall_pos[3];              /* wall[0].dynamic->coordinate[1..3]: the axis / mid-plane / centre the distance is taken from */

    /* flat obstruction walls without holes ("Obstruct?" line of parameter.txt, Initialization.c:637-676; pseudo-atoms
     * built by ReadWall :1981-2012; ObstTime TimePrediction.c:740-829, ObstEvent Events.c:543-619).  Event partner code
     * -dim - 10 n (TimePrediction.c:823). */
    int32_t n_obst;                   /* obstObj.num, <= SDMD_MAX_OBST */
    int32_t n_sphobst;                /* SphObstObj.num, <= SDMD_MAX_OBST */
    double  obst_pos[SDMD_MAX_OBST][3];   /* obstObj.position[n][1..3]; negative = no plane in that dimension */
    int32_t obst_type[SDMD_MAX_OBST];     /* obstObj.obst[n].property->typeofAtom */
    /* hard spherical obstacles ("SphObst?" line, Initialization.c:701-726; SphObstTime TimePrediction.c:1020-1071,
     * SphObstEvent Events.c:752-796).  Event partner code -1000 n. */
    double  sphobst_pos[SDMD_MAX_OBST][3];
    double  sphobst_radius[SDMD_MAX_OBST]; /* SphObstObj.radius[n] (the file gives the diameter) */

    /* tunnels through obstruction walls ("Tunnel?" line, Initialization.c:678-699; pseudo-atom built by ReadWall :2014-2049;
     * TunnelTime TimePrediction.c:832-930, TunnelEvent Events.c:621-682, EnterTunnel ToolFunctions.c:921-952).  Event partner
     * code -10000 n.  obst_hole[n] != 0: obstruction wall n sits at a tunnel mouth (obstObj.hole). */
    int32_t n_tunnel;                 /* tunlObj.num, <= SDMD_MAX_OBST */
    int32_t tunnel_type;              /* tunlObj.tunnel.property->typeofAtom */
    double  tunnel_pos[SDMD_MAX_OBST][4];  /* tunlObj.position[n][0..3]: x of the two mouths, then y, z of the axis */
    double  tunnel_diameter[SDMD_MAX_OBST];
    int32_t obst_hole[SDMD_MAX_OBST];

    /* flow ("Flow" line of parameter.txt, Initialization.c:562-633; CreateFlow :2056-2068): 1 = constant drift velocity,
     * 2 = constant force.  Both act where the Andersen thermostat redraws a velocity (ThermostatEvent, Events.c:396-420) and
     * in CalKinetE (ToolFunctions.c:626-656).  3 = charge sources (ChargeTime TimePrediction.c:933-1017, ChargeEvent
     * Events.c:685-749, partner code -100 n; with `-COM` ThermostatEvent adds the chain's centre-of-mass velocity). */
    int32_t flow_mark;                /* flow.mark */
    int32_t pad_flow_;
    double  flow_v[3];                /* flow.constV.v[1..3] */
    double *flow_a;                   /* flow.force.a[i][1..3], 3 per atom; may be NULL (= 0) */
    double *flow_time_rec;            /* flow.force.timeRec[i] (mutable); may be NULL when flow_mark == 0 */
    double *protein_mass;             /* per atom: protein[proteinNum].mass of its chain (CalculateMass, Initialization.c:1557-1567,
                                         counts hydrogens as 1); only read by CalKinetE when flow_mark >= 2; may be NULL otherwise */
    double *charge;                   /* property->charge per atom (InitializeCharge, Initialization.c:2092-2141); may be NULL */
    int32_t n_charge;                 /* charge sources ChargeTime loops over (its totalCNum, TimePrediction.c:936), <= SDMD_MAX_CHARGE */
    int32_t charge_com;               /* chargeAA[COM] (command-line -COM) */
    double  charge_pos[SDMD_MAX_CHARGE][3], charge_vel[SDMD_MAX_CHARGE][3];
    double  charge_potgrad[SDMD_MAX_CHARGE], charge_gap[SDMD_MAX_CHARGE];

    /* hydrogen-bond tallies the reference prints with -HBn and stores in savedData.dat (DataSave.c:461-470, 574): HBnumformed and
     * alphaHBformed (Events.c:296-311,352-389).  hb_name: 1 where property->name is "H", 2 where it is "O", else 0 — what
     * CheckAlphaHB (Events.c:1121-1128) looks at; may be NULL (no alpha tally). */
    uint8_t *hb_name;
    int32_t hb_num_formed;            /* HBnumformed at the start of the run */
    int32_t alpha_hb_formed;          /* alphaHBformed at the start of the run */

    /* moving smooth wall ("WallDyn compress | expand | pulse", Initialization.c:1937-1978): the wall surface follows
     * boxDimension, which DoWallDyn (Events.c:502-540, sdmd_cuda_wall_dyn_step) changes once per time unit and WallTime's retry
     * (TimePrediction.c:675-697) pushes back out when it finds an atom outside.  wall_dyn = 0: static walls. */
    int32_t wall_dyn;                 /* strcmp(wallDyn.mark, "no") != 0 */
    int32_t wall_dyn_touch;           /* wallDyn.touch */
    double  wall_dyn_step;            /* wallDyn.step */
    double  wall_dyn_orig_box[3];     /* wallDyn.origBoxDim[1..3] */
};

/* wallType (parameter.txt line 11), first letter decides (strncmp(...,1)): which coordinates take part in the wall distance
 * and which dimensions stay periodic */
enum sdmd_wall_kind {
    SDMD_WALL_NONE = 0,
    SDMD_WALL_PARALLEL = 1,           /* two plates normal to y: distance from y = box_y/2; x and z periodic */
    SDMD_WALL_CYLINDER = 2,           /* tube along x: distance from the axis (y,z) = box/2; x periodic */
    SDMD_WALL_SPHERE   = 3            /* cavity centred in the box; nothing periodic */
};

/* counters mirroring the reference's global tallies (DMD.c:50-61) */
struct sdmd_run_stats {
    int64_t frame;                    /* committed events (frame, SGThread.c:100) */
    int64_t iterations;               /* loop iterations incl. hazard-denied */
    int64_t count_recal;              /* countReCal */
    int64_t collision, bond, hb, hb_neighbor, thermostat, pcc;
    int64_t warnings;                 /* warningsum: FindPair's too-small / too-large warnings (ToolFunctions.c:127-139,171-178,203-210);
                                       * the too-small one is suppressed while a tunnel exists, like the reference's */
    int64_t pair_predictions;         /* CalculateParameters-equivalent evaluations */
    int64_t gaussians_used;           /* thermostat RNG samples consumed */
    double  currenttime;
    int32_t stop_reason;              /* enum sdmd_stop_reason */
    int32_t error;                    /* first device-side SDMD_ERR_* or 0 */
    int64_t wall;                     /* walleventsum: wall, obstruction and spherical-obstacle events (Events.c:73-100) */
    int64_t error_prints;             /* "!!ERROR!!" messages the reference would have printed and carried on from (negative
                                       * predicted time, bond long-/short-error: TimePrediction.c:168-173,240-269); NOT in warningsum */
    int64_t hb_num_formed;            /* HBnumformed: hydrogen bonds in existence now (Events.c:345,384) */
    int64_t alpha_hb_formed;          /* alphaHBformed: those between backbone H and O four residues apart (CheckAlphaHB, Events.c:1121-1128) */
};

enum sdmd_stop_reason {
    SDMD_STOP_TIME_END   = 0,         /* currenttime > timestep (loop condition SGThread.c:75) */
    SDMD_STOP_T_UNTIL    = 1,         /* an executed event carried currenttime past t_until (output / exchange boundary) */
    SDMD_STOP_MAX_ITER   = 2,
    SDMD_STOP_NEED_RNG   = 3,         /* gaussian ring ran low; refill and call again */
    SDMD_STOP_ERROR      = 4
};

/* one line of the DEBUG_PRINTF event log (SGThread.c:87-99) plus the time increment */
struct sdmd_trace_rec {
    int64_t frame;
    int32_t target;
    int32_t partner;                  /* target's event.partner (may be SDMD_INVALID) */
    int32_t type;
    int32_t sub;
    int32_t executed;                 /* 1 executed, 0 denied */
    int32_t pad_;
    double  dt;
};

typedef struct sdmd_ctx sdmd_ctx;

/* ---- lifetime --------------------------------------------------------------------------------- */
/* Flatten-and-upload.  Replaces nothing in the reference; it is what a maintainer calls once after
 * InputData() (DMD.c:196).  Computes the cell grid exactly like CreateCell (Initialization.c:1060-1073). */
int  sdmd_cuda_create(const struct sdmd_host_view *view, int device, sdmd_ctx **out);
void sdmd_cuda_destroy(sdmd_ctx *ctx);
const char *sdmd_cuda_last_error(void);
int  sdmd_cuda_device_count(void);

/* CreateCell result: cellnum[1..3], cellsize[1..3] */
int  sdmd_cuda_cell_grid(const sdmd_ctx *ctx, int cellnum[3], double cellsize[3]);

/* ---- state transfer (host buffers <-> HBM) ------------------------------------------------------ */
/* positions / velocities only: natom+1 doubles each, entry 0 ignored */
int  sdmd_cuda_upload_coords(sdmd_ctx *ctx, const double *x, const double *y, const double *z,
                             const double *vx, const double *vy, const double *vz);
/* restore a saved link list (celllist[natom + ncell + 1]); the resume path of ReadBreakPoint,
 * Initialization.c:360-509, where the list order inside a cell is part of the state */
int  sdmd_cuda_upload_celllist(sdmd_ctx *ctx, const int32_t *celllist);
/* refresh every mutable per-atom array of `view` from the device (what SaveData/DoEvent need,
 * SURVEY 8b "sync_to_host"); arrays that are NULL in `view` are skipped */
int  sdmd_cuda_download(sdmd_ctx *ctx, struct sdmd_host_view *view);
/* event slots only (any pointer may be NULL); natom+1 entries each */
int  sdmd_cuda_download_slots(sdmd_ctx *ctx, double *time, double *potential, int32_t *partner, uint8_t *type, uint8_t *sub);
/* reference-format link list: celllist[natom + ncell + 1] (Events.c:125,152-153) */
int  sdmd_cuda_download_celllist(sdmd_ctx *ctx, int32_t *celllist);
/* CBT.node[2*natom], CBT.leaf[natom+1] (CBT.c:16-30) */
int  sdmd_cuda_download_cbt(sdmd_ctx *ctx, int32_t *node, int32_t *leaf);

/* ---- thermostat random stream ------------------------------------------------------------------- */
/* Append standard-normal samples, in the order the reference would draw them from
 * Maxwell_Boltzmann_Distribution (ToolFunctions.c:544-569).  The device consumes them sequentially. */
int  sdmd_cuda_rng_feed(sdmd_ctx *ctx, const double *gauss, int64_t n);
/* Generate n samples on the host with libc rand()/log()/sqrt() exactly as ToolFunctions.c:517-569 does
 * and feed them.  Uses the process-global rand() stream the reference seeded with srand(seed). */
int  sdmd_cuda_rng_fill_libc(sdmd_ctx *ctx, int64_t n);
int64_t sdmd_cuda_rng_available(const sdmd_ctx *ctx);
/* move the consumption cursor back/forward inside the fed window (benchmarks replay one synthetic stream) */
int  sdmd_cuda_rng_seek(sdmd_ctx *ctx, int64_t pos);
/* drop everything fed so far and start a fresh stream */
int  sdmd_cuda_rng_reset(sdmd_ctx *ctx);
/* the host-side generator by itself (no ctx): one draw of Maxwell_Boltzmann_Distribution(0,1) */
double sdmd_host_gaussian_libc(void);

/* ---- the hot path, one call per reference function ----------------------------------------------- */
/* PBC("full") Events.c:218-228 */
int  sdmd_cuda_pbc_full(sdmd_ctx *ctx);
/* LinkList("full") Events.c:121-156 — GPU counting sort; also produces the cell-sorted SoA tiles */
int  sdmd_cuda_cell_build(sdmd_ctx *ctx);
/* Prepare()'s Predict over all atoms, SGThread.c:47-52 → Predict SGThread.c:158-211.
 * Consumes one gaussian per non-constrained atom when the thermostat is on (constrained atoms skip ThermostatTime,
 * SGThread.c:165-168). */
int  sdmd_cuda_predict_all(sdmd_ctx *ctx);
/* Predict(renewList) SGThread.c:158-211; renew_list follows the reference layout:
 * [0] = count, [1..count] = atom ids, 0-terminated */
int  sdmd_cuda_predict(sdmd_ctx *ctx, const int32_t *renew_list);
/* CreateCBT CBT.c:16-30 / UpdateCBT CBT.c:33-39 */
int  sdmd_cuda_cbt_build(sdmd_ctx *ctx);
/* Prepare() SGThread.c:44-62 from HOST inputs in one call: upload coordinates and a fresh thermostat stream
 * (gauss[0..n_gauss), n_gauss >= natom when the thermostat is on), then PBC("full"), LinkList("full"), Predict(all),
 * CreateCBT.  Same results as upload_coords + rng_reset + rng_feed + pbc_full + cell_build + predict_all + cbt_build;
 * velocities and gaussians travel on a second stream while the cell list and the neighbour filter already run.
 * Arrays are 1-based (natom+1 entries); page-locked host memory (sdmd_host_register) makes the copies asynchronous. */
int  sdmd_cuda_prepare(sdmd_ctx *ctx, const double *x, const double *y, const double *z, const double *vx,
                       const double *vy, const double *vz, const double *gauss, int64_t n_gauss);
/* Prepare() SGThread.c:44-62 on the state that is already on the device: PBC("full"), LinkList("full"), Predict(all),
 * CreateCBT in one call (PBC rides along with the cell-index kernel; the thermostat stream continues where it stands).
 * Same results as pbc_full + cell_build + predict_all + cbt_build. */
int  sdmd_cuda_prepare_resident(sdmd_ctx *ctx);
/* DoWallDyn Events.c:502-540, called by the host loop where the reference calls it (SGThread.c:129, after an executed event) once
 * currenttime has reached wd->cur_time: advances the box edges by one step (pulse: turns around first where the reference does),
 * bumps the event counter of every atom whose pending event is a wall event, ends the motion ("compress" / "expand" reached).
 * In/out: the reference's wallDyn fields; `box` returns the box edges afterwards.  sdmd_cuda_wall_dyn_state reads the edges and the
 * retry tally as WallTime left them (they change during sdmd_cuda_run). */
struct sdmd_wall_dyn {
    int32_t mode;                     /* 0 "no", 1 "compress", 2 "expand", 3 "pulse" */
    int32_t sign, touch, pad_;
    double  size, rate, step, cur_time;
    double  orig_box[3];
    double  box[3];
};
int  sdmd_cuda_wall_dyn_step(sdmd_ctx *ctx, struct sdmd_wall_dyn *wd);
int  sdmd_cuda_wall_dyn_state(sdmd_ctx *ctx, double box[3], int32_t *touch);
int  sdmd_cuda_cbt_update(sdmd_ctx *ctx, const int32_t *renew_list);
/* AssignJob SGThread.c:237-259: target = CBT.node[1], dt = its event time, partner (0 if none) */
int  sdmd_cuda_assign_job(sdmd_ctx *ctx, int32_t *target, int32_t *partner, double *dt, int32_t *renew_list);
/* UpdateData + TimeForward, ToolFunctions.c:23-36 (one fused sweep) */
int  sdmd_cuda_advance(sdmd_ctx *ctx, double dt);
/* currenttime += timeInr, SGThread.c:80 (for callers that drive the loop one call at a time) */
int  sdmd_cuda_add_time(sdmd_ctx *ctx, double dt);
/* HazardCheck SGThread.c:215-235; *hazard = 0 ok / -1 denied */
int  sdmd_cuda_hazard_check(sdmd_ctx *ctx, int32_t target, int32_t partner, int32_t *hazard);
/* PBCandCrossCellEvent Events.c:248-272 = PBC("part") + LinkList("part") for one atom */
int  sdmd_cuda_cell_move(sdmd_ctx *ctx, int32_t atom);
/* DoEvent Events.c:25-114 for the pending (target, partner) of the last assign_job, on the device */
int  sdmd_cuda_do_event(sdmd_ctx *ctx);
/* CalPotenE ToolFunctions.c:659-752 / CalKinetE :626-656 */
int  sdmd_cuda_potential_energy(sdmd_ctx *ctx, double *energy);
int  sdmd_cuda_kinetic_energy(sdmd_ctx *ctx, double *energy);

/* SingleThreadRun / REMDRun event loop (SGThread.c:67-156, REMD.c:29-126) resident on the device:
 * runs until currenttime > timestep, or an executed event reaches t_until, or max_iter iterations.
 * trace (optional, capacity trace_cap records) receives one record per loop iteration. */
int  sdmd_cuda_run(sdmd_ctx *ctx, double t_until, int64_t max_iter,
                   struct sdmd_trace_rec *trace, int64_t trace_cap, struct sdmd_run_stats *stats);
/* Many replicas on one GPU (REMD ensembles: the reference runs one sDMD process per replica, REMDExec.c:111-138):
 * the same loop for n contexts concurrently, one persistent thread block each, one launch.  stats: n entries. */
int  sdmd_cuda_run_many(sdmd_ctx **ctxs, int n, const double *t_until, int64_t max_iter, struct sdmd_run_stats *stats);
int  sdmd_cuda_potential_energy_many(sdmd_ctx **ctxs, int n, double *energy);
int  sdmd_cuda_get_stats(sdmd_ctx *ctx, struct sdmd_run_stats *stats);
/* targetTemperature after a replica exchange (REMD.c:100) */
int  sdmd_cuda_set_temperature(sdmd_ctx *ctx, double t_reduced);

/* ---- REMD exchange (REMD.c:135-192 client + Server.c:121-211 server) ----------------------------- */
struct sdmd_remd_temp { int32_t num; int32_t pad_; double T; };   /* struct REMDTempStr, DMD.h:377-380 */

/* The sequential Metropolis sweep of Server.c:145-192 over nrep replicas (host arrays, in place).
 * `u01` supplies the uniform draws of RANDOMVALUE(0,1) (Server.c:16-18) in order; *n_used returns how
 * many were consumed.  accept/attempt (size nrep, may be NULL) accumulate Server.c's `probability`. */
int  sdmd_remd_rearrange(double *epot, struct sdmd_remd_temp *T, int nrep,
                         const double *u01, int n_u01, int *n_used,
                         double *attempt, double *accept);

/* NCCL communicator for the exchange step: one rank per GPU.  `unique_id` is the 128-byte
 * ncclUniqueId produced by rank 0 (sdmd_remd_nccl_unique_id) and distributed by the launcher. */
int  sdmd_remd_nccl_unique_id(unsigned char id[128]);
int  sdmd_remd_nccl_init(sdmd_ctx *ctx, int nranks, int rank, const unsigned char id[128]);
/* srand(seed) of sServer (Server.c:57-58): seeds the replicated glibc rand() stream every rank advances
 * identically during the exchange sweep */
int  sdmd_remd_seed_server(sdmd_ctx *ctx, uint32_t seed);
/* n outputs of glibc's rand() after srand(seed), from the library's restatement of the TYPE_3 generator */
int  sdmd_host_glibc_rand(uint32_t seed, int64_t n, uint32_t *out);
/* ReplicaExchange: all-gather (E, T, num) of every rank over NCCL, run the replicated sweep with the
 * shared server seed stream, return this rank's new (T, num). */
int  sdmd_remd_exchange(sdmd_ctx *ctx, double epot, struct sdmd_remd_temp *T, uint32_t reserved);
/* The same with CalPotenE fused in (REMD.c:98 + 135-192 in one call): the energy is computed on the device and written
 * straight into the all-gather's send record — no device->host->device hop in front of the collective.  natom <= 2048.
 * *epot (may be NULL) returns the energy this replica contributed.  Both calls fail with SDMD_ERR_STATE when
 * sdmd_remd_seed_server has not been called for the context. */
int  sdmd_remd_exchange_energy(sdmd_ctx *ctx, struct sdmd_remd_temp *T, double *epot);
/* End of this replica's run: the (E = 0, T = 0) sentinel of REMD.c:119-122.  The reference's server goes on serving the replicas
 * that are still running and skips the ones that left (Server.c:108-113,125); a collective needs every rank in every round, so
 * this call keeps contributing the sentinel — one all-gather per round of the others — until every replica has finished.
 * Call it once after the last sdmd_cuda_run of a replica that took part in sdmd_remd_exchange*. */
int  sdmd_remd_finish(sdmd_ctx *ctx);

/* ---- measurement helpers ------------------------------------------------------------------------- */
/* average device time (ms) of the last launch of each named kernel, measured with CUDA events on the
 * context's stream; name in {"advance","predict_all","cell_build","cbt_build"}; <0 if never launched */
double sdmd_cuda_kernel_ms(const sdmd_ctx *ctx, const char *name);
/* accumulated device time and launch count of a named kernel group since the last reset:
 * "advance", "predict_all" (round 0), "predict_all_fixup", "cell_build", "cbt_build", "run" */
int  sdmd_cuda_kernel_stats(const sdmd_ctx *ctx, const char *name, double *total_ms, int64_t *count);
void sdmd_cuda_reset_kernel_stats(sdmd_ctx *ctx);
/* device-time bracket (CUDA events on the context's stream) around any sequence of calls */
int  sdmd_cuda_timer_begin(sdmd_ctx *ctx);
int  sdmd_cuda_timer_end(sdmd_ctx *ctx, double *ms);
int  sdmd_cuda_synchronize(sdmd_ctx *ctx);
/* page-lock caller-owned host buffers used with upload_* / download_* / rng_feed */
int  sdmd_cuda_host_register(void *ptr, size_t bytes);
int  sdmd_cuda_host_unregister(void *ptr);
/* How sdmd_cuda_run drives the loop: AUTO = one persistent thread block for natom <= 2048, CLUSTER = one persistent
 * thread-block cluster (block 0 runs the events, the others help with UpdateData/TimeForward) up to 98304, otherwise SWEEP = ONE
 * cooperative launch for the whole run, one block per SM: block 0 executes the events, the others sweep
 * UpdateData/TimeForward over all atoms side by side with it (the event block applies the sweep's arithmetic on the fly to
 * what it reads), a grid-wide barrier per iteration.  SWEEP_LAUNCH is the same overlap with one launch per iteration,
 * SWEEP_SERIAL runs the sweep and the event block one after the other.  Results are identical; tests force all of them. */
enum sdmd_loop_mode { SDMD_LOOP_AUTO = 0, SDMD_LOOP_ONE_CTA = 1, SDMD_LOOP_SWEEP = 2, SDMD_LOOP_SWEEP_SERIAL = 3, SDMD_LOOP_CLUSTER = 4,
                      SDMD_LOOP_SWEEP_LAUNCH = 5 };
int  sdmd_cuda_set_loop_mode(sdmd_ctx *ctx, int mode);
void sdmd_cuda_set_timing(sdmd_ctx *ctx, int on);   /* CUDA-event brackets per stage (asynchronous, resolved on query): 0 off, 1 default, 2 also predict_all's fix-up sub-stages */
int64_t sdmd_cuda_launch_count(const sdmd_ctx *ctx);
/* debug: accumulated SM cycles per event-loop phase {head, advance, event, gather, eval, fold, cbt, iterations} */
int  sdmd_cuda_phase_cycles(sdmd_ctx *ctx, int64_t out[12]);
void *sdmd_cuda_stream(const sdmd_ctx *ctx);   /* cudaStream_t */

#ifdef __cplusplus
}
#endif
#endif /* SDMD_CUDA_H */
