// sdmd_dev.cuh — device data model and device functions of the sDMD hot path for sm_100a.
//
// Everything here is exact IEEE binary64 in the reference's association order; the translation unit is
// compiled with -fmad=false so no multiply-add is contracted (the reference is x86-64 SSE2, no FMA).
// Reference citations are relative to the sDMD source tree (src/).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>
#include "sdmd_cuda.h"

#define SD_NT (SDMD_NATOMTYPE + 1)
#define SD_INF SDMD_INFINIT
#define SD_ZERO SDMD_ZERO

typedef long long i64;

// device-resident control block: the loop state the reference keeps in process globals (DMD.c:12-61)
struct Ctl {
    double currenttime;        // currenttime
    double timestep;           // timestep
    double temperature;        // targetTemperature
    double dt;                 // AssignJob: atomList.timeInr
    i64 frame, iterations, count_recal;
    i64 n_coli, n_bond, n_hb, n_hbne, n_ther, n_pcc, n_wall;
    i64 warnings, pair_predictions;
    i64 error_prints;          // "!!ERROR!!" prints the reference carries on from (not part of warningsum)
    i64 pa_passes, pa_rounds, pa_requeued;   // full-system passes: count, re-evaluation rounds, atoms queued by the resolve stage
    i64 g_pos;                 // gaussians consumed (absolute)
    i64 g_filled;              // gaussians fed (absolute)
    int target, partner;       // AssignJob result
    int stop;                  // 0 = running, else 1 + sdmd_stop_reason
    int error;
    int hbnumformed;           // HBnumformed
    int alphaformed;           // alphaHBformed (Events.c:310-311,363-364)
    int changed;               // predict_all fix-point flag
    i64 trace_pos;
    int pad_cc_;
    int live;                  // overlapped sweep: which of the two coordinate/time buffers holds the current state
    // candidate-list cache: cc_epoch is a tick bumped by every change of the link list or of the HB state; the change stamps
    // the cells it touched with the new tick (DevP::cell_epoch).  A cached list is valid while its stamp is at least
    // cc_floor (last wholesale invalidation: new link list) and no cell of its 27-cell neighbourhood carries a later tick.
    i64 cc_epoch, cc_floor;
    // moving walls (WallDyn): the box edges WallTime / PBCShift see, changed by DoWallDyn (host, once per time unit) and by
    // WallTime's own retry when an atom is found outside the wall (TimePrediction.c:675-697)
    double dbox[3];
    int wd_touch, wd_pad_;     // wallDyn.touch
    i64 phase_cycles[12];      // clock64() per loop phase (thread 0): head, advance, event, gather, eval, fold, cbt
    // overlapped sweep: AssignJob result per launch parity.  The event block of launch n publishes the job of launch
    // n+1 in the other slot while sweep blocks of launch n may still be reading theirs.
    struct Job { double dt; i64 seq; int target, partner, go, pad_; } job[2];   // seq = loop iteration the job belongs to
};

// thin view of "the slots other atoms currently hold" (live arrays in the event loop, the previous
// relaxation round in the full-system pass)
struct SlotSrc {
    const double *time, *pot;
    const int *partner;
    const i64 *pcounter;
    const uint8_t *type, *sub;
};

// one set of event-slot arrays (the live ones or a relaxation buffer of the full-system pass)
struct SlotBuf {
    double *time, *pot;
    int *partner;
    i64 *pcounter;
    uint8_t *type, *sub;
};

struct Slot {
    double time, pot;
    double m2;      // value of the best candidate that did NOT win (used by the full-system pass)
    double bp_time, sp_time;   // best / second-best PAIR candidate seen so far (full-system pass)
    int bp_j;
    i64 pcounter;
    int partner, type, sub;
};

// One cell-sorted atom as the evaluation stage of Predict(all) wants it: a survivor costs one 64-byte gather (two
// sectors, four 16-byte loads) instead of eight scattered 4/8-byte loads
struct __align__(16) SRec {
    double x, y, z, vx, vy, vz;
    int type, atom;
    unsigned cxy, cz;            // cell index: x | y << 16, z
};

// Obstruction walls, tunnels, spherical obstacles, charge sources: rarely present, so their tables live in device memory
// instead of the kernel parameter block (the sweep driver launches once per loop iteration and pays for every parameter byte)
struct SingleBody {
    int obst_type[SDMD_MAX_OBST], obst_hole[SDMD_MAX_OBST];
    double obst_pos[SDMD_MAX_OBST][3], sph_pos[SDMD_MAX_OBST][3], sph_r[SDMD_MAX_OBST];
    double tun_pos[SDMD_MAX_OBST][4], tun_diam[SDMD_MAX_OBST];
    double chg_pos[SDMD_MAX_CHARGE][3], chg_vel[SDMD_MAX_CHARGE][3], chg_grad[SDMD_MAX_CHARGE], chg_gap[SDMD_MAX_CHARGE];
};

struct DevP {
    int n, ncell;
    int cellnum[3];
    double cellsize[3], box[3];
    double cutoffr, cut2, cutfull2, thermoF;
    int thermostat_on, method_ding, hydrogen_bond, gel_mode;
    double hb_barrier[3];
    int wall_kind, wall_type;  // smooth confining wall (enum sdmd_wall_kind), its row of the collision table
    int pbc_mask;              // bit k: dimension k is periodic (Events.c:210-216)
    double wall_pos[3];
    int n_obst, n_sphobst;     // flat obstruction walls / hard spherical obstacles
    int flow_mark;             // 1 constant drift velocity, 2 constant force (ThermostatEvent / CalKinetE terms)
    double flow_v[3];
    const double *flow_a;      // 3 per atom or NULL
    double *flow_time_rec;     // per atom or NULL
    const double *protein_mass;
    const double *charge;      // per atom or NULL (charge sources: flow_mark == 3)
    int n_charge, charge_com;
    int n_tunnel, tunnel_type; // tunnels through (holed) obstruction walls
    int single_body;           // any of wall / obstruction / sphere present
    const struct SingleBody *sb;   // geometry tables of the single-body predictors (device memory; NULL when single_body == 0)
    double *x, *y, *z, *vx, *vy, *vz;
    const double *mass;
    int *type;
    const int *aa_type, *protein, *aa_num;
    const uint8_t *is_cons;
    int *hb_bond;
    const int *hb_neighbor;
    const uint8_t *hb_role;
    const uint8_t *hb_name;    // 1 = "H", 2 = "O" (CheckAlphaHB, Events.c:1121-1128), NULL: no alpha tally
    uint8_t *hb_itype;
    int *hbn_partner, *hbn_pnum, *hbn_status;
    double *ev_time, *ev_pot;
    int *ev_partner;
    i64 *ev_counter, *ev_pcounter;
    uint8_t *ev_type, *ev_sub;
    int *cell_idx;     // 3 per atom
    int *celllist;     // reference layout: next[1..n], heads at [n + cell], cell 1-based
    int *cbt_node, *cbt_leaf;
    const int *bond_off, *bond_j;
    const double *bond_dmin2, *bond_dmax2;
    const int *con_off, *con_j, *con_table;
    const int *cm_off, *cm_j, *cm_flags;
    const int *cm_table;   // constraint table of a CONSTRAINT_CONNECT entry (resolved from the constr list at create)
    int *cell_cnt;         // atoms per cell, kept in step with the link list
    // per-atom cache of the 27-cell candidate list (scan order); validity: see Ctl::cc_epoch
    i64 *cc_stamp;
    i64 *cell_epoch;       // per cell: tick of the last change of its membership or of an HB attribute of one of its atoms
    int *cc_n, *cc_j;
    unsigned char *cc_k, *cc_pass;
    int cc_cap;            // entries per row
    unsigned short *cc_passidx;   // positions (within the row) of entries that pass the free-HB filters, 32 per row
    int *cc_npass;
    const unsigned int *cm_dense; // (n+1)^2 entries flags | table << 8 for small systems, else NULL
    int cc_per_atom;       // 1: one row per atom, reused across events; 0: two scratch rows, rebuilt every event (huge systems)
    const sdmd_step_table *tables;
    const int *coll_index, *hb_index;
    const double *gauss;
    i64 g_cap;
    const int *ther_rank;  // index of the gaussian atom i consumes in a full pass
    int no_cons;           // no constrained atom at all: ther_rank[i] == i - 1
    Ctl *ctl;
    double *pe_terms;      // CalPotenE addends, 64 per atom (natom <= 2048 only, else NULL)
    int *pe_cnt;
    // cell-sorted tiles produced by cell_build (fresh link-list order: descending atom id inside a cell)
    const int *cell_start;   // [ncell + 1], 0-based cell
    const int *sorted_atom;  // [n]
    const struct SRec *srec;      // full record per sorted position: what the evaluation stage gathers
    // round-0 filter output: per sorted position up to R0_SURV_CAP candidates inside the scan sphere
    unsigned int *surv;      // (sorted position of j) | (scan slot k << 27)
    int surv_cap;            // entries per row: 16 for sparse bond-free boxes, else R0_SURV_CAP (chosen at create)
    int *surv_n;             // count, or -1 when the list overflowed (atom re-scanned in the evaluation kernel)
    int *surv_np;            // pair evaluations counted by the filter
    int *spos;               // sorted position of every atom (inverse of sorted_atom)
    double cut2_loose;       // cut2 * (1 + 1e-9): the filter keeps a superset, the evaluation re-tests exactly
    // single-precision working set of the neighbour filter: position relative to the atom's own cell corner (so a
    // neighbour cell is just an offset of whole cells, periodic wrap included) + the atom id; home cell per position
    const float4 *sloc;
    const int *scell;
    float cut2_f32;          // cut2 * (1 + 1e-4), rounded up: superset sphere of the single-precision test
    float cellsize_f[3];
    // round 0 as one tile kernel (sdmd_k2_tile.cu): own cells per thread block, halo capacity in atoms, dynamic shared memory
    int tile_tx, tile_ty, tile_tz, tile_hcap, tile_smem, tile_mode;
    float tile_cut2f;        // superset sphere of the tile kernel's single-precision test (tile-relative coordinates)
    // Overlapped sweep (large-system loop): while the grid-wide UpdateData/TimeForward sweep of event n is still
    // running, the event block already works on event n.  x/y/z/ev_time above are then the buffer the sweep WRITES;
    // lz_* is the buffer it reads (state before the sweep).  The event block reads every atom other than the event's
    // own two through d_rx/d_ry/d_rz/d_rt, which apply the sweep's arithmetic (x + v*dt, t - dt) on the fly — the same
    // two roundings, so the same bits; the event's own atoms are written to the new buffer up front and live there.
    const double *lz_x, *lz_y, *lz_z, *lz_t;
    double lz_dt;
    int lz_on, lz_a, lz_b;
    // round 0 by sdmd_k2_rows.cu: survivor rows are written in row order, not in visiting order (re-evaluation rounds sort them)
    int rows_mode;
    int cm_empty;            // the connection map has no entry at all (no bond, no constraint)
    int wall_dyn;            // moving walls: the box lives in Ctl::dbox (d_box)
    int wd_active;           // the motion is still on (wallDyn.mark != "no"): WallTime may push the box back out
    double wd_step, wd_orig_y;   // wallDyn.step, wallDyn.origBoxDim[2]
    int uni_table;           // >= 0: every atom has the same type; its collision table (index) has <= 2 steps and no dmax2.  Else -1
};

#define SDMD_LZ 0
namespace lz0 {
#include "sdmd_dev_body.cuh"
}
#undef SDMD_LZ
#define SDMD_LZ 1
namespace lz1 {
#include "sdmd_dev_body.cuh"
}
#undef SDMD_LZ
using namespace lz0;
