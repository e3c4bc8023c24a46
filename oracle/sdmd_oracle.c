/*
 * sdmd_oracle.c — TEST INFRASTRUCTURE, NOT PRODUCT CODE (see sdmd_oracle.h).
 *
 * CPU restatement, in plain C, of the sDMD hot path.  Every function names the reference lines it
 * follows (paths relative to the reference's src/).  Arithmetic is IEEE binary64 with the reference's
 * association order and no contraction (build with -ffp-contract=off; the reference is x86-64 SSE2 -O2).
 *
 * Known, deliberate deviations (none reachable on the shipped inputs):
 *  - FindElemInList's unbounded scan (ToolFunctions.c:1078-1084) can run past neighborPartner[4] into the
 *    next struct members; here the scan stops at index 4.
 *  - fatal paths of the reference (exit(EXIT_FAILURE)) set stats.error instead of exiting.
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "sdmd_oracle.h"

#define NT (SDMD_NATOMTYPE + 1)
#define INF SDMD_INFINIT
#define ZERO SDMD_ZERO

/* neighbour-cell offsets, DMD.c:118-120 */
static const int OFF_X[27] = {0, 1, 0, 0, 0, 1, 1, 1, -1, 0, 0, 0, -1, -1, -1, 0, 0, 1, -1, 1, -1, -1, -1, 1, 1, 1, -1};
static const int OFF_Y[27] = {0, 0, 1, 0, 1, 0, 1, 1, 0, -1, 0, -1, 0, -1, -1, -1, 1, 0, 0, -1, 1, -1, 1, -1, 1, -1, 1};
static const int OFF_Z[27] = {0, 0, 0, 1, 1, 1, 0, 1, 0, 0, -1, -1, -1, 0, -1, 1, -1, -1, 1, 0, 0, 1, -1, -1, -1, 1, 1};

enum { FP_COLLISION, FP_CONSTRAINT, FP_HB, FP_NEIGHBOR };

struct oracle_ctx {
    struct sdmd_host_view *v;
    int n;
    int cellnum[3];
    double cellsize[3];
    int ncell;
    int32_t *celllist;
    int32_t *node, *leaf;
    int cbt_count;
    double *gauss;
    int64_t g_n, g_cap, g_pos;
    struct sdmd_run_stats st;
    int hbnumformed;
    struct sdmd_wall_dyn wd; /* moving wall (wallDyn.*); wd.mode == 0: none */
    int target, partner; /* AssignJob result */
    double time_inr;
};

struct evt { double time, potential; int sub; };
struct par {
    int coreorShell, s_time;
    double r_ij[3], v_ij[3], b_ij, v_2, r_2, d_ij2;
    double lowerPotential, upperPotential, shortlimit2, longlimit2;
};

/* ------------------------------------------------------------------------------------------------ */
oracle_ctx *oracle_create(struct sdmd_host_view *v) {
    oracle_ctx *c = calloc(1, sizeof(*c));
    c->v = v;
    c->n = v->natom;
    /* CreateCell, Initialization.c:1060-1073 */
    for (int k = 0; k < 3; k++) {
        c->cellnum[k] = (int)floor(v->box[k] / v->cutoffr);
        if (c->cellnum[k] < 3) c->cellnum[k] = 3;
    }
    for (int k = 0; k < 3; k++) c->cellsize[k] = v->box[k] / c->cellnum[k];
    c->ncell = c->cellnum[0] * c->cellnum[1] * c->cellnum[2];
    c->celllist = calloc((size_t)c->n + c->ncell + 1, sizeof(int32_t));
    c->st.currenttime = v->currenttime;
    c->st.frame = v->frame;
    c->hbnumformed = v->hb_num_formed;
    c->st.hb_num_formed = v->hb_num_formed;
    c->st.alpha_hb_formed = v->alpha_hb_formed;
    return c;
}

void oracle_destroy(oracle_ctx *c) {
    if (!c) return;
    free(c->celllist); free(c->node); free(c->leaf); free(c->gauss); free(c);
}

void oracle_rng_feed(oracle_ctx *c, const double *g, int64_t n) {
    if (c->g_n + n > c->g_cap) {
        c->g_cap = (c->g_n + n) * 2;
        c->gauss = realloc(c->gauss, sizeof(double) * c->g_cap);
    }
    memcpy(c->gauss + c->g_n, g, sizeof(double) * n);
    c->g_n += n;
}

static double next_gauss(oracle_ctx *c) {
    if (c->g_pos >= c->g_n) { c->st.error = SDMD_ERR_STATE; return 1.0; }
    c->st.gaussians_used++;
    return c->gauss[c->g_pos++];
}

void oracle_cell_grid(const oracle_ctx *c, int cellnum[3], double cellsize[3]) {
    for (int k = 0; k < 3; k++) { cellnum[k] = c->cellnum[k]; cellsize[k] = c->cellsize[k]; }
}
const int32_t *oracle_celllist(const oracle_ctx *c) { return c->celllist; }
void oracle_set_celllist(oracle_ctx *c, const int32_t *l) { memcpy(c->celllist, l, sizeof(int32_t) * ((size_t)c->n + c->ncell + 1)); }
const int32_t *oracle_cbt_node(const oracle_ctx *c) { return c->node; }
const int32_t *oracle_cbt_leaf(const oracle_ctx *c) { return c->leaf; }
void oracle_get_stats(const oracle_ctx *c, struct sdmd_run_stats *s) { *s = c->st; }
void oracle_set_temperature(oracle_ctx *c, double t) { c->v->target_temperature = t; }

/* ------------------------------------------------------------------------------------------------ */
/* which dimensions PBC wraps, Events.c:207-216: all three without a wall; x and z between parallel plates
 * (dim = 3, step = 2); x only inside a cylinder (dim = 1); none inside a sphere (early return) */
static int pbc_dim_mask(const struct sdmd_host_view *v) {
    switch (v->wall_kind) {
        case SDMD_WALL_PARALLEL: return 1 | 4;
        case SDMD_WALL_CYLINDER: return 1;
        case SDMD_WALL_SPHERE: return 0;
        default: return 7;
    }
}

/* PBC("full"), Events.c:218-228 */
void oracle_pbc_full(oracle_ctx *c) {
    struct sdmd_host_view *v = c->v;
    double *co[3] = {v->x, v->y, v->z};
    int mask = pbc_dim_mask(v);
    for (int i = 1; i <= c->n; i++)
        for (int k = 0; k < 3; k++) {
            if (!(mask >> k & 1)) continue;
            if (co[k][i] >= v->box[k]) co[k][i] -= v->box[k];
            else if (co[k][i] < 0) co[k][i] += v->box[k];
        }
}

static int cell_of(oracle_ctx *c, int i, int idx[3]) {
    struct sdmd_host_view *v = c->v;
    double co[3] = {v->x[i], v->y[i], v->z[i]};
    for (int k = 0; k < 3; k++) {
        idx[k] = (int)(co[k] / c->cellsize[k]); /* Events.c:132-133 */
        if (idx[k] == c->cellnum[k]) idx[k]--;  /* Events.c:143-145 */
    }
    return idx[2] * c->cellnum[0] * c->cellnum[1] + idx[1] * c->cellnum[0] + idx[0] + 1;
}

/* LinkList("full"), Events.c:121-156 */
void oracle_cell_build(oracle_ctx *c) {
    memset(c->celllist, 0, sizeof(int32_t) * ((size_t)c->n + c->ncell + 1));
    for (int i = 1; i <= c->n; i++) {
        int idx[3];
        int cell = cell_of(c, i, idx);
        for (int k = 0; k < 3; k++) c->v->cell_idx[3 * i + k] = idx[k];
        c->celllist[i] = c->celllist[c->n + cell];
        c->celllist[c->n + cell] = i;
    }
}

/* LinkList("part"), Events.c:158-199 */
static void cell_move(oracle_ctx *c, int i, const int old[3]) {
    int idx[3];
    int oldCell = old[2] * c->cellnum[0] * c->cellnum[1] + old[1] * c->cellnum[0] + old[0] + 1;
    int newCell = cell_of(c, i, idx);
    for (int k = 0; k < 3; k++) c->v->cell_idx[3 * i + k] = idx[k];
    int next = oldCell + c->n;
    if (c->celllist[next] == 0) c->st.error = SDMD_ERR_STATE;
    while (c->celllist[next] != i) {
        next = c->celllist[next];
        if (next == 0) { c->st.error = SDMD_ERR_STATE; return; }
    }
    c->celllist[next] = c->celllist[i];
    next = newCell + c->n;
    c->celllist[i] = c->celllist[next];
    c->celllist[next] = i;
}

/* ------------------------------------------------------------------------------------------------ */
/* AtomTypeChange, ToolFunctions.c:233-270 */
static int type_change(int t, int direct) {
    if (direct == 1) { if (t == 1) return 3; if (t == 2) return 4; if (t == 26) return 27; }
    else { if (t == 3) return 1; if (t == 4) return 2; if (t == 27) return 26; }
    return t;
}

/* HBModel, Models.c:212-290; -1 = "no HB" (ASN/GLN), -2 = the reference's fatal exit */
static int hb_model(const struct sdmd_host_view *v, int a1, int a2) {
    if (v->hb_role[a1] == 'A' || v->hb_role[a1] == 'S') { int s = a1; a1 = a2; a2 = s; }
    int o1 = type_change(v->type[a1], 0), o2 = type_change(v->type[a2], 0);
    int aa = v->aa_type[a2];
    if (o1 == 2 && o2 == 26) return 1;
    if (o1 == 2) {
        switch (aa) {
            case 4: case 7: return 2;
            case 16: case 17: case 19: return 3;
            case 9: return 4;
            case 5: return 10;
            case 3: case 6: return -1;
            default: break;
        }
    } else if (o2 == 26) return 5;
    else if (o1 == 31 && o2 == 31) return 9;
    else if (o1 == 1) {
        switch (aa) {
            case 4: case 7: return 6;
            case 16: case 17: case 19: return 7;
            case 9: return 8;
            case 3: case 6: return 11;
            default: break;
        }
    }
    return -2;
}

static int cm_flags(const struct sdmd_host_view *v, int i, int j) {
    for (int k = v->cm_off[i]; k < v->cm_off[i + 1]; k++)
        if (v->cm_j[k] == j) return v->cm_flags[k];
    return 0;
}

/* NEIGHBOR_CONNECT bit of connectionMap[a][b]: set exactly while b sits in a's neighborPartner list
 * (HBEvent flips both together, Events.c:297-329,350-382) */
static int neighbor_connected(const struct sdmd_host_view *v, int a, int b) {
    for (int k = 1; k <= 4; k++) {
        if (v->hbn_partner[5 * a + k] == 0) break;
        if (v->hbn_partner[5 * a + k] == b) return 1;
    }
    return 0;
}

/* FindElemInList(list, value, -1, -1, 0), ToolFunctions.c:1078-1084 */
static int find_in_hbn(const struct sdmd_host_view *v, int a, int value) {
    for (int k = 1; k <= 4; k++) {
        if (v->hbn_partner[5 * a + k] == 0) return 0;
        if (v->hbn_partner[5 * a + k] == value) return k;
    }
    return 0;
}

/* FindPair, ToolFunctions.c:39-224.  Output pointers may alias exactly like at the reference call sites;
 * the write order below is the reference's. */
static int find_pair(oracle_ctx *c, int a1, int a2, int mode, double direction, double dist2, double *lowerLimit,
                     double *upperLimit, double *lowerPot, double *upperPot, double *accum, int hbPartner, int typeChange) {
    const struct sdmd_host_view *v = c->v;
    int coreorShell = -1;
    int idx = -1;
    direction = (direction < 0) ? 1 : -1;        /* :46-50 */
    double tol = direction * ZERO;
    if (mode == FP_COLLISION) {                    /* :53-66 */
        int t1 = v->type[a1];
        if (typeChange == 1) t1 = type_change(t1, 1);
        else if (typeChange == -1) t1 = type_change(t1, 0);
        idx = v->coll_index[t1 * NT + (a2 == -1 ? v->wall_type : a2 == -100 ? v->tunnel_type : a2 < -1 ? v->obst_type[-a2 - 2] : v->type[a2])];   /* -1: wall pseudo-atom, -2-n: obstruction n, -100: tunnel */
    } else if (mode == FP_CONSTRAINT) {            /* :68-77 */
        for (int k = v->con_off[a1]; k < v->con_off[a1 + 1]; k++)
            if (v->con_j[k] == a2) { idx = v->con_table[k]; break; }
    } else if (mode == FP_HB) {                    /* :79-82 */
        int cls = hb_model(v, a1, a2);
        if (cls >= 0) idx = v->hb_index[(cls * NT + 0) * NT + 0];
    } else {                                       /* "neighbor" :84-90 */
        int cls = hbPartner > 0 ? hb_model(v, a2, hbPartner) : -2;
        int t1 = type_change(v->type[a1], 0), t2 = type_change(v->type[a2], 0);
        if (cls >= 0) idx = v->hb_index[(cls * NT + t1) * NT + t2];
    }
    if (idx < 0 || v->tables[idx].dmin2 == 0) {    /* :114-117 fatal in the reference */
        c->st.error = SDMD_ERR_TABLE;
        *lowerLimit = INF; *upperLimit = INF; *accum = INF; *lowerPot = 0; *upperPot = 0;
        return -1;
    }
    const struct sdmd_step_table *t = &v->tables[idx];
    int noStep = (t->nstep == 0);

    if (dist2 < t->dmin2 - ZERO) {                 /* :119-139 */
        *lowerLimit = INF; *upperLimit = INF; *accum = INF; *lowerPot = 0; *upperPot = 0;
        /* :127-130: counted only while wallDyn is "no" (always, here) and no tunnel exists */
        if (!(mode == FP_NEIGHBOR && v->hb_bond[a2] == 0) && !(mode == FP_HB && v->hb_bond[a1] == 0) && !v->n_tunnel) c->st.warnings++;
        return coreorShell;
    }
    *lowerLimit = t->dmin2;                        /* :141-142 */
    *lowerPot = 0;
    for (int k = 0; k < t->nstep; k++) {           /* :144-169 */
        if (dist2 < t->d2[k] + tol) {
            *upperLimit = t->d2[k];
            *upperPot = -1 * t->e[k];
            *accum = t->acc[k];
            if (k == 0) return 0;
            if (mode == FP_HB && k == t->nstep - 1 && v->hb_bond[a1] > 0) return 1;
            if (mode == FP_NEIGHBOR && k == t->nstep - 1 && v->hb_bond[a2] > 0) { *upperPot = 0; return 1; }
            return coreorShell;
        }
        *lowerLimit = t->d2[k];
        *lowerPot = t->e[k];
    }
    if (mode == FP_HB && v->hb_bond[a1] > 0) c->st.warnings++;   /* :171-178 */
    if (t->dmax2 > 0) {                            /* :180-211 */
        if (dist2 < t->dmax2 + ZERO) {
            *upperLimit = t->dmax2; *upperPot = 0; *accum = 0;
            coreorShell = 1;
            if (noStep && direction == 1) coreorShell = 0;
            return coreorShell;
        }
        *lowerLimit = INF; *upperLimit = INF; *accum = INF; *lowerPot = 0; *upperPot = 0;
        if (!(mode == FP_NEIGHBOR && v->hb_bond[a2] == 0) && !(mode == FP_HB && v->hb_bond[a1] == 0)) c->st.warnings++;
    } else {                                       /* :212-221 */
        *upperLimit = INF; *upperPot = INF; *accum = INF;
        return 1;
    }
    return coreorShell;
}

/* PBCShift, ToolFunctions.c:801-817 */
static void pbc_shift(const oracle_ctx *c, int i, int j, double shift[3]) {
    const int32_t *ci = &c->v->cell_idx[3 * i], *cj = &c->v->cell_idx[3 * j];
    for (int k = 0; k < 3; k++) {
        shift[k] = 0;
        if (ci[k] != cj[k]) {
            if (ci[k] == 0 && cj[k] == c->cellnum[k] - 1) shift[k] = -1 * c->v->box[k];
            else if (ci[k] == c->cellnum[k] - 1 && cj[k] == 0) shift[k] = c->v->box[k];
        }
    }
}

/* CalculateParameters, TimePrediction.c:1081-1090 */
static void calc_params(oracle_ctx *c, int i, int j, struct par *p) {
    const struct sdmd_host_view *v = c->v;
    double shift[3], rt[3];
    pbc_shift(c, i, j, shift);
    rt[0] = v->x[j] + shift[0]; rt[1] = v->y[j] + shift[1]; rt[2] = v->z[j] + shift[2];
    p->r_ij[0] = v->x[i] - rt[0]; p->r_ij[1] = v->y[i] - rt[1]; p->r_ij[2] = v->z[i] - rt[2];
    p->v_ij[0] = v->vx[i] - v->vx[j]; p->v_ij[1] = v->vy[i] - v->vy[j]; p->v_ij[2] = v->vz[i] - v->vz[j];
    p->r_2 = p->r_ij[0] * p->r_ij[0] + p->r_ij[1] * p->r_ij[1] + p->r_ij[2] * p->r_ij[2];
    p->b_ij = p->r_ij[0] * p->v_ij[0] + p->r_ij[1] * p->v_ij[1] + p->r_ij[2] * p->v_ij[2];
    p->v_2 = p->v_ij[0] * p->v_ij[0] + p->v_ij[1] * p->v_ij[1] + p->v_ij[2] * p->v_ij[2];
    c->st.pair_predictions++;
}
/* CalculateDisc :1092-1094, CalculateTime :1096-1099 */
static double calc_disc(const struct par *p) { return p->b_ij * p->b_ij - p->v_2 * (p->r_2 - p->shortlimit2); }
static double calc_time(const struct par *p) {
    return (-1 * p->b_ij + p->s_time * sqrt(p->b_ij * p->b_ij - p->v_2 * (p->r_2 - p->d_ij2))) / p->v_2;
}

/* CheckDuplication, TimePrediction.c:1101-1117 */
static int check_dup(oracle_ctx *c, int i, int j, int type) {
    struct sdmd_host_view *v = c->v;
    if (v->ev_partner[j] == i && v->ev_type[j] == type && v->ev_pcounter[j] == v->ev_counter[i] && v->ev_time[i] > v->ev_time[j]) {
        v->ev_partner[i] = j;
        v->ev_potential[i] = v->ev_potential[j];
        v->ev_sub[i] = v->ev_sub[j];
        v->ev_time[i] = v->ev_time[j];
        v->ev_type[i] = (uint8_t)type;
        v->ev_pcounter[i] = v->ev_counter[j];
        return 1;
    }
    return 0;
}

/* JobAssign, TimePrediction.c:1140-1158 (j == 0: no partner) */
static int job_assign(oracle_ctx *c, int i, int j, const struct evt *e, int type) {
    struct sdmd_host_view *v = c->v;
    if (e->time > 0 && v->ev_time[i] > e->time) {
        v->ev_potential[i] = e->potential;
        v->ev_sub[i] = (uint8_t)e->sub;
        v->ev_time[i] = e->time;
        v->ev_type[i] = (uint8_t)type;
        if (j) { v->ev_partner[i] = j; v->ev_pcounter[i] = v->ev_counter[j]; }
        else v->ev_partner[i] = SDMD_INVALID;
        return 1;
    }
    return 0;
}
static void init_event(struct evt *e) { e->time = INF; e->potential = 0; e->sub = SDMD_SUB_INVALID; } /* :1074-1079 */

static int neighbor_cell(const oracle_ctx *c, int i, int k) {
    const int32_t *ci = &c->v->cell_idx[3 * i];
    int nb[3] = {ci[0] + OFF_X[k], ci[1] + OFF_Y[k], ci[2] + OFF_Z[k]};
    for (int d = 0; d < 3; d++) {               /* TimePrediction.c:66-72 */
        if (nb[d] < 0) nb[d] = c->cellnum[d] - 1;
        else if (nb[d] >= c->cellnum[d]) nb[d] = 0;
    }
    return nb[2] * c->cellnum[0] * c->cellnum[1] + nb[1] * c->cellnum[0] + nb[0] + 1;
}

/* CollisionTime, TimePrediction.c:44-183 */
static void collision_time(oracle_ctx *c, int i) {
    struct sdmd_host_view *v = c->v;
    struct evt ev; struct par p; double tmpDouble;
    init_event(&ev);
    double cut2 = v->cutoffr * SDMD_EXTEND_RATIO;
    cut2 *= cut2;
    for (int k = 0; k < 27; k++) {
        int cell = neighbor_cell(c, i, k);
        for (int j = c->celllist[c->n + cell]; j; j = c->celllist[j]) {
            if (i == j) continue;
            if (check_dup(c, i, j, SDMD_EV_COLI)) continue;
            int cm = cm_flags(v, i, j);
            if ((cm & SDMD_BOND_CONNECT) || v->hb_bond[i] == j || find_in_hbn(v, i, j)) continue;
            calc_params(c, i, j, &p);
            if (!(p.r_2 <= cut2)) continue;
            if (cm & SDMD_CONSTRAINT_CONNECT) {
                p.coreorShell = find_pair(c, i, j, FP_CONSTRAINT, p.b_ij, p.r_2, &p.shortlimit2, &p.longlimit2,
                                          &p.lowerPotential, &p.upperPotential, &tmpDouble, 0, 0);
            } else {
                p.coreorShell = find_pair(c, i, j, FP_COLLISION, p.b_ij, p.r_2, &p.shortlimit2, &p.longlimit2,
                                          &p.lowerPotential, &p.upperPotential, &tmpDouble, 0, 0);
                if (p.lowerPotential == 0 && p.upperPotential == 0) continue;
            }
            if (p.b_ij < 0 && calc_disc(&p) > 0) {
                p.s_time = -1; p.d_ij2 = p.shortlimit2; ev.potential = p.lowerPotential;
                ev.sub = (p.coreorShell == 0) ? SDMD_SUB_CORECOLLI : SDMD_SUB_TBDCAPTR;
            } else {
                p.s_time = 1; p.d_ij2 = p.longlimit2; ev.potential = p.upperPotential;
                if (v->method_ding) {
                    if (cm & SDMD_CONSTRAINT_CONNECT) ev.sub = (p.coreorShell == 1) ? SDMD_SUB_WELLBOUNC : SDMD_SUB_TBDESCAP;
                    else { if (p.coreorShell == 1) continue; ev.sub = SDMD_SUB_TBDESCAP; }
                }
            }
            ev.time = calc_time(&p);
            if (ev.time < 0) c->st.error_prints++; /* reference prints an error and carries on, :168-173 */
            job_assign(c, i, j, &ev, SDMD_EV_COLI);
        }
    }
}

/* BondTime, TimePrediction.c:186-295 */
static void bond_time(oracle_ctx *c, int i) {
    struct sdmd_host_view *v = c->v;
    struct evt ev; struct par p;
    init_event(&ev);
    for (int k = v->bond_off[i]; k < v->bond_off[i + 1]; k++) {
        int j = v->bond_j[k];
        if (j == 0) break;
        if (check_dup(c, i, j, SDMD_EV_BOND)) continue;
        p.longlimit2 = v->bond_dmax2[k];
        p.shortlimit2 = v->bond_dmin2[k];
        calc_params(c, i, j, &p);
        if (p.r_2 - p.longlimit2 > ZERO) {           /* :240-253 */
            c->st.error_prints++;
            if (p.b_ij >= 0) {
                ev.sub = SDMD_SUB_WELLBOUNC;
                ev.time = 1e-16 * i * j;
                job_assign(c, i, j, &ev, SDMD_EV_BOND);
                continue;
            }
        } else if (p.shortlimit2 - p.r_2 > ZERO) {   /* :255-269 */
            c->st.error_prints++;
            if (p.b_ij < 0) {
                ev.sub = SDMD_SUB_CORECOLLI;
                ev.time = 1e-16 * i * j;
                job_assign(c, i, j, &ev, SDMD_EV_BOND);
                continue;
            }
        }
        if (p.b_ij < 0 && calc_disc(&p) > 0) { p.s_time = -1; p.d_ij2 = p.shortlimit2; ev.sub = SDMD_SUB_CORECOLLI; }
        else { p.s_time = 1; p.d_ij2 = p.longlimit2; ev.sub = SDMD_SUB_WELLBOUNC; }
        ev.time = calc_time(&p);
        if (ev.time < 0) { c->st.error_prints++; ev.time = INF; }
        job_assign(c, i, j, &ev, SDMD_EV_BOND);
    }
}

/* CheckPairMatch, TimePrediction.c:1119-1138 */
static int check_pair_match(oracle_ctx *c, int i, int j) {
    const struct sdmd_host_view *v = c->v;
    int cls = hb_model(v, i, j);
    if (cls == -2) { c->st.error = SDMD_ERR_TABLE; return -1; }
    if (cls < 0) return -1;
    int ti = v->type[i], tj = v->type[j];
#define HAS_STEP(a, b) (v->hb_index[(cls * NT + (a)) * NT + (b)] >= 0 && v->tables[v->hb_index[(cls * NT + (a)) * NT + (b)]].nstep > 0)
    if (!HAS_STEP(ti, tj)) return -1;
    int ni = type_change(v->type[v->hb_neighbor[i]], 0);
    int nj = type_change(v->type[v->hb_neighbor[j]], 0);
    if (!HAS_STEP(ti, nj) || !HAS_STEP(tj, ni)) return -1;
#undef HAS_STEP
    return 0;
}

/* HBTime, TimePrediction.c:298-483 */
static void hb_time(oracle_ctx *c, int i) {
    struct sdmd_host_view *v = c->v;
    struct evt ev; struct par p; double tmpDouble;
    init_event(&ev);
    if (v->hb_itype[i] == SDMD_HB_SKIP) { v->hb_itype[i] = SDMD_HB_NOEVENT; return; }
    v->hb_itype[i] = SDMD_HB_NOEVENT;
    if (v->hb_bond[i] > 0) {
        int j = v->hb_bond[i];
        if (check_dup(c, i, j, SDMD_EV_HB)) return;
        calc_params(c, i, j, &p);
        p.coreorShell = find_pair(c, i, j, FP_HB, p.b_ij, p.r_2, &p.shortlimit2, &p.longlimit2, &p.lowerPotential,
                                  &p.upperPotential, &tmpDouble, 0, 0);
        if (p.b_ij < 0 && calc_disc(&p) > 0) {
            p.s_time = -1; p.d_ij2 = p.shortlimit2; ev.potential = p.lowerPotential;
            ev.sub = (p.coreorShell == 0) ? SDMD_SUB_CORECOLLI : SDMD_SUB_TBDCAPTR;
        } else {
            p.s_time = 1; p.d_ij2 = p.longlimit2; ev.potential = p.upperPotential;
            if (v->method_ding) ev.sub = (p.coreorShell == 1) ? SDMD_SUB_TBDHBBRIK : SDMD_SUB_TBDESCAP;
        }
        ev.time = calc_time(&p);
        if (ev.time < 0) { c->st.error_prints++; ev.time = INF; }
        job_assign(c, i, j, &ev, SDMD_EV_HB);
    } else if ((v->hb_role[i] == 'A' || v->hb_role[i] == 'D' || v->hb_role[i] == 'S') && v->hb_bond[i] == 0) {
        double cut2 = v->cutoffr * SDMD_EXTEND_RATIO;
        cut2 *= cut2;
        for (int k = 0; k < 27; k++) {
            int cell = neighbor_cell(c, i, k);
            for (int j = c->celllist[c->n + cell]; j; j = c->celllist[j]) {
                if (i == j) continue;
                int ad = (v->hb_role[i] == 'A' && v->hb_role[j] == 'D') || (v->hb_role[i] == 'D' && v->hb_role[j] == 'A');
                int sd = (v->hb_role[i] == 'S' && v->type[j] == 2) || (v->hb_role[j] == 'S' && v->type[i] == 2);
                int ss = (v->hb_role[i] == 'S' && v->hb_role[j] == 'S');
                if (!(ad || sd || ss)) continue;
                if (v->hb_bond[j] != 0) continue;
                int d = v->aa_num[i] - v->aa_num[j];
                if (d < 0) d = -d;
                if (!((v->protein_num[i] == v->protein_num[j] && d >= 4) || v->protein_num[i] != v->protein_num[j])) continue;
                if (check_dup(c, i, j, SDMD_EV_HB)) continue;
                calc_params(c, i, j, &p);
                if (!(p.r_2 <= cut2)) continue;
                if (check_pair_match(c, i, j) < 0) continue;
                p.coreorShell = find_pair(c, i, j, FP_HB, p.b_ij, p.r_2, &p.shortlimit2, &p.longlimit2, &p.lowerPotential,
                                          &p.upperPotential, &tmpDouble, 0, 0);
                if (p.b_ij < 0 && p.coreorShell == 1 && calc_disc(&p) > 0) {
                    p.s_time = -1; p.d_ij2 = p.shortlimit2; ev.potential = p.lowerPotential; ev.sub = SDMD_SUB_TBDHBFORM;
                } else continue;
                ev.time = calc_time(&p);
                if (ev.time < 0) { c->st.error_prints++; ev.time = INF; }
                job_assign(c, i, j, &ev, SDMD_EV_HB);
            }
        }
    }
}

/* HBNeighborTime, TimePrediction.c:486-579 */
static void hbn_time(oracle_ctx *c, int i) {
    struct sdmd_host_view *v = c->v;
    struct evt ev; struct par p; double tmpDouble;
    int flag = 0;
    init_event(&ev);
    for (int n = 1; n <= v->hbn_status[i]; n++) {
        int j = v->hbn_partner[5 * i + n];
        if (flag == 0) v->hbn_partner_num[i] = n;
        if (check_dup(c, i, j, SDMD_EV_HBNE)) { v->hbn_partner_num[i] = n; flag = 1; continue; }
        calc_params(c, i, j, &p);
        int host, nbr, third;
        int r = v->hb_role[i];
        if ((r == 'A' || r == 'D' || r == 'S') && (v->hb_bond[i] && v->hb_neighbor[v->hb_bond[i]] == j)) {
            host = i; nbr = j; third = v->hb_bond[i];
        } else {
            host = j; nbr = i; third = v->hb_bond[j];
        }
        p.coreorShell = find_pair(c, nbr, host, FP_NEIGHBOR, p.b_ij, p.r_2, &p.shortlimit2, &p.longlimit2, &p.lowerPotential,
                                  &p.upperPotential, &tmpDouble, third, 0);
        if (p.b_ij < 0 && calc_disc(&p) > 0) {
            p.s_time = -1; p.d_ij2 = p.shortlimit2; ev.potential = p.lowerPotential;
            ev.sub = (p.coreorShell == 0) ? SDMD_SUB_CORECOLLI : SDMD_SUB_TBDCAPTR;
        } else {
            p.s_time = 1; p.d_ij2 = p.longlimit2; ev.potential = p.upperPotential;
            if (v->method_ding) ev.sub = (p.coreorShell == 1) ? SDMD_SUB_WELLBOUNC : SDMD_SUB_TBDESCAP;
        }
        ev.time = calc_time(&p);
        if (ev.time < 0) { c->st.error_prints++; ev.time = INF; }
        if (job_assign(c, i, j, &ev, SDMD_EV_HBNE)) { v->hbn_partner_num[i] = n; flag = 1; }
    }
}

/* ThermostatTime, TimePrediction.c:582-595 (Maxwell_Boltzmann_Distribution(0,1) == the raw gaussian) */
static void thermostat_time(oracle_ctx *c, int i) {
    if (!c->v->thermostat_on) return;
    struct evt ev;
    init_event(&ev);
    double g = next_gauss(c);
    ev.time = c->v->thermoF * ((g < 0) ? (-g) : (g));
    job_assign(c, i, 0, &ev, SDMD_EV_THER);
}

/* PBCandCrossCellTime, TimePrediction.c:598-632 */
static void pcc_time(oracle_ctx *c, int i) {
    const struct sdmd_host_view *v = c->v;
    struct evt ev;
    init_event(&ev);
    double pos[3] = {v->x[i], v->y[i], v->z[i]}, spd[3] = {v->vx[i], v->vy[i], v->vz[i]};
    double lo[3], hi[3];
    for (int k = 0; k < 3; k++) {
        lo[k] = v->cell_idx[3 * i + k] * c->cellsize[k];
        hi[k] = (v->cell_idx[3 * i + k] + 1) * c->cellsize[k];
    }
    for (int k = 0; k < 3; k++) {
        double t;
        if (spd[k] > 0) {
            t = (hi[k] - pos[k] + ZERO) / spd[k];
            if (ev.time > t && t > 0) ev.time = t;
        } else if (spd[k] < 0) {
            t = (lo[k] - pos[k] - ZERO) / spd[k];
            if (ev.time > t && t > 0) ev.time = t;
        }
    }
    job_assign(c, i, 0, &ev, SDMD_EV_PCC);
}

/* WallTime, TimePrediction.c:635-737, smooth wall, wallDyn "no".  The wall is a pseudo-atom at rest sitting on the
 * mid-plane / axis / centre; `distance2` is the squared distance from the wall SURFACE, (box_y/2 - r)^2 written the
 * reference's way; the quadratic itself is solved in the distance r from the centre. */
static void wall_time(oracle_ctx *c, int i) {
    const struct sdmd_host_view *v = c->v;
    struct evt ev;
    struct par p;
    init_event(&ev);
    double pi[3] = {v->x[i], v->y[i], v->z[i]}, si[3] = {v->vx[i], v->vy[i], v->vz[i]};
    if (v->wall_kind == SDMD_WALL_PARALLEL) { pi[0] = pi[2] = 0; si[0] = si[2] = 0; }   /* :654-656 */
    else if (v->wall_kind == SDMD_WALL_CYLINDER) { pi[0] = 0; si[0] = 0; }              /* :657-660 */
    for (int k = 0; k < 3; k++) { p.r_ij[k] = pi[k] - v->wall_pos[k]; p.v_ij[k] = si[k] - 0.0; }
    p.r_2 = p.r_ij[0] * p.r_ij[0] + p.r_ij[1] * p.r_ij[1] + p.r_ij[2] * p.r_ij[2];
    p.b_ij = p.r_ij[0] * p.v_ij[0] + p.r_ij[1] * p.v_ij[1] + p.r_ij[2] * p.v_ij[2];
    p.v_2 = p.v_ij[0] * p.v_ij[0] + p.v_ij[1] * p.v_ij[1] + p.v_ij[2] * p.v_ij[2];
    double by, accum;
    for (;;) {                                                                           /* ReCal, :641 */
        by = v->box[1];
        double distance2 = by * by * 0.25 + p.r_2 - by * sqrt(p.r_2);                  /* :667 */
        p.coreorShell = find_pair(c, i, -1, FP_COLLISION, -1 * p.b_ij, distance2, &p.shortlimit2, &p.longlimit2,
                                  &p.lowerPotential, &p.upperPotential, &accum, 0, 0);
        if (!(p.lowerPotential == 0 && p.upperPotential == 0)) break;
        if (!c->wd.mode) { c->st.error = SDMD_ERR_STATE; return; }                       /* :699-701 exit() */
        /* moving wall: the atom is outside -> the box is pushed back out by one step and the prediction repeated, :678-697 */
        double *box = c->v->box;
        c->wd.touch++;
        if (v->wall_kind == SDMD_WALL_PARALLEL) box[1] += c->wd.step;
        else if (v->wall_kind == SDMD_WALL_CYLINDER) { box[1] += c->wd.step; box[2] += c->wd.step; }
        else { box[0] += c->wd.step; box[1] += c->wd.step; box[2] += c->wd.step; }
        if (box[1] > c->wd.orig_box[1]) { c->st.error = SDMD_ERR_STATE; return; }        /* :691-694 exit() */
    }
    if (p.b_ij < 0 && calc_disc(&p) > 0) {                                               /* towards the centre :704-714 */
        if (p.coreorShell == 1) return;
        ev.sub = SDMD_SUB_TBDCAPTR;
        p.s_time = -1;
        p.d_ij2 = by * by * 0.25 + p.longlimit2 - by * sqrt(p.longlimit2);
        ev.potential = p.upperPotential;
    } else {                                                                             /* towards the wall :716-726 */
        ev.sub = (p.coreorShell == 0) ? SDMD_SUB_WELLBOUNC : SDMD_SUB_TBDESCAP;
        p.s_time = 1;
        p.d_ij2 = by * by * 0.25 + p.shortlimit2 - by * sqrt(p.shortlimit2);
        ev.potential = p.lowerPotential;
    }
    ev.time = calc_time(&p);
    if (ev.time < 0) ev.time = INF;                                                      /* :729-732 */
    job_assign(c, i, 0, &ev, SDMD_EV_WALL);
}

/* FindPotWellWidth, ToolFunctions.c:954-970: squared range of the pair potential of (type_i, type_j) */
static double pot_well_width(const struct sdmd_host_view *v, int type_i, int type_j) {
    int idx = v->coll_index[type_i * NT + type_j];
    if (idx < 0) return 0;
    const struct sdmd_step_table *t = &v->tables[idx];
    if (t->dmax2 > 0) return t->dmax2;
    return t->nstep > 0 ? t->d2[t->nstep - 1] : 0;
}

/* EnterTunnel, ToolFunctions.c:921-952: will the atom be inside a tunnel mouth's cross-section after `time`? */
static int enter_tunnel(oracle_ctx *c, double time, int i) {
    const struct sdmd_host_view *v = c->v;
    for (int n = 0; n < v->n_tunnel; n++) {
        double radius2 = v->tunnel_diameter[n] * 0.5;
        radius2 *= radius2;
        double shift2 = pot_well_width(v, v->tunnel_type, v->type[i]);
        double py = v->y[i], pz = v->z[i];
        py += v->vy[i] * time;
        pz += v->vz[i] * time;
        double r1 = v->tunnel_pos[n][1] - v->tunnel_pos[n][1], r2 = py - v->tunnel_pos[n][2], r3 = pz - v->tunnel_pos[n][3];
        double r_2 = r1 * r1 + r2 * r2 + r3 * r3;
        double distance2 = radius2 + shift2 - 2 * sqrt(radius2 * shift2);
        if (r_2 <= distance2) return 1;
    }
    return 0;
}

/* TunnelTime, TimePrediction.c:832-930: inside a tunnel the wall distance is taken from the tunnel's axis (y, z) */
static void tunnel_time(oracle_ctx *c, int i) {
    struct sdmd_host_view *v = c->v;
    int idx = v->coll_index[v->type[i] * NT + v->tunnel_type];
    if (idx < 0) { c->st.error = SDMD_ERR_TABLE; return; }
    double distanceShift2 = v->tables[idx].dmin2;
    double extraRange = sqrt(pot_well_width(v, v->tunnel_type, v->type[i]));
    double pos0 = v->x[i], spd0 = v->vx[i];
    for (int n = 0; n < v->n_tunnel; n++) {
        struct evt ev;
        struct par p;
        double radius2 = v->tunnel_diameter[n] / 2;
        radius2 *= radius2;
        init_event(&ev);
        p.r_ij[0] = v->tunnel_pos[n][1] - v->tunnel_pos[n][1]; p.r_ij[1] = v->y[i] - v->tunnel_pos[n][2]; p.r_ij[2] = v->z[i] - v->tunnel_pos[n][3];
        p.v_ij[0] = 0 - 0.0; p.v_ij[1] = v->vy[i] - 0.0; p.v_ij[2] = v->vz[i] - 0.0;
        p.r_2 = p.r_ij[0] * p.r_ij[0] + p.r_ij[1] * p.r_ij[1] + p.r_ij[2] * p.r_ij[2];
        p.b_ij = p.r_ij[0] * p.v_ij[0] + p.r_ij[1] * p.v_ij[1] + p.r_ij[2] * p.v_ij[2];
        p.v_2 = p.v_ij[0] * p.v_ij[0] + p.v_ij[1] * p.v_ij[1] + p.v_ij[2] * p.v_ij[2];
        if (p.r_2 > radius2 + distanceShift2 - 2 * sqrt(radius2 * distanceShift2) + ZERO) continue;   /* outside the tunnel */
        double distance2 = radius2 + p.r_2 - 2 * sqrt(p.r_2 * radius2);
        double accum;
        p.coreorShell = find_pair(c, i, -100, FP_COLLISION, -1 * p.b_ij, distance2, &p.shortlimit2, &p.longlimit2,
                                  &p.lowerPotential, &p.upperPotential, &accum, 0, 0);
        if (p.lowerPotential == 0 && p.upperPotential == 0) { c->st.error = SDMD_ERR_STATE; return; }   /* exit(), :880-884 */
        if (p.b_ij < 0 && calc_disc(&p) > 0) {
            if (p.coreorShell == 1) continue;
            ev.sub = SDMD_SUB_TBDCAPTR;
            p.s_time = -1;
            p.d_ij2 = radius2 + p.longlimit2 - 2 * sqrt(p.longlimit2 * radius2);
            ev.potential = p.upperPotential;
        } else {
            ev.sub = (p.coreorShell == 0) ? SDMD_SUB_WELLBOUNC : SDMD_SUB_TBDESCAP;
            p.s_time = 1;
            p.d_ij2 = radius2 + p.shortlimit2 - 2 * sqrt(p.shortlimit2 * radius2);
            ev.potential = p.lowerPotential;
        }
        ev.time = calc_time(&p);
        double futurePos = pos0 + spd0 * ev.time;                                          /* :912-917 */
        if (futurePos < v->tunnel_pos[n][0] - extraRange || futurePos > v->tunnel_pos[n][1] + extraRange) continue;
        if (ev.time < 0) ev.time = INF;
        if (job_assign(c, i, 0, &ev, SDMD_EV_TUNL)) v->ev_partner[i] = -10000 * n;
    }
}

/* ObstTime, TimePrediction.c:740-829, obstruction walls: one-dimensional distance from the plane */
static void obst_time(oracle_ctx *c, int i) {
    struct sdmd_host_view *v = c->v;
    const double pos[3] = {v->x[i], v->y[i], v->z[i]}, spd[3] = {v->vx[i], v->vy[i], v->vz[i]};
    for (int n = 0; n < v->n_obst; n++) {
        struct evt ev;
        struct par p;
        init_event(&ev);
        for (int dim = 0; dim < 3; dim++) {
            if (v->obst_pos[n][dim] < 0) continue;
            for (int k = 0; k < 3; k++) {
                double pi = (k == dim) ? pos[k] : 0, pj = (k == dim) ? v->obst_pos[n][k] : 0, si = (k == dim) ? spd[k] : 0;
                p.r_ij[k] = pi - pj; p.v_ij[k] = si - 0.0;
            }
            p.r_2 = p.r_ij[0] * p.r_ij[0] + p.r_ij[1] * p.r_ij[1] + p.r_ij[2] * p.r_ij[2];
            p.b_ij = p.r_ij[0] * p.v_ij[0] + p.r_ij[1] * p.v_ij[1] + p.r_ij[2] * p.v_ij[2];
            p.v_2 = p.v_ij[0] * p.v_ij[0] + p.v_ij[1] * p.v_ij[1] + p.v_ij[2] * p.v_ij[2];
            double accum;
            p.coreorShell = find_pair(c, i, -2 - n, FP_COLLISION, p.b_ij, p.r_2, &p.shortlimit2, &p.longlimit2,
                                      &p.lowerPotential, &p.upperPotential, &accum, 0, 0);
            if (p.lowerPotential == 0 && p.upperPotential == 0) continue;           /* :783-787 */
            if (p.b_ij < 0) {                                                         /* approaching :789-798 */
                ev.sub = (p.coreorShell == 0) ? SDMD_SUB_CORECOLLI : SDMD_SUB_TBDCAPTR;
                p.s_time = -1; p.d_ij2 = p.shortlimit2; ev.potential = p.lowerPotential;
            } else {                                                                  /* leaving :800-810 */
                if (p.coreorShell == 1) continue;
                ev.sub = SDMD_SUB_TBDESCAP;
                p.s_time = 1; p.d_ij2 = p.longlimit2; ev.potential = p.upperPotential;
            }
            ev.time = calc_time(&p);
            if (ev.time < 0) { c->st.error = SDMD_ERR_STATE; return; }                /* exit(), :813-816 */
            /* at the next collision the atom would already be in the hole: the prediction is void, :818-824 */
            if (!(v->obst_hole[n] && enter_tunnel(c, ev.time, i)) && job_assign(c, i, 0, &ev, SDMD_EV_OBST))
                v->ev_partner[i] = -1 * (dim + 1) - 10 * n;
        }
    }
}

/* ChargeTime, TimePrediction.c:933-1017: concentric potential steps of width `gap` around each charge source */
static void charge_time(oracle_ctx *c, int i) {
    struct sdmd_host_view *v = c->v;
    if (!v->charge || !v->charge[i]) return;
    double charge = v->charge[i];
    int flag = 0;
    for (int cn = 0; cn < v->n_charge; cn++) {
        if (v->x[i] > v->charge_pos[cn][0] - ZERO) continue;
        if (v->charge_pos[cn][0] < v->box[0]) flag = 1;                               /* :948-956 */
        else if (!(flag == 0 && v->charge_pos[cn][0] > v->box[0])) break;
        struct evt ev;
        struct par p;
        init_event(&ev);
        p.r_ij[0] = v->x[i] - v->charge_pos[cn][0]; p.r_ij[1] = v->y[i] - v->charge_pos[cn][1]; p.r_ij[2] = v->z[i] - v->charge_pos[cn][2];
        p.v_ij[0] = v->vx[i] - v->charge_vel[cn][0]; p.v_ij[1] = v->vy[i] - v->charge_vel[cn][1]; p.v_ij[2] = v->vz[i] - v->charge_vel[cn][2];
        p.r_2 = p.r_ij[0] * p.r_ij[0] + p.r_ij[1] * p.r_ij[1] + p.r_ij[2] * p.r_ij[2];
        p.b_ij = p.r_ij[0] * p.v_ij[0] + p.r_ij[1] * p.v_ij[1] + p.r_ij[2] * p.v_ij[2];
        p.v_2 = p.v_ij[0] * p.v_ij[0] + p.v_ij[1] * p.v_ij[1] + p.v_ij[2] * p.v_ij[2];
        int gapNum = sqrt(p.r_2) / v->charge_gap[cn];
        double gap2 = v->charge_gap[cn] * v->charge_gap[cn];
        double distance2 = gapNum * gapNum * gap2;
        if (p.b_ij < 0 && p.r_2 < distance2 + ZERO) gapNum--;
        else if (p.b_ij > 0 && p.r_2 > (gapNum + 1) * (gapNum + 1) * gap2 - ZERO) gapNum++;
        p.upperPotential = v->charge_potgrad[cn] * charge;
        p.longlimit2 = (gapNum + 1) * (gapNum + 1) * gap2;
        p.lowerPotential = -1 * v->charge_potgrad[cn] * charge;
        p.shortlimit2 = gapNum * gapNum * gap2;
        if (p.b_ij < 0 && calc_disc(&p) > 0) {
            p.s_time = -1; p.d_ij2 = p.shortlimit2; ev.potential = p.lowerPotential; ev.sub = SDMD_SUB_TBDCAPTR;
        } else {
            p.s_time = 1; p.d_ij2 = p.longlimit2; ev.potential = p.upperPotential; ev.sub = SDMD_SUB_TBDESCAP;
        }
        ev.time = calc_time(&p);
        if (ev.time < 0) ev.time = INF;
        if (job_assign(c, i, 0, &ev, SDMD_EV_CHRG)) v->ev_partner[i] = -100 * cn;
    }
}

/* SphObstTime, TimePrediction.c:1020-1071: hard sphere at rest, core collision only */
static void sphobst_time(oracle_ctx *c, int i) {
    struct sdmd_host_view *v = c->v;
    for (int n = 0; n < v->n_sphobst; n++) {
        struct evt ev;
        struct par p;
        init_event(&ev);
        p.r_ij[0] = v->x[i] - v->sphobst_pos[n][0]; p.r_ij[1] = v->y[i] - v->sphobst_pos[n][1]; p.r_ij[2] = v->z[i] - v->sphobst_pos[n][2];
        p.v_ij[0] = v->vx[i] - 0.0; p.v_ij[1] = v->vy[i] - 0.0; p.v_ij[2] = v->vz[i] - 0.0;
        p.r_2 = p.r_ij[0] * p.r_ij[0] + p.r_ij[1] * p.r_ij[1] + p.r_ij[2] * p.r_ij[2];
        p.b_ij = p.r_ij[0] * p.v_ij[0] + p.r_ij[1] * p.v_ij[1] + p.r_ij[2] * p.v_ij[2];
        p.v_2 = p.v_ij[0] * p.v_ij[0] + p.v_ij[1] * p.v_ij[1] + p.v_ij[2] * p.v_ij[2];
        p.shortlimit2 = v->sphobst_radius[n] * v->sphobst_radius[n];
        if (!(p.b_ij < 0 && calc_disc(&p) > 0)) continue;
        ev.sub = SDMD_SUB_CORECOLLI;
        p.s_time = -1; p.d_ij2 = p.shortlimit2; ev.potential = 0;
        ev.time = calc_time(&p);
        if (ev.time < 0) { c->st.error = SDMD_ERR_STATE; return; }
        if (job_assign(c, i, 0, &ev, SDMD_EV_SPHO)) v->ev_partner[i] = -1000 * n;
    }
}

/* Predict, SGThread.c:158-211 (+ ResetTarget, ToolFunctions.c:504-514) */
void oracle_predict(oracle_ctx *c, const int32_t *renew) {
    struct sdmd_host_view *v = c->v;
    for (int n = 1; renew[n]; n++) {
        int i = renew[n];
        v->ev_type[i] = SDMD_EV_INVD; v->ev_time[i] = INF; v->ev_partner[i] = SDMD_INVALID;
    }
    for (int n = 1; renew[n]; n++) {
        int i = renew[n];
        if (v->is_cons[i]) { v->ev_time[i] = INF; continue; }
        if (!v->gel_mode) bond_time(c, i);
        collision_time(c, i);
        pcc_time(c, i);
        thermostat_time(c, i);
        if (v->hydrogen_bond) { hb_time(c, i); hbn_time(c, i); }
        if (v->wall_kind) wall_time(c, i);                       /* SGThread.c:183-185 */
        if (v->n_obst) obst_time(c, i);                          /* :187-189 */
        if (v->n_tunnel) tunnel_time(c, i);                      /* :197-199 */
        if (v->flow_mark == 3) charge_time(c, i);                /* :201-203 */
        if (v->n_sphobst) sphobst_time(c, i);                    /* :205-207 */
    }
}

void oracle_predict_all(oracle_ctx *c) {
    int32_t *renew = calloc((size_t)c->n + 2, sizeof(int32_t));
    for (int i = 1; i <= c->n; i++) renew[i] = i;
    oracle_predict(c, renew);
    free(renew);
}

/* ------------------------------------------------------------------------------------------------ */
/* UpdateCBTRun, CBT.c:56-83 */
static void cbt_update_run(oracle_ctx *c, int a) {
    const double *t = c->v->ev_time;
    int father, left, right, old;
    for (father = c->leaf[a] / 2; father > 0; father /= 2) {
        if (c->node[father] != a) break;
        left = c->node[father * 2]; right = c->node[father * 2 + 1];
        c->node[father] = (t[left] < t[right]) ? left : right;
    }
    for (; father > 0; father /= 2) {
        old = c->node[father];
        left = c->node[father * 2]; right = c->node[father * 2 + 1];
        c->node[father] = (t[left] < t[right]) ? left : right;
        if (c->node[father] == old) return;
    }
}
/* CreateCBT + InsertCBTRun, CBT.c:16-30,86-105 */
void oracle_cbt_build(oracle_ctx *c) {
    free(c->node); free(c->leaf);
    c->node = calloc((size_t)c->n * 2 + 2, sizeof(int32_t));
    c->leaf = calloc((size_t)c->n + 1, sizeof(int32_t));
    c->cbt_count = 0;
    for (int a = 1; a <= c->n; a++) {
        if (c->cbt_count == 0) { c->node[1] = a; c->cbt_count++; continue; }
        int i = c->node[c->cbt_count];
        c->node[c->cbt_count * 2] = i;
        c->node[c->cbt_count * 2 + 1] = a;
        c->leaf[i] = c->cbt_count * 2;
        c->leaf[a] = c->cbt_count * 2 + 1;
        c->cbt_count++;
        cbt_update_run(c, i);
    }
}
void oracle_cbt_update(oracle_ctx *c, const int32_t *renew) {
    for (int n = 1; renew[n]; n++) cbt_update_run(c, renew[n]);
}

/* UpdateData + TimeForward, ToolFunctions.c:23-36 */
void oracle_advance(oracle_ctx *c, double dt) {
    struct sdmd_host_view *v = c->v;
    for (int i = 1; i <= c->n; i++) {
        v->x[i] += v->vx[i] * dt;
        v->y[i] += v->vy[i] * dt;
        v->z[i] += v->vz[i] * dt;
    }
    for (int i = 1; i <= c->n; i++) v->ev_time[i] -= dt;
}

/* ------------------------------------------------------------------------------------------------ */
/* HBBarrier, Events.c:1097-1119 */
static double hb_barrier(const struct sdmd_host_view *v, int a1, int a2) {
    if (v->hb_role[a1] == 'A' || v->hb_role[a1] == 'S') { int s = a1; a1 = a2; a2 = s; }
    int o1 = type_change(v->type[a1], 0), o2 = type_change(v->type[a2], 0);
    if (o1 == 2 && o2 == 26) return v->hb_barrier[0];
    if (o1 != 2 && o2 != 26) return v->hb_barrier[1];
    return v->hb_barrier[2];
}

/* HBEnergyChange, Events.c:1027-1095; hbType: 0 "_nonHB", 1 "_HB" */
static double hb_energy_change(oracle_ctx *c, int i, int skipAtom, int hbType) {
    const struct sdmd_host_view *v = c->v;
    double cutoffRange = v->cutoffr * v->cutoffr;
    double energyChange = 0, tmp = 0, accumPoten = 0;
    for (int k = 0; k < 27; k++) {
        int cell = neighbor_cell(c, i, k);
        for (int j = c->celllist[c->n + cell]; j; j = c->celllist[j]) {
            int cm = cm_flags(v, i, j);
            if (i != j && !(cm & SDMD_BOND_CONNECT) && !(cm & SDMD_CONSTRAINT_CONNECT) && j != skipAtom) {
                double shift[3], rt[3], r[3];
                pbc_shift(c, i, j, shift);
                rt[0] = v->x[j] + shift[0]; rt[1] = v->y[j] + shift[1]; rt[2] = v->z[j] + shift[2];
                r[0] = v->x[i] - rt[0]; r[1] = v->y[i] - rt[1]; r[2] = v->z[i] - rt[2];
                double r2 = r[0] * r[0] + r[1] * r[1] + r[2] * r[2];
                if (r2 <= cutoffRange) {
                    if (hbType == 0) find_pair(c, i, j, FP_COLLISION, tmp, r2, &tmp, &tmp, &tmp, &tmp, &accumPoten, 0, -1);
                    else if (j != v->hb_neighbor[skipAtom]) find_pair(c, i, j, FP_COLLISION, tmp, r2, &tmp, &tmp, &tmp, &tmp, &accumPoten, 0, 1);
                    if (accumPoten == INF) accumPoten = 0;
                    energyChange += accumPoten;
                }
            }
        }
    }
    return energyChange;
}

/* position of b as seen from a under the cell-index minimum image, Events.c:815-823 */
static void image_of(const oracle_ctx *c, int a, int b, double pos[3]) {
    const struct sdmd_host_view *v = c->v;
    pos[0] = v->x[b]; pos[1] = v->y[b]; pos[2] = v->z[b];
    for (int k = 0; k < 3; k++) {
        int ca = v->cell_idx[3 * a + k], cb = v->cell_idx[3 * b + k];
        if (ca != cb) {
            if (ca == 0 && cb == c->cellnum[k] - 1) pos[k] -= v->box[k];
            else if (ca == c->cellnum[k] - 1 && cb == 0) pos[k] += v->box[k];
        }
    }
}

/* InteractionEvent, Events.c:798-1024 */
static void interaction_event(oracle_ctx *c, int i, int j, int hbEvent) {
    struct sdmd_host_view *v = c->v;
    double pi[3] = {v->x[i], v->y[i], v->z[i]}, pj[3];
    image_of(c, i, j, pj);
    double r[3] = {pi[0] - pj[0], pi[1] - pj[1], pi[2] - pj[2]};
    double w[3] = {v->vx[i] - v->vx[j], v->vy[i] - v->vy[j], v->vz[i] - v->vz[j]};
    double d2 = r[0] * r[0] + r[1] * r[1] + r[2] * r[2];
    double b = r[0] * w[0] + r[1] * w[1] + r[2] * w[2];
    double mi = v->mass[i], mj = v->mass[j];
    double reduced = mi * mj / (mi + mj);
    double potential = v->ev_potential[i];
    int s_vel = 0;

    if (hbEvent && (v->ev_sub[i] == SDMD_SUB_TBDHBFORM || v->ev_sub[i] == SDMD_SUB_TBDHBBRIK)) {   /* :832-943 */
        int ni = v->hb_neighbor[i], nj = v->hb_neighbor[j];
        if ((neighbor_connected(v, i, nj) || neighbor_connected(v, j, ni)) && v->ev_sub[i] == SDMD_SUB_TBDHBFORM) {
            v->hb_itype[i] = SDMD_HB_SKIP;
            return;
        }
        double pni[3], pnj[3];
        image_of(c, i, ni, pni);
        image_of(c, i, nj, pnj);
        double ri[3] = {pi[0] - pnj[0], pi[1] - pnj[1], pi[2] - pnj[2]};
        double rj[3] = {pj[0] - pni[0], pj[1] - pni[1], pj[2] - pni[2]};
        double rni2 = ri[0] * ri[0] + ri[1] * ri[1] + ri[2] * ri[2];
        double rnj2 = rj[0] * rj[0] + rj[1] * rj[1] + rj[2] * rj[2];
        double acc[6], tmp = 0, temp, nlo, nup;
        acc[0] = hb_energy_change(c, i, j, 0);
        acc[1] = hb_energy_change(c, j, i, 0);
        acc[2] = hb_energy_change(c, i, j, 1);
        acc[3] = hb_energy_change(c, j, i, 1);
        find_pair(c, nj, i, FP_NEIGHBOR, tmp, rni2, &temp, &temp, &nlo, &nup, &acc[4], j, 0);
        find_pair(c, ni, j, FP_NEIGHBOR, tmp, rnj2, &temp, &temp, &nlo, &nup, &acc[5], i, 0);
        if (v->ev_sub[i] == SDMD_SUB_TBDHBFORM) {
            potential -= hb_barrier(v, i, j);
            potential += (acc[2] - acc[0] + acc[3] - acc[1] + acc[4] + acc[5]);
        } else {
            potential += hb_barrier(v, i, j);
            potential += (acc[0] - acc[2] + acc[1] - acc[3] - acc[4] - acc[5]);
        }
        if (potential < -0.5 * INF && v->ev_sub[i] == SDMD_SUB_TBDHBBRIK) { v->hb_itype[i] = SDMD_HB_SKIP; return; }
        else if (potential > 0.5 * INF && v->ev_sub[i] == SDMD_SUB_TBDHBFORM) { v->hb_itype[i] = SDMD_HB_SKIP; return; }
        if (b * b - 2 * d2 * potential / reduced > 0) {
            if (v->ev_sub[i] == SDMD_SUB_TBDHBFORM) { v->ev_sub[i] = SDMD_SUB_WELLCAPTR; v->hb_itype[i] = SDMD_HB_FORM; }
            else { v->ev_sub[i] = SDMD_SUB_WELLESCAP; v->hb_itype[i] = SDMD_HB_BREAK; }
        } else {
            if (v->ev_sub[i] == SDMD_SUB_TBDHBFORM) { v->hb_itype[i] = SDMD_HB_SKIP; return; }
            potential = 0;
            v->ev_sub[i] = SDMD_SUB_WELLBOUNC;
            v->hb_itype[i] = SDMD_HB_NOEVENT;
        }
    }

    switch (v->ev_sub[i]) {                                                     /* :945-1003 */
        case SDMD_SUB_CORECOLLI: s_vel = 1; potential = 0; break;
        case SDMD_SUB_WELLBOUNC: s_vel = (b < 0) ? 1 : -1; potential = 0; break;
        case SDMD_SUB_WELLCAPTR: s_vel = -1; break;
        case SDMD_SUB_WELLESCAP: s_vel = (b < 0) ? -1 : 1; break;
        case SDMD_SUB_TBDCAPTR:
            if (b * b - 2 * d2 * potential / reduced <= 0) { potential = 0; s_vel = 1; } else s_vel = -1;
            break;
        case SDMD_SUB_TBDESCAP:
            if (b * b - 2 * d2 * potential / reduced <= 0) { potential = 0; s_vel = (b < 0) ? 1 : -1; }
            else s_vel = (b < 0) ? -1 : 1;
            break;
        default: c->st.error = SDMD_ERR_STATE; return;
    }
    double phi = (-1 * b + s_vel * sqrt((b * b - 2 * d2 * potential / reduced))) / ((mi + mj) * d2);    /* :1005-1006 */
    if (isnan(phi)) { c->st.error = SDMD_ERR_STATE; return; }
    if (!v->is_cons[i]) {
        double f = mj * phi;
        v->vx[i] = v->vx[i] + f * r[0]; v->vy[i] = v->vy[i] + f * r[1]; v->vz[i] = v->vz[i] + f * r[2];
    }
    if (!v->is_cons[j]) {
        double f = mi * phi;
        v->vx[j] = v->vx[j] - f * r[0]; v->vy[j] = v->vy[j] - f * r[1]; v->vz[j] = v->vz[j] - f * r[2];
    }
}

/* ListRefresh(target, list, 1, 4), ToolFunctions.c:572-597 */
static void list_refresh(struct sdmd_host_view *v, int a, int target) {
    int32_t *l = &v->hbn_partner[5 * a];
    int i = find_in_hbn(v, a, target);
    if (i) {
        int n = i + 1;
        while (n <= 4 && l[n] != 0) { l[i] = l[n]; i++; n++; }
        l[i] = 0;
    }
}

/* CheckAlphaHB, Events.c:1121-1128 (hb_name: 1 = "H", 2 = "O") */
static int check_alpha_hb(const struct sdmd_host_view *v, int i, int j) {
    if (!v->hb_name || v->protein_num[i] != v->protein_num[j]) return 0;
    int ni = v->hb_name[i], nj = v->hb_name[j];
    return (ni == 1 && nj == 2 && v->aa_num[j] == v->aa_num[i] - 4) || (ni == 2 && nj == 1 && v->aa_num[i] == v->aa_num[j] - 4);
}

/* HBEvent, Events.c:275-393 */
static void hb_event(oracle_ctx *c, int i, int j, int ni, int nj) {
    struct sdmd_host_view *v = c->v;
    if (v->hb_itype[i] == SDMD_HB_FORM) {
        v->hb_bond[i] = j; v->hb_bond[j] = i;
        v->type[i] = type_change(v->type[i], 1);
        v->type[j] = type_change(v->type[j], 1);
        int t;
        t = ++v->hbn_status[i];  v->hbn_partner[5 * i + t] = v->hb_neighbor[j];
        t = ++v->hbn_status[j];  v->hbn_partner[5 * j + t] = v->hb_neighbor[i];
        t = ++v->hbn_status[ni]; v->hbn_partner[5 * ni + t] = j;
        t = ++v->hbn_status[nj]; v->hbn_partner[5 * nj + t] = i;
        if (check_alpha_hb(v, i, j)) c->st.alpha_hb_formed++;    /* Events.c:310-311 */
        c->hbnumformed++;
        c->st.hb_num_formed = c->hbnumformed;
    } else if (v->hb_itype[i] == SDMD_HB_BREAK) {
        v->hb_bond[i] = 0; v->hb_bond[j] = 0;
        v->type[i] = type_change(v->type[i], 0);
        v->type[j] = type_change(v->type[j], 0);
        --v->hbn_status[i];  list_refresh(v, i, v->hb_neighbor[j]);
        --v->hbn_status[j];  list_refresh(v, j, v->hb_neighbor[i]);
        --v->hbn_status[ni]; list_refresh(v, ni, j);
        --v->hbn_status[nj]; list_refresh(v, nj, i);
        if (check_alpha_hb(v, i, j)) c->st.alpha_hb_formed--;    /* Events.c:363-364 */
        c->hbnumformed--;
        c->st.hb_num_formed = c->hbnumformed;
        if (c->st.hb_num_formed < c->st.alpha_hb_formed) c->st.warnings++;   /* Events.c:386-389 */
    }
}

/* ThermostatEvent, Events.c:396-420 (Andersen; RandomVelocity ToolFunctions.c:600-610 with mean 0) */
static void thermostat_event(oracle_ctx *c, int i) {
    struct sdmd_host_view *v = c->v;
    double speed[3];
    double netV[3] = {0, 0, 0};
    if (v->charge_com && v->flow_mark == 3) {       /* CalCOMV of the atom's chain from the raw velocities, Events.c:400-402 */
        int a = i, b = i;
        while (a > 1 && v->protein_num[a - 1] == v->protein_num[i]) a--;
        while (b < c->n && v->protein_num[b + 1] == v->protein_num[i]) b++;
        for (int n = a; n <= b; n++) { netV[0] += v->vx[n] * v->mass[n]; netV[1] += v->vy[n] * v->mass[n]; netV[2] += v->vz[n] * v->mass[n]; }
        for (int k = 0; k < 3; k++) netV[k] /= v->protein_mass[i];
    }
    for (int k = 0; k < 3; k++) {
        do {
            double variance = v->target_temperature / v->mass[i];
            speed[k] = next_gauss(c) * sqrt(variance) + 0;
            /* flow terms, Events.c:408-410 */
            double a = v->flow_a ? v->flow_a[3 * i + k] : 0.0, rec = v->flow_time_rec ? v->flow_time_rec[i] : 0.0;
            speed[k] += v->flow_v[k] + netV[k] + a * (c->st.currenttime - rec);
        } while (speed[k] == 0 && !c->st.error);
    }
    v->vx[i] = speed[0]; v->vy[i] = speed[1]; v->vz[i] = speed[2];
    if (v->flow_time_rec) v->flow_time_rec[i] = c->st.currenttime;                       /* :418 */
}

/* PBCandCrossCellEvent, Events.c:248-272 (+ PBC("part") :230-238) */
static void pcc_event(oracle_ctx *c, int i) {
    struct sdmd_host_view *v = c->v;
    int old[3] = {v->cell_idx[3 * i], v->cell_idx[3 * i + 1], v->cell_idx[3 * i + 2]};
    double *co[3] = {v->x, v->y, v->z};
    int mask = pbc_dim_mask(v);
    for (int k = 0; k < 3; k++) {
        if (old[k] == 0 || old[k] == c->cellnum[k] - 1) {
            for (int d = 0; d < 3; d++) {
                if (!(mask >> d & 1)) continue;
                if (co[d][i] >= v->box[d]) co[d][i] -= v->box[d];
                else if (co[d][i] < 0) co[d][i] += v->box[d];
            }
            break;
        }
    }
    cell_move(c, i, old);
}

/* WallEvent, Events.c:423-499 (smooth wall): radial velocity kick of the atom alone, the wall has no recoil */
static void wall_event(oracle_ctx *c, int i) {
    struct sdmd_host_view *v = c->v;
    double pi[3] = {v->x[i], v->y[i], v->z[i]}, si[3] = {v->vx[i], v->vy[i], v->vz[i]};
    double r[3], vv[3];
    if (v->wall_kind == SDMD_WALL_PARALLEL) { pi[0] = pi[2] = 0; si[0] = si[2] = 0; }
    else if (v->wall_kind == SDMD_WALL_CYLINDER) { pi[0] = 0; si[0] = 0; }
    for (int k = 0; k < 3; k++) { r[k] = pi[k] - v->wall_pos[k]; vv[k] = si[k] - 0.0; }
    double d_ij2 = r[0] * r[0] + r[1] * r[1] + r[2] * r[2];
    double b_ij = r[0] * vv[0] + r[1] * vv[1] + r[2] * vv[2];
    double mass = v->mass[i];
    double potential = v->ev_potential[i];
    int s_vel = 0;
    switch (v->ev_sub[i]) {
        case SDMD_SUB_WELLBOUNC: s_vel = -1; potential = 0; break;
        case SDMD_SUB_TBDCAPTR:
            if (b_ij * b_ij - 2 * d_ij2 * potential / mass <= 0) { potential = 0; s_vel = 1; } else s_vel = -1;
            break;
        case SDMD_SUB_TBDESCAP:
            if (b_ij * b_ij - 2 * d_ij2 * potential / mass <= 0) { potential = 0; s_vel = -1; } else s_vel = 1;
            break;
        default: c->st.error = SDMD_ERR_STATE; return;
    }
    double phi = (-1 * b_ij + s_vel * sqrt(b_ij * b_ij - 2 * d_ij2 * potential / mass)) / d_ij2;
    if (phi != phi) { c->st.error = SDMD_ERR_STATE; return; }
    v->vx[i] = v->vx[i] + phi * r[0]; v->vy[i] = v->vy[i] + phi * r[1]; v->vz[i] = v->vz[i] + phi * r[2];
}

/* ObstEvent, Events.c:543-619, and SphObstEvent, Events.c:752-796: the kick along the one-dimensional / radial
 * separation; the obstacle has no recoil.  sph: spherical obstacle, else obstruction plane (partner code decides which) */
static void obst_event(oracle_ctx *c, int i, int sph) {
    struct sdmd_host_view *v = c->v;
    double r[3], vv[3];
    const double pos[3] = {v->x[i], v->y[i], v->z[i]}, spd[3] = {v->vx[i], v->vy[i], v->vz[i]};
    int partner = v->ev_partner[i];
    if (sph) {
        int n = partner / -1000;
        for (int k = 0; k < 3; k++) { r[k] = pos[k] - v->sphobst_pos[n][k]; vv[k] = spd[k] - 0.0; }
    } else {
        int n = partner / -10, dim = (partner * -1) % 10 - 1;
        for (int k = 0; k < 3; k++) {
            double pi = (k == dim) ? pos[k] : 0, pj = (k == dim) ? v->obst_pos[n][k] : 0, si = (k == dim) ? spd[k] : 0;
            r[k] = pi - pj; vv[k] = si - 0.0;
        }
    }
    double d_ij2 = r[0] * r[0] + r[1] * r[1] + r[2] * r[2];
    double b_ij = r[0] * vv[0] + r[1] * vv[1] + r[2] * vv[2];
    double mass = v->mass[i];
    double potential = v->ev_potential[i];
    int s_vel = 0;
    switch (v->ev_sub[i]) {
        case SDMD_SUB_CORECOLLI: s_vel = 1; potential = 0; break;
        case SDMD_SUB_TBDCAPTR:
            if (sph) { c->st.error = SDMD_ERR_STATE; return; }
            if (b_ij * b_ij - 2 * d_ij2 * potential / mass <= 0) { potential = 0; s_vel = 1; } else s_vel = -1;
            break;
        case SDMD_SUB_TBDESCAP:
            if (sph) { c->st.error = SDMD_ERR_STATE; return; }
            if (b_ij * b_ij - 2 * d_ij2 * potential / mass <= 0) { potential = 0; s_vel = -1; } else s_vel = 1;
            break;
        default: c->st.error = SDMD_ERR_STATE; return;
    }
    double phi = (-1 * b_ij + s_vel * sqrt(b_ij * b_ij - 2 * d_ij2 * potential / mass)) / d_ij2;
    if (phi != phi) { c->st.error = SDMD_ERR_STATE; return; }
    v->vx[i] = v->vx[i] + phi * r[0]; v->vy[i] = v->vy[i] + phi * r[1]; v->vz[i] = v->vz[i] + phi * r[2];
}

/* ChargeEvent, Events.c:685-749 */
static void charge_event(oracle_ctx *c, int i) {
    struct sdmd_host_view *v = c->v;
    int cn = v->ev_partner[i] / -100;
    double r[3] = {v->x[i] - v->charge_pos[cn][0], v->y[i] - v->charge_pos[cn][1], v->z[i] - v->charge_pos[cn][2]};
    double vv[3] = {v->vx[i] - v->charge_vel[cn][0], v->vy[i] - v->charge_vel[cn][1], v->vz[i] - v->charge_vel[cn][2]};
    double d_ij2 = r[0] * r[0] + r[1] * r[1] + r[2] * r[2];
    double b_ij = r[0] * vv[0] + r[1] * vv[1] + r[2] * vv[2];
    double mass = v->mass[i];
    double potential = v->ev_potential[i];
    int s_vel = 0;
    switch (v->ev_sub[i]) {
        case SDMD_SUB_TBDCAPTR:
            if (b_ij * b_ij - 2 * d_ij2 * potential / mass <= 0) { potential = 0; s_vel = 1; } else s_vel = -1;
            break;
        case SDMD_SUB_TBDESCAP:
            if (b_ij * b_ij - 2 * d_ij2 * potential / mass <= 0) { potential = 0; s_vel = -1; } else s_vel = 1;
            break;
        default: c->st.error = SDMD_ERR_STATE; return;
    }
    double phi = (-1 * b_ij + s_vel * sqrt(b_ij * b_ij - 2 * d_ij2 * potential / mass)) / d_ij2;
    if (phi != phi) { c->st.error = SDMD_ERR_STATE; return; }
    v->vx[i] = v->vx[i] + phi * r[0]; v->vy[i] = v->vy[i] + phi * r[1]; v->vz[i] = v->vz[i] + phi * r[2];
}

/* TunnelEvent, Events.c:621-682: WallEvent's kick with the separation taken from the tunnel's axis */
static void tunnel_event(oracle_ctx *c, int i) {
    struct sdmd_host_view *v = c->v;
    int n = v->ev_partner[i] / -10000;
    double r[3] = {v->tunnel_pos[n][1] - v->tunnel_pos[n][1], v->y[i] - v->tunnel_pos[n][2], v->z[i] - v->tunnel_pos[n][3]};
    double vv[3] = {0 - 0.0, v->vy[i] - 0.0, v->vz[i] - 0.0};
    double d_ij2 = r[0] * r[0] + r[1] * r[1] + r[2] * r[2];
    double b_ij = r[0] * vv[0] + r[1] * vv[1] + r[2] * vv[2];
    double mass = v->mass[i];
    double potential = v->ev_potential[i];
    int s_vel = 0;
    switch (v->ev_sub[i]) {
        case SDMD_SUB_WELLBOUNC: s_vel = -1; potential = 0; break;
        case SDMD_SUB_TBDCAPTR:
            if (b_ij * b_ij - 2 * d_ij2 * potential / mass <= 0) { potential = 0; s_vel = 1; } else s_vel = -1;
            break;
        case SDMD_SUB_TBDESCAP:
            if (b_ij * b_ij - 2 * d_ij2 * potential / mass <= 0) { potential = 0; s_vel = -1; } else s_vel = 1;
            break;
        default: c->st.error = SDMD_ERR_STATE; return;
    }
    double phi = (-1 * b_ij + s_vel * sqrt(b_ij * b_ij - 2 * d_ij2 * potential / mass)) / d_ij2;
    if (phi != phi) { c->st.error = SDMD_ERR_STATE; return; }
    v->vx[i] = v->vx[i] + phi * r[0]; v->vy[i] = v->vy[i] + phi * r[1]; v->vz[i] = v->vz[i] + phi * r[2];
}

/* DoEvent, Events.c:25-114 */
static void do_event(oracle_ctx *c, int i, int j) {
    struct sdmd_host_view *v = c->v;
    switch (v->ev_type[i]) {
        case SDMD_EV_COLI: interaction_event(c, i, j, 0); c->st.collision++; break;
        case SDMD_EV_BOND: interaction_event(c, i, j, 0); c->st.bond++; break;
        case SDMD_EV_HBNE: interaction_event(c, i, j, 0); c->st.hb_neighbor++; break;
        case SDMD_EV_HB: {
            int ni = v->hb_neighbor[i], nj = v->hb_neighbor[j];
            interaction_event(c, i, j, 1);
            hb_event(c, i, j, ni, nj);
            c->st.hb++;
            break;
        }
        case SDMD_EV_THER: thermostat_event(c, i); c->st.thermostat++; break;
        case SDMD_EV_PCC: pcc_event(c, i); c->st.pcc++; break;
        case SDMD_EV_WALL: wall_event(c, i); c->st.wall++; break;
        case SDMD_EV_OBST: obst_event(c, i, 0); c->st.wall++; break;
        case SDMD_EV_TUNL: tunnel_event(c, i); c->st.wall++; break;
        case SDMD_EV_CHRG: charge_event(c, i); break;                       /* no counter, Events.c:93-95 */
        case SDMD_EV_SPHO: obst_event(c, i, 1); c->st.wall++; break;
        default: c->st.error = SDMD_ERR_STATE; break;
    }
}

/* SingleThreadRun, SGThread.c:67-156 */
/* moving wall: wallDyn.* of the reference; box edges are the host view's box[] */
void oracle_set_wall_dyn(oracle_ctx *c, const struct sdmd_wall_dyn *wd) { c->wd = *wd; }
void oracle_get_wall_dyn(const oracle_ctx *c, struct sdmd_wall_dyn *wd) {
    *wd = c->wd;
    for (int k = 0; k < 3; k++) wd->box[k] = c->v->box[k];
}
/* DoWallDyn, Events.c:502-540 */
static void do_wall_dyn(oracle_ctx *c) {
    struct sdmd_host_view *v = c->v;
    struct sdmd_wall_dyn *wd = &c->wd;
    double *box = v->box;
    const double currenttime = c->st.currenttime;
    if (!(currenttime >= wd->cur_time)) return;
    if (wd->mode == 3) {                                                                  /* "pulse", :504-514 */
        if ((wd->orig_box[1] - box[1] >= wd->size) || (wd->orig_box[1] == box[1])) {
            wd->sign *= -1;
            wd->touch = 0;
        } else if (fabs(box[1] + (int)(currenttime / wd->rate + 1) * wd->size - (currenttime - 1) * wd->step - wd->orig_box[1]) <= wd->step &&
                   wd->touch >= 5) {
            wd->sign *= -1;
            wd->touch = 0;
        }
    }
    wd->cur_time += 1;
    if (v->wall_kind == SDMD_WALL_PARALLEL) box[1] += wd->sign * wd->step;                /* :518-527 */
    else if (v->wall_kind == SDMD_WALL_CYLINDER) { box[1] += wd->sign * wd->step; box[2] += wd->sign * wd->step; }
    else { box[0] += wd->sign * wd->step; box[1] += wd->sign * wd->step; box[2] += wd->sign * wd->step; }
    for (int i = 1; i <= c->n; i++)                                                       /* :529-533 */
        if (v->ev_type[i] == SDMD_EV_WALL) v->ev_counter[i]++;
    if (((wd->orig_box[1] - box[1] >= wd->size) && wd->mode == 1) || ((wd->orig_box[1] - box[1] <= wd->step) && wd->mode == 2)) wd->mode = 0;   /* :535-538 */
}

int oracle_run(oracle_ctx *c, double t_until, int64_t max_iter, struct sdmd_trace_rec *trace, int64_t trace_cap,
               struct sdmd_run_stats *stats) {
    struct sdmd_host_view *v = c->v;
    int32_t renew[4];
    int64_t iter = 0;
    c->st.stop_reason = SDMD_STOP_TIME_END;
    while (c->st.currenttime <= v->timestep) {
        if (iter >= max_iter) { c->st.stop_reason = SDMD_STOP_MAX_ITER; break; }
        if (c->st.error) { c->st.stop_reason = SDMD_STOP_ERROR; break; }
        /* AssignJob, SGThread.c:237-259 */
        int target = c->node[1];
        double dt = v->ev_time[target];
        int partner = v->ev_partner[target] > 0 ? v->ev_partner[target] : 0;
        renew[0] = partner ? 2 : 1; renew[1] = target; renew[2] = partner; renew[3] = 0;
        oracle_advance(c, dt);
        c->st.currenttime += dt;
        /* HazardCheck, SGThread.c:215-235 */
        int hazard = (partner > 0 && v->ev_pcounter[target] < v->ev_counter[partner]) ? -1 : 0;
        if (trace && iter < trace_cap) {
            struct sdmd_trace_rec *r = &trace[iter];
            r->frame = c->st.frame; r->target = target; r->partner = v->ev_partner[target];
            r->type = v->ev_type[target]; r->sub = v->ev_sub[target]; r->executed = !hazard; r->pad_ = 0; r->dt = dt;
        }
        iter++;
        c->st.iterations++;
        if (!hazard) {
            c->st.frame++;
            do_event(c, target, partner);
            v->ev_counter[target]++;
            if (partner && !v->is_cons[partner]) v->ev_counter[partner]++;
            oracle_predict(c, renew);
            oracle_cbt_update(c, renew);
            if (c->wd.mode) do_wall_dyn(c);                                              /* SGThread.c:129 */
            if (c->st.currenttime >= t_until) { c->st.stop_reason = SDMD_STOP_T_UNTIL; break; }
        } else {
            renew[0] = 1; renew[2] = 0; renew[3] = 0;
            oracle_predict(c, renew);
            oracle_cbt_update(c, renew);
            c->st.count_recal++;
        }
    }
    v->currenttime = c->st.currenttime;
    v->frame = c->st.frame;
    if (stats) *stats = c->st;
    return c->st.error;
}

/* ------------------------------------------------------------------------------------------------ */
/* CalPotenE, ToolFunctions.c:659-752 */
double oracle_potential_energy(oracle_ctx *c) {
    const struct sdmd_host_view *v = c->v;
    double TPE = 0, accum = 0, tmp;
    double cut2 = v->cutoffr * v->cutoffr;
    for (int i = 1; i <= c->n; i++) {
        const int32_t *ci = &v->cell_idx[3 * i];
        for (int n = 0; n <= 26; n++) {
            int nb[3] = {ci[0] + OFF_X[n], ci[1] + OFF_Y[n], ci[2] + OFF_Z[n]};
            double shift[3];
            for (int d = 0; d < 3; d++) {
                shift[d] = 0;
                if (nb[d] < 0) { nb[d] = c->cellnum[d] - 1; shift[d] = -1 * v->box[d]; }
                else if (nb[d] >= c->cellnum[d]) { nb[d] = 0; shift[d] = v->box[d]; }
            }
            int cell = nb[2] * c->cellnum[0] * c->cellnum[1] + nb[1] * c->cellnum[0] + nb[0] + 1;
            for (int pair = c->celllist[c->n + cell]; pair != 0; pair = c->celllist[pair]) {
                if (pair <= i) continue;
                double co[3] = {v->x[pair] + shift[0], v->y[pair] + shift[1], v->z[pair] + shift[2]};
                double r[3] = {v->x[i] - co[0], v->y[i] - co[1], v->z[i] - co[2]};
                double d2 = r[0] * r[0] + r[1] * r[1] + r[2] * r[2];
                int cm = cm_flags(v, i, pair);
                if (d2 <= cut2 && !(cm & SDMD_BOND_CONNECT)) {
                    int ai = i, aj = pair, hbp = 0, mode;
                    if (v->hb_bond[ai] == aj) { mode = FP_HB; TPE -= hb_barrier(v, ai, aj); }
                    else if (cm & SDMD_CONSTRAINT_CONNECT) mode = FP_CONSTRAINT;
                    else if (neighbor_connected(v, ai, aj)) {
                        if (v->hb_neighbor[v->hb_bond[ai]] == aj) { int s = ai; ai = aj; aj = s; }
                        hbp = v->hb_bond[aj];
                        mode = FP_NEIGHBOR;
                    } else mode = FP_COLLISION;
                    double w[3] = {v->vx[ai] - v->vx[aj], v->vy[ai] - v->vy[aj], v->vz[ai] - v->vz[aj]};
                    double direction = r[0] * w[0] + r[1] * w[1] + r[2] * w[2];
                    find_pair(c, ai, aj, mode, direction, d2, &tmp, &tmp, &tmp, &tmp, &accum, hbp, 0);
                    if (accum < INF / 2) TPE += accum;
                }
            }
        }
    }
    return TPE;
}

/* CalKinetE, ToolFunctions.c:626-656 (flow.mark < 2, constV = 0) */
double oracle_kinetic_energy(oracle_ctx *c) {
    const struct sdmd_host_view *v = c->v;
    double TKE = 0;
    if (v->flow_mark < 2) {                                     /* ToolFunctions.c:630-636: relative to the drift velocity */
        for (int i = 1; i <= c->n; i++) {
            double t[3] = {v->vx[i] - v->flow_v[0], v->vy[i] - v->flow_v[1], v->vz[i] - v->flow_v[2]};
            double s = t[0] * t[0] + t[1] * t[1] + t[2] * t[2];
            TKE += 0.5 * v->mass[i] * s;
        }
    } else {                                                    /* :637-649: relative to each protein's centre of mass (CalCOMV :613-623) */
        int i = 1;
        while (i <= c->n) {
            int j = i;
            double pm = v->protein_mass ? v->protein_mass[i] : 0, net[3] = {0, 0, 0};     /* CalculateMass, Initialization.c:1557-1567 */
            while (j <= c->n && v->protein_num[j] == v->protein_num[i]) j++;
            for (int n = i; n < j; n++) { net[0] += v->vx[n] * v->mass[n]; net[1] += v->vy[n] * v->mass[n]; net[2] += v->vz[n] * v->mass[n]; }
            for (int k = 0; k < 3; k++) net[k] /= pm;
            for (int n = i; n < j; n++) {
                double t[3] = {v->vx[n] - net[0], v->vy[n] - net[1], v->vz[n] - net[2]};
                double s = t[0] * t[0] + t[1] * t[1] + t[2] * t[2];
                TKE += 0.5 * v->mass[n] * s;
            }
            i = j;
        }
    }
    return TKE;
}

/* rearrange, Server.c:145-192 */
int oracle_remd_rearrange(double *epot, struct sdmd_remd_temp *T, int nrep, const double *u01, int n_u01, int *n_used,
                          double *attempt, double *accept) {
    int used = 0;
    for (int i = 0; i < nrep; i++) {
        if (!T[i].T) continue;
        for (int j = 0; j < nrep; j++) {
            if (!T[j].T) continue;
            if ((T[i].num == T[j].num + 1) || (T[i].num == 1 && T[j].num == 2)) {
                double beta1 = 1 / T[i].T, beta2 = 1 / T[j].T;
                double dE = (beta1 - beta2) * (epot[j] - epot[i]);
                int lo = (T[i].num < T[j].num ? T[i].num : T[j].num) - 1;
                if (attempt) attempt[lo]++;
                int ok = dE <= 0.0;
                if (!ok) {
                    if (used >= n_u01) return SDMD_ERR_BAD_ARG;
                    ok = u01[used++] < exp(-1 * dE);
                }
                if (ok) {
                    struct sdmd_remd_temp z = T[i]; T[i] = T[j]; T[j] = z;
                    lo = (T[i].num < T[j].num ? T[i].num : T[j].num) - 1;
                    if (accept) accept[lo]++;
                }
                break;
            }
        }
    }
    if (n_used) *n_used = used;
    return 0;
}

/* Maxwell_Boltzmann_Distribution(0,1), ToolFunctions.c:517-522,544-569 */
double oracle_gaussian_libc(void) {
    double r0, r1, r2, choose, out;
    do {
        do {
            r0 = 2 * (0 + (double)rand() / ((double)RAND_MAX / (1 - 0))) - 1;
            r1 = 2 * (0 + (double)rand() / ((double)RAND_MAX / (1 - 0))) - 1;
            r2 = r0 * r0 + r1 * r1;
        } while (r2 >= 1);
        r2 = sqrt(-2 * log(r2) / r2);
        choose = 0 + (double)rand() / ((double)RAND_MAX / (1 - 0));
        out = (choose >= 0.5) ? r0 * r2 : r1 * r2;
    } while (out == 0);
    return out * sqrt(1.0) + 0;
}
