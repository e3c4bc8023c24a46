/*
 * sdmd_refglue.c — flatten the sDMD process globals into `struct sdmd_host_view` and back.
 * Compiled against the reference's DMD.h (include path supplied by the build), see sdmd_refglue.h.
 */
#include "DMD.h"
#include "sdmd_refglue.h"

#define NT (NATOMTYPE + 1)

/* ------------------------------------------------------------------------------------------------ */
/* potential-table pool                                                                             */
/* ------------------------------------------------------------------------------------------------ */
struct pool {
    int n, cap;
    struct sdmd_step_table *t;
    /* dedupe key: the step-list head pointer plus dmin/dmax (struct copies share the list) */
    struct StepPotenStr **key_step;
    double *key_dmin, *key_dmax;
};

static int pool_add(struct pool *p, const struct ConstraintStr *c, int *err) {
    for (int i = 0; i < p->n; i++) {
        if (p->key_step[i] == c->step && p->key_dmin[i] == c->dmin && p->key_dmax[i] == c->dmax) return i;
    }
    if (p->n == p->cap) {
        p->cap = p->cap ? p->cap * 2 : 256;
        p->t = realloc(p->t, sizeof(*p->t) * p->cap);
        p->key_step = realloc(p->key_step, sizeof(*p->key_step) * p->cap);
        p->key_dmin = realloc(p->key_dmin, sizeof(double) * p->cap);
        p->key_dmax = realloc(p->key_dmax, sizeof(double) * p->cap);
    }
    struct sdmd_step_table *t = &p->t[p->n];
    memset(t, 0, sizeof(*t));
    t->dmin2 = c->dmin;
    t->dmax2 = c->dmax;
    int k = 0;
    for (struct StepPotenStr *s = c->step; s != NULL; s = s->next) {
        if (k >= SDMD_MAX_STEPS || s->d == 0) { *err = SDMD_ERR_UNSUPPORTED; break; }
        t->d2[k] = s->d;
        t->e[k] = s->e;
        t->acc[k] = s->accumulated;
        k++;
    }
    t->nstep = k;
    p->key_step[p->n] = c->step;
    p->key_dmin[p->n] = c->dmin;
    p->key_dmax[p->n] = c->dmax;
    return p->n++;
}

/* ------------------------------------------------------------------------------------------------ */

#define ALLOC(ptr, count) ptr = calloc((size_t)(count), sizeof(*(ptr)))

void sdmd_glue_gather_dynamic(struct sdmd_host_view *v) {
    int n = v->natom;
    v->target_temperature = targetTemperature;
    v->currenttime = currenttime;
    v->timestep = timestep;
    v->frame = frame;
    for (int i = 1; i <= n; i++) {
        struct DynamicStr *d = atom[i].dynamic;
        struct PropertyStr *p = atom[i].property;
        v->x[i] = d->coordinate[1]; v->y[i] = d->coordinate[2]; v->z[i] = d->coordinate[3];
        v->vx[i] = d->velocity[1];  v->vy[i] = d->velocity[2];  v->vz[i] = d->velocity[3];
        v->mass[i] = p->mass;
        v->type[i] = p->typeofAtom;
        v->hb_bond[i] = d->HB.bondConnection;
        v->hb_neighbor[i] = d->HB.neighbor;
        v->hb_role[i] = (uint8_t)d->HB.role;
        v->hb_itype[i] = (uint8_t)d->HB.interactionType;
        for (int k = 0; k < 5; k++) v->hbn_partner[5 * i + k] = d->HBNeighbor.neighborPartner[k];
        v->hbn_partner_num[i] = d->HBNeighbor.partnerNum;
        v->hbn_status[i] = d->HBNeighbor.neighborStatus;
        v->ev_time[i] = d->event.time;
        v->ev_potential[i] = d->event.potential;
        v->ev_partner[i] = d->event.partner;
        v->ev_counter[i] = d->event.counter;
        v->ev_pcounter[i] = d->event.partnerCounter;
        v->ev_type[i] = (uint8_t)d->event.eventType;
        v->ev_sub[i] = (uint8_t)d->event.subEventType;
        for (int k = 0; k < 3; k++) v->cell_idx[3 * i + k] = d->cellIndex[k + 1];
        if (v->flow_time_rec && flow.force.timeRec) v->flow_time_rec[i] = flow.force.timeRec[i];
    }
}

/* HB_CONNECT / NEIGHBOR_CONNECT bits of connectionMap follow the HB lists (HBEvent flips them together,
 * Events.c:297-329,350-382): clear what the previous lists set, set what the new lists say. */
static void rebuild_hb_bits(const struct sdmd_host_view *v) {
    int n = v->natom;
    for (int i = 1; i <= n; i++) {
        struct DynamicStr *d = atom[i].dynamic;
        if (d->HB.bondConnection > 0) connectionMap[i][d->HB.bondConnection] &= ~HB_CONNECT;
        for (int k = 1; k <= 4 && d->HBNeighbor.neighborPartner[k]; k++) connectionMap[i][d->HBNeighbor.neighborPartner[k]] &= ~NEIGHBOR_CONNECT;
    }
    for (int i = 1; i <= n; i++) {
        if (v->hb_bond[i] > 0) connectionMap[i][v->hb_bond[i]] |= HB_CONNECT;
        for (int k = 1; k <= 4 && v->hbn_partner[5 * i + k]; k++) connectionMap[i][v->hbn_partner[5 * i + k]] |= NEIGHBOR_CONNECT;
    }
}

void sdmd_glue_scatter_dynamic(const struct sdmd_host_view *v) {
    int n = v->natom;
    rebuild_hb_bits(v);
    for (int i = 1; i <= n; i++) {
        struct DynamicStr *d = atom[i].dynamic;
        struct PropertyStr *p = atom[i].property;
        d->coordinate[1] = v->x[i]; d->coordinate[2] = v->y[i]; d->coordinate[3] = v->z[i];
        d->velocity[1] = v->vx[i];  d->velocity[2] = v->vy[i];  d->velocity[3] = v->vz[i];
        p->typeofAtom = v->type[i];
        d->HB.bondConnection = v->hb_bond[i];
        d->HB.interactionType = (enum EHBEventType)v->hb_itype[i];
        for (int k = 0; k < 5; k++) d->HBNeighbor.neighborPartner[k] = v->hbn_partner[5 * i + k];
        d->HBNeighbor.partnerNum = v->hbn_partner_num[i];
        d->HBNeighbor.neighborStatus = v->hbn_status[i];
        d->event.time = v->ev_time[i];
        d->event.potential = v->ev_potential[i];
        d->event.partner = v->ev_partner[i];
        d->event.counter = v->ev_counter[i];
        d->event.partnerCounter = v->ev_pcounter[i];
        d->event.eventType = (enum EEventType)v->ev_type[i];
        d->event.subEventType = (enum EInteractionType)v->ev_sub[i];
        for (int k = 0; k < 3; k++) d->cellIndex[k + 1] = v->cell_idx[3 * i + k];
        if (v->flow_time_rec && flow.force.timeRec) flow.force.timeRec[i] = v->flow_time_rec[i];
    }
}

int sdmd_glue_flatten(struct sdmd_host_view *v) {
    int err = 0;
    int n = atomnum;
    memset(v, 0, sizeof(*v));
    v->abi_version = SDMD_ABI_VERSION;
    v->natom = n;
    for (int k = 0; k < 3; k++) v->box[k] = boxDimension[k + 1];
    v->cutoffr = cutoffr;
    v->thermoF = thermoF;
    v->hb_num_formed = HBnumformed;
    v->alpha_hb_formed = alphaHBformed;
    v->thermostat_on = strcmp(thermostatType, "no") != 0;
    /* ThermostatEvent redraws velocities only for a type starting with 'A' (Events.c:404); any other string still schedules
     * Ther_Event but leaves the velocities alone — not a mode the device implements, so refuse it instead of diverging */
    if (v->thermostat_on && strncmp(thermostatType, "Andersen", 1) != 0) return SDMD_ERR_UNSUPPORTED;
    v->method_ding = strncmp(Methodtype, "Ding", 1) == 0;
#ifdef HYDROGEN_BOND
    v->hydrogen_bond = 1;
#endif
#ifdef GEL
    v->gel_mode = 1;
#endif
    /* features that live outside the hot-path scope (rough / moving walls, obstacles, tunnels, flow, CG) */
    if (CG.mark) return SDMD_ERR_UNSUPPORTED;
    if (strcmp(wallDyn.mark, "no")) {            /* moving wall: Initialization.c:1937-1978 has filled wallDyn */
        v->wall_dyn = 1;
        v->wall_dyn_touch = wallDyn.touch;
        v->wall_dyn_step = wallDyn.step;
        for (int k = 0; k < 3; k++) v->wall_dyn_orig_box[k] = wallDyn.origBoxDim[k + 1];
    }
    if (flow.mark == 3) {
        int total = flow.charge.PBCMark ? flow.charge.num : flow.charge.num / 2;     /* TimePrediction.c:936 */
        if (total > SDMD_MAX_CHARGE) return SDMD_ERR_UNSUPPORTED;
        v->n_charge = total;
        v->charge_com = chargeAA[COM];
        for (int n = 0; n < total; n++) {
            for (int k = 0; k < 3; k++) { v->charge_pos[n][k] = flow.charge.position[n][k + 1]; v->charge_vel[n][k] = flow.charge.velocity[n][k + 1]; }
            v->charge_potgrad[n] = flow.charge.potGrad[n];
            v->charge_gap[n] = flow.charge.gap[n];
        }
    }
    v->flow_mark = flow.mark;
    for (int k = 0; k < 3; k++) v->flow_v[k] = flow.constV.v[k + 1];
    /* flat obstruction walls, tunnels through them, hard spherical obstacles */
    if (obstObj.mark) {
        if (obstObj.num > SDMD_MAX_OBST) return SDMD_ERR_UNSUPPORTED;
        v->n_obst = obstObj.num;
        for (int n = 0; n < obstObj.num; n++) {
            v->obst_hole[n] = obstObj.hole[n];
            for (int k = 0; k < 3; k++) v->obst_pos[n][k] = obstObj.position[n][k + 1];
            v->obst_type[n] = obstObj.obst[n].property->typeofAtom;
        }
    }
    if (tunlObj.mark) {
        if (tunlObj.num > SDMD_MAX_OBST) return SDMD_ERR_UNSUPPORTED;
        v->n_tunnel = tunlObj.num;
        v->tunnel_type = tunlObj.tunnel.property->typeofAtom;
        for (int n = 0; n < tunlObj.num; n++) {
            for (int k = 0; k < 4; k++) v->tunnel_pos[n][k] = tunlObj.position[n][k];
            v->tunnel_diameter[n] = tunlObj.diameter[n];
        }
    }
    if (SphObstObj.mark) {
        if (SphObstObj.num > SDMD_MAX_OBST) return SDMD_ERR_UNSUPPORTED;
        v->n_sphobst = SphObstObj.num;
        for (int n = 0; n < SphObstObj.num; n++) {
            for (int k = 0; k < 3; k++) v->sphobst_pos[n][k] = SphObstObj.position[n][k + 1];
            v->sphobst_radius[n] = SphObstObj.radius[n];
        }
    }
    /* smooth confining wall: the pseudo-atom ReadWall built (Initialization.c:1907-1935) */
    if (strcmp(wallExist, "no")) {
        if (strcmp(wallExist, "smooth") || wall == NULL) return SDMD_ERR_UNSUPPORTED;
        v->wall_kind = strncmp(wallType, "parallel", 1) == 0 ? SDMD_WALL_PARALLEL
                     : strncmp(wallType, "cylinder", 1) == 0 ? SDMD_WALL_CYLINDER
                     : strncmp(wallType, "sphere", 1) == 0 ? SDMD_WALL_SPHERE : SDMD_WALL_NONE;
        if (v->wall_kind == SDMD_WALL_NONE) return SDMD_ERR_UNSUPPORTED;
        v->wall_type = wall[0].property->typeofAtom;
        for (int k = 0; k < 3; k++) v->wall_pos[k] = wall[0].dynamic->coordinate[k + 1];
    }

    ALLOC(v->x, n + 1); ALLOC(v->y, n + 1); ALLOC(v->z, n + 1);
    ALLOC(v->vx, n + 1); ALLOC(v->vy, n + 1); ALLOC(v->vz, n + 1);
    ALLOC(v->mass, n + 1); ALLOC(v->type, n + 1); ALLOC(v->aa_type, n + 1);
    ALLOC(v->protein_num, n + 1); ALLOC(v->aa_num, n + 1); ALLOC(v->is_cons, n + 1);
    ALLOC(v->hb_bond, n + 1); ALLOC(v->hb_neighbor, n + 1); ALLOC(v->hb_role, n + 1); ALLOC(v->hb_itype, n + 1); ALLOC(v->hb_name, n + 1);
    ALLOC(v->hbn_partner, 5 * (n + 1)); ALLOC(v->hbn_partner_num, n + 1); ALLOC(v->hbn_status, n + 1);
    ALLOC(v->ev_time, n + 1); ALLOC(v->ev_potential, n + 1); ALLOC(v->ev_partner, n + 1);
    ALLOC(v->ev_counter, n + 1); ALLOC(v->ev_pcounter, n + 1); ALLOC(v->ev_type, n + 1); ALLOC(v->ev_sub, n + 1);
    ALLOC(v->cell_idx, 3 * (n + 1));
    ALLOC(v->flow_a, 3 * (n + 1)); ALLOC(v->flow_time_rec, n + 1); ALLOC(v->protein_mass, n + 1);
    ALLOC(v->charge, n + 1);
    for (int i = 1; i <= n; i++) {
        v->protein_mass[i] = protein[atom[i].property->sequence.proteinNum].mass;
        v->charge[i] = atom[i].property->charge;
    }
    if (flow.force.a)
        for (int i = 1; i <= n; i++)
            for (int k = 0; k < 3; k++) v->flow_a[3 * i + k] = flow.force.a[i][k + 1];
    sdmd_glue_gather_dynamic(v);

    struct pool pool = {0};

    /* static per-atom properties + bond / constraint lists (CSR, list order) */
    ALLOC(v->bond_off, n + 2);
    ALLOC(v->con_off, n + 2);
    ALLOC(v->cm_off, n + 2);
    int nb = 0, nc = 0, ncm = 0;
    for (int i = 1; i <= n; i++) {
        struct PropertyStr *p = atom[i].property;
        v->aa_type[i] = p->typeofAA;
        v->protein_num[i] = p->sequence.proteinNum;
        v->aa_num[i] = p->sequence.aminoacidNum;
        v->is_cons[i] = strcmp(p->extraProperty[1], "_CONS") == 0;
        v->hb_name[i] = strcmp(p->name, "H") == 0 ? 1 : (strcmp(p->name, "O") == 0 ? 2 : 0);   /* CheckAlphaHB, Events.c:1121-1128 */
        for (struct ConstraintStr *b = p->bond; b; b = b->next) nb++;
        for (struct ConstraintStr *c = p->constr; c; c = c->next) nc++;
    }
    /* Sparse rows of connectionMap (no O(N^2) scan of the dense matrix).  A BOND / CONSTRAINT bit is only ever set
     * for a pair that sits in a bond or constraint list (Initialization.c:1176-1214, 1413, 1458), on either side; so the
     * candidates of row i are the partners of i's lists plus every atom that lists i, and the bits are read from the matrix
     * for those pairs only.  Rows come out in ascending j like the dense scan.  SDMD_GLUE_CHECK_CM=1 compares with the full scan. */
    int *cand_off = NULL, *cand = NULL;
    ALLOC(cand_off, n + 2);
    for (int i = 1; i <= n; i++) {
        struct PropertyStr *p = atom[i].property;
        for (struct ConstraintStr *b = p->bond; b; b = b->next) if (b->connection >= 1 && b->connection <= n) { cand_off[i + 1]++; cand_off[b->connection + 1]++; }
        for (struct ConstraintStr *c = p->constr; c; c = c->next) if (c->connection >= 1 && c->connection <= n) { cand_off[i + 1]++; cand_off[c->connection + 1]++; }
    }
    for (int i = 1; i <= n + 1; i++) cand_off[i] += cand_off[i - 1];      /* row i = [cand_off[i], cand_off[i + 1]) */
    ALLOC(cand, cand_off[n + 1] + 1);
    {
        int *fill = NULL;
        ALLOC(fill, n + 2);
        for (int i = 1; i <= n; i++) {
            struct PropertyStr *p = atom[i].property;
            for (struct ConstraintStr *b = p->bond; b; b = b->next) if (b->connection >= 1 && b->connection <= n) {
                cand[cand_off[i] + fill[i]++] = b->connection; cand[cand_off[b->connection] + fill[b->connection]++] = i; }
            for (struct ConstraintStr *c = p->constr; c; c = c->next) if (c->connection >= 1 && c->connection <= n) {
                cand[cand_off[i] + fill[i]++] = c->connection; cand[cand_off[c->connection] + fill[c->connection]++] = i; }
        }
        free(fill);
    }
    for (int i = 1; i <= n; i++) {      /* sort every row (insertion sort: rows hold a handful of entries), drop repeats, keep flagged pairs */
        int *r = cand + cand_off[i], m = cand_off[i + 1] - cand_off[i], u = 0;
        for (int a = 1; a < m; a++) { int key = r[a], b = a; while (b > 0 && r[b - 1] > key) { r[b] = r[b - 1]; b--; } r[b] = key; }
        for (int a = 0; a < m; a++)
            if ((a == 0 || r[a] != r[a - 1]) && (connectionMap[i][r[a]] & (BOND_CONNECT | CONSTRAINT_CONNECT))) r[u++] = r[a];
        for (int a = u; a < m; a++) r[a] = 0;       /* 0 terminates the kept part of the row */
        ncm += u;
    }
    ALLOC(v->bond_j, nb + 1); ALLOC(v->bond_dmin2, nb + 1); ALLOC(v->bond_dmax2, nb + 1);
    ALLOC(v->con_j, nc + 1); ALLOC(v->con_table, nc + 1);
    ALLOC(v->cm_j, ncm + 1); ALLOC(v->cm_flags, ncm + 1);
    nb = nc = ncm = 0;
    for (int i = 1; i <= n; i++) {
        struct PropertyStr *p = atom[i].property;
        v->bond_off[i] = nb;
        for (struct ConstraintStr *b = p->bond; b; b = b->next) {
            v->bond_j[nb] = b->connection;
            v->bond_dmin2[nb] = b->dmin;
            v->bond_dmax2[nb] = b->dmax;
            nb++;
        }
        v->con_off[i] = nc;
        for (struct ConstraintStr *c = p->constr; c; c = c->next) {
            v->con_j[nc] = c->connection;
            v->con_table[nc] = pool_add(&pool, c, &err);
            nc++;
        }
        v->cm_off[i] = ncm;
        for (int a = cand_off[i]; a < cand_off[i + 1] && cand[a]; a++) {
            v->cm_j[ncm] = cand[a]; v->cm_flags[ncm] = connectionMap[i][cand[a]] & (BOND_CONNECT | CONSTRAINT_CONNECT); ncm++;
        }
    }
    free(cand); free(cand_off);
    if (getenv("SDMD_GLUE_CHECK_CM")) {      /* the dense scan the sparse construction replaces, as a cross-check */
        int k = 0, bad = 0;
        for (int i = 1; i <= n && !bad; i++)
            for (int j = 1; j <= n; j++) {
                int f = connectionMap[i][j] & (BOND_CONNECT | CONSTRAINT_CONNECT);
                if (!f) continue;
                if (k >= ncm || v->cm_j[k] != j || v->cm_flags[k] != f || k < v->cm_off[i] - 0) { bad = 1; break; }
                k++;
            }
        if (bad || k != ncm) { fprintf(stderr, "sdmd glue: sparse connection rows differ from the dense connectionMap\n"); return SDMD_ERR_STATE; }
    }
    v->bond_off[n + 1] = nb; v->bond_off[0] = 0;
    v->con_off[n + 1] = nc;  v->con_off[0] = 0;
    v->cm_off[n + 1] = ncm;  v->cm_off[0] = 0;

    /* collision tables with the RightPair fallback folded in (ToolFunctions.c:278-292) */
    ALLOC(v->coll_index, NT * NT);
    for (int a = 0; a < NT; a++) {
        for (int b = 0; b < NT; b++) {
            struct ConstraintStr *c = (a == 0 || b == 0) ? NULL : RightPair(a, b, 0);
            v->coll_index[a * NT + b] = c ? pool_add(&pool, c, &err) : -1;
        }
    }
    /* HB tables (Initialization.c:1572-1669); [class][0][0] holds the first row of the class */
    ALLOC(v->hb_index, SDMD_NHBCLASS * NT * NT);
    for (int t = 0; t < SDMD_NHBCLASS; t++) {
        for (int a = 0; a < NT; a++) {
            for (int b = 0; b < NT; b++) {
                struct ConstraintStr *c = &potentialPairHB[t][a][b];
                int idx = -1;
                if (c->dmin != 0 || c->step != NULL) idx = pool_add(&pool, c, &err);
                v->hb_index[(t * NT + a) * NT + b] = idx;
            }
        }
    }
    v->hb_barrier[0] = HBPotential.BB_v;
    v->hb_barrier[1] = HBPotential.SS_v;
    v->hb_barrier[2] = HBPotential.BS_v;

    v->ntable = pool.n;
    v->tables = pool.t;
    free(pool.key_step); free(pool.key_dmin); free(pool.key_dmax);
    return err;
}

void sdmd_glue_free(struct sdmd_host_view *v) {
    free(v->x); free(v->y); free(v->z); free(v->vx); free(v->vy); free(v->vz);
    free(v->mass); free(v->type); free(v->aa_type); free(v->protein_num); free(v->aa_num); free(v->is_cons);
    free(v->hb_bond); free(v->hb_neighbor); free(v->hb_role); free(v->hb_itype); free(v->hb_name);
    free(v->hbn_partner); free(v->hbn_partner_num); free(v->hbn_status);
    free(v->ev_time); free(v->ev_potential); free(v->ev_partner); free(v->ev_counter); free(v->ev_pcounter);
    free(v->ev_type); free(v->ev_sub); free(v->cell_idx); free(v->flow_a); free(v->flow_time_rec); free(v->protein_mass); free(v->charge);
    free(v->bond_off); free(v->bond_j); free(v->bond_dmin2); free(v->bond_dmax2);
    free(v->con_off); free(v->con_j); free(v->con_table);
    free(v->cm_off); free(v->cm_j); free(v->cm_flags);
    free(v->tables); free(v->coll_index); free(v->hb_index);
    memset(v, 0, sizeof(*v));
}

/* ------------------------------------------------------------------------------------------------ */
/* snapshot writer: "SDMDSNP1" then records {char name[24]; char dtype[8]; uint64 count; data (8-byte padded)} */
/* ------------------------------------------------------------------------------------------------ */
struct sdmd_snap { FILE *f; };

sdmd_snap *sdmd_snap_open(const char *path) {
    FILE *f = fopen(path, "wb");
    if (!f) return NULL;
    sdmd_snap *s = malloc(sizeof(*s));
    s->f = f;
    fwrite("SDMDSNP1", 1, 8, f);
    return s;
}

static size_t dtype_size(const char *dtype) {
    if (!strcmp(dtype, "f8") || !strcmp(dtype, "i8")) return 8;
    if (!strcmp(dtype, "i4")) return 4;
    return 1;
}

void sdmd_snap_put(sdmd_snap *s, const char *name, const char *dtype, const void *data, uint64_t count) {
    char hdr[32];
    memset(hdr, 0, sizeof(hdr));
    strncpy(hdr, name, 23);
    strncpy(hdr + 24, dtype, 7);
    fwrite(hdr, 1, 32, s->f);
    fwrite(&count, 8, 1, s->f);
    size_t bytes = dtype_size(dtype) * count;
    if (bytes) fwrite(data, 1, bytes, s->f);
    static const char zeros[8] = {0};
    if (bytes % 8) fwrite(zeros, 1, 8 - bytes % 8, s->f);
}

void sdmd_snap_put_view(sdmd_snap *s, const struct sdmd_host_view *v) {
    int n = v->natom;
    uint64_t n1 = (uint64_t)n + 1;
    int32_t ints[8] = {v->abi_version, v->natom, v->thermostat_on, v->method_ding, v->hydrogen_bond, v->gel_mode, v->ntable, 0};
    double dbl[12] = {v->box[0], v->box[1], v->box[2], v->cutoffr, v->thermoF, v->target_temperature,
                      v->currenttime, v->timestep, v->hb_barrier[0], v->hb_barrier[1], v->hb_barrier[2], 0};
    int64_t i64[1] = {v->frame};
    sdmd_snap_put(s, "scalars_i4", "i4", ints, 8);
    sdmd_snap_put(s, "scalars_f8", "f8", dbl, 12);
    sdmd_snap_put(s, "scalars_i8", "i8", i64, 1);
    int32_t wall_i[2] = {v->wall_kind, v->wall_type};
    sdmd_snap_put(s, "wall_i4", "i4", wall_i, 2);
    sdmd_snap_put(s, "wall_f8", "f8", v->wall_pos, 3);
    sdmd_snap_put(s, "obst_pos", "f8", v->obst_pos, 3 * (uint64_t)v->n_obst);
    sdmd_snap_put(s, "obst_type", "i4", v->obst_type, (uint64_t)v->n_obst);
    sdmd_snap_put(s, "sphobst_pos", "f8", v->sphobst_pos, 3 * (uint64_t)v->n_sphobst);
    sdmd_snap_put(s, "sphobst_radius", "f8", v->sphobst_radius, (uint64_t)v->n_sphobst);
    sdmd_snap_put(s, "tunnel_pos", "f8", v->tunnel_pos, 4 * (uint64_t)v->n_tunnel);
    sdmd_snap_put(s, "tunnel_diameter", "f8", v->tunnel_diameter, (uint64_t)v->n_tunnel);
    sdmd_snap_put(s, "tunnel_i4", "i4", &v->tunnel_type, 1);
    sdmd_snap_put(s, "obst_hole", "i4", v->obst_hole, (uint64_t)v->n_obst);
#define PUT(field, dt, cnt) sdmd_snap_put(s, #field, dt, v->field, cnt)
    PUT(x, "f8", n1); PUT(y, "f8", n1); PUT(z, "f8", n1);
    PUT(vx, "f8", n1); PUT(vy, "f8", n1); PUT(vz, "f8", n1);
    PUT(mass, "f8", n1); PUT(type, "i4", n1); PUT(aa_type, "i4", n1);
    PUT(protein_num, "i4", n1); PUT(aa_num, "i4", n1); PUT(is_cons, "u1", n1);
    PUT(hb_bond, "i4", n1); PUT(hb_neighbor, "i4", n1); PUT(hb_role, "u1", n1); PUT(hb_itype, "u1", n1); PUT(hb_name, "u1", n1);
    PUT(hbn_partner, "i4", 5 * n1); PUT(hbn_partner_num, "i4", n1); PUT(hbn_status, "i4", n1);
    PUT(ev_time, "f8", n1); PUT(ev_potential, "f8", n1); PUT(ev_partner, "i4", n1);
    PUT(ev_counter, "i8", n1); PUT(ev_pcounter, "i8", n1); PUT(ev_type, "u1", n1); PUT(ev_sub, "u1", n1);
    PUT(cell_idx, "i4", 3 * n1);
    PUT(flow_a, "f8", 3 * n1); PUT(flow_time_rec, "f8", n1); PUT(protein_mass, "f8", n1); PUT(charge, "f8", n1);
    int32_t flow_i[2] = {v->flow_mark, v->charge_com};
    sdmd_snap_put(s, "flow_i4", "i4", flow_i, 2);
    int32_t hb_i[2] = {v->hb_num_formed, v->alpha_hb_formed};
    sdmd_snap_put(s, "hb_i4", "i4", hb_i, 2);
    double csrc[SDMD_MAX_CHARGE][8];
    for (int q = 0; q < v->n_charge; q++) {
        for (int k = 0; k < 3; k++) { csrc[q][k] = v->charge_pos[q][k]; csrc[q][3 + k] = v->charge_vel[q][k]; }
        csrc[q][6] = v->charge_potgrad[q]; csrc[q][7] = v->charge_gap[q];
    }
    sdmd_snap_put(s, "charge_src", "f8", csrc, 8 * (uint64_t)v->n_charge);
    sdmd_snap_put(s, "flow_f8", "f8", v->flow_v, 3);
    PUT(bond_off, "i4", n1 + 1); PUT(bond_j, "i4", v->bond_off[n + 1]);
    PUT(bond_dmin2, "f8", v->bond_off[n + 1]); PUT(bond_dmax2, "f8", v->bond_off[n + 1]);
    PUT(con_off, "i4", n1 + 1); PUT(con_j, "i4", v->con_off[n + 1]); PUT(con_table, "i4", v->con_off[n + 1]);
    PUT(cm_off, "i4", n1 + 1); PUT(cm_j, "i4", v->cm_off[n + 1]); PUT(cm_flags, "i4", v->cm_off[n + 1]);
    sdmd_snap_put(s, "tables", "u1", v->tables, (uint64_t)v->ntable * sizeof(struct sdmd_step_table));
    PUT(coll_index, "i4", NT * NT);
    PUT(hb_index, "i4", SDMD_NHBCLASS * NT * NT);
#undef PUT
}

void sdmd_snap_close(sdmd_snap *s) {
    char hdr[32];
    memset(hdr, 0, sizeof(hdr));
    strcpy(hdr, "END");
    fwrite(hdr, 1, 32, s->f);
    uint64_t zero = 0;
    fwrite(&zero, 8, 1, s->f);
    fclose(s->f);
    free(s);
}
