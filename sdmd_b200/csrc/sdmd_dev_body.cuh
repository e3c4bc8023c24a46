// Device functions of the hot path.  Included twice by sdmd_dev.cuh: once with SDMD_LZ 0 (namespace lz0: every kernel but
// one) and once with SDMD_LZ 1 (namespace lz1: the event block of the overlapped sweep, which reads other atoms' positions
// and event times through the on-the-fly advance).  No include guard on purpose.
__device__ __forceinline__ bool d_lz(const DevP &D, int j) { return SDMD_LZ && j != D.lz_a && j != D.lz_b; }
__device__ __forceinline__ double d_rx(const DevP &D, int j) { return d_lz(D, j) ? D.lz_x[j] + D.vx[j] * D.lz_dt : D.x[j]; }
__device__ __forceinline__ double d_ry(const DevP &D, int j) { return d_lz(D, j) ? D.lz_y[j] + D.vy[j] * D.lz_dt : D.y[j]; }
__device__ __forceinline__ double d_rz(const DevP &D, int j) { return d_lz(D, j) ? D.lz_z[j] + D.vz[j] * D.lz_dt : D.z[j]; }
__device__ __forceinline__ double d_rt(const DevP &D, int j) { return d_lz(D, j) ? D.lz_t[j] - D.lz_dt : D.ev_time[j]; }
__device__ __forceinline__ void d_prefetch_l1(const void *p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }
// everything the evaluation of a collision candidate reads of its partner, asked into this SM's L1 ahead of time
__device__ __forceinline__ void d_prefetch_atom(const DevP &D, int j) {
#if SDMD_LZ
    if (D.lz_on) { d_prefetch_l1(&D.lz_x[j]); d_prefetch_l1(&D.lz_y[j]); d_prefetch_l1(&D.lz_z[j]); }
    else
#endif
    { d_prefetch_l1(&D.x[j]); d_prefetch_l1(&D.y[j]); d_prefetch_l1(&D.z[j]); }
    d_prefetch_l1(&D.vx[j]); d_prefetch_l1(&D.vy[j]); d_prefetch_l1(&D.vz[j]);
    d_prefetch_l1(&D.ev_partner[j]); d_prefetch_l1(&D.ev_type[j]); d_prefetch_l1(&D.ev_pcounter[j]); d_prefetch_l1(&D.type[j]);
}
// box edge k: a kernel parameter, except under moving walls (WallDyn), where DoWallDyn and WallTime change it during the run
__device__ __forceinline__ double d_box(const DevP &D, int k) { return D.wall_dyn ? D.ctl->dbox[k] : D.box[k]; }
// event time of another atom's slot: the live slots go through d_rt, the relaxation buffers of Predict(all) do not
__device__ __forceinline__ double d_stime(const DevP &D, const SlotSrc &S, int j) { return S.time == D.ev_time ? d_rt(D, j) : S.time[j]; }

// neighbour-cell offsets, DMD.c:118-120
static __constant__ int c_offx[27] = {0, 1, 0, 0, 0, 1, 1, 1, -1, 0, 0, 0, -1, -1, -1, 0, 0, 1, -1, 1, -1, -1, -1, 1, 1, 1, -1};
static __constant__ int c_offy[27] = {0, 0, 1, 0, 1, 0, 1, 1, 0, -1, 0, -1, 0, -1, -1, -1, 1, 0, 0, -1, 1, -1, 1, -1, 1, -1, 1};
static __constant__ int c_offz[27] = {0, 0, 0, 1, 1, 1, 0, 1, 0, 0, -1, -1, -1, 0, -1, 1, -1, -1, 1, 0, 0, 1, -1, -1, -1, 1, 1};

enum { FP_COLLISION = 0, FP_CONSTRAINT = 1, FP_HB = 2, FP_NEIGHBOR = 3 };

struct FindRes {
    double lower, upper, lowPot, upPot, accum;
    int cs;  // coreorShell
};

struct Pair {
    double rx, ry, rz, r2, b, v2;
};

__device__ __forceinline__ FindRes d_find_pair_table_fast(const DevP &D, const sdmd_step_table *t, int mode, double direction,
                                                          double dist2, int bond_a1, int bond_a2);

// ---------------------------------------------------------------------------------------------------
// AtomTypeChange, ToolFunctions.c:233-270
__device__ __forceinline__ int d_type_change(int t, int direct) {
    if (direct == 1) {
        if (t == 1) return 3;
        if (t == 2) return 4;
        if (t == 26) return 27;
    } else {
        if (t == 3) return 1;
        if (t == 4) return 2;
        if (t == 27) return 26;
    }
    return t;
}

// HBModel, Models.c:212-290; -1 "no HB" (ASN/GLN), -2 = the reference's fatal exit
__device__ inline int d_hb_model(const DevP &D, int a1, int a2) {
    if (D.hb_role[a1] == 'A' || D.hb_role[a1] == 'S') { int s = a1; a1 = a2; a2 = s; }
    int o1 = d_type_change(D.type[a1], 0), o2 = d_type_change(D.type[a2], 0);
    int aa = D.aa_type[a2];
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

// static bits of connectionMap[i][j] (BOND / CONSTRAINT)
__device__ __forceinline__ int d_cm_flags(const DevP &D, int i, int j) {
    for (int k = D.cm_off[i]; k < D.cm_off[i + 1]; k++)
        if (D.cm_j[k] == j) return D.cm_flags[k];
    return 0;
}

// FindElemInList(neighborPartner, value, -1, -1, 0), ToolFunctions.c:1078-1084 (scan bounded to [1,4])
__device__ __forceinline__ int d_find_in_hbn(const DevP &D, int a, int value) {
    for (int k = 1; k <= 4; k++) {
        int p = D.hbn_partner[5 * a + k];
        if (p == 0) return 0;
        if (p == value) return k;
    }
    return 0;
}
// NEIGHBOR_CONNECT bit: set exactly while b is in a's neighborPartner list (Events.c:297-329,350-382)
__device__ __forceinline__ int d_neighbor_connected(const DevP &D, int a, int b) { return d_find_in_hbn(D, a, b) != 0; }

// the step walk of FindPair, ToolFunctions.c:104-224, on one flattened table
__device__ inline FindRes d_find_pair_table(const DevP &D, const sdmd_step_table *t, int mode, double direction, double dist2,
                                            int bond_a1, int bond_a2) {
    FindRes r;
    r.cs = -1;
    double dir = (direction < 0) ? 1.0 : -1.0;  // :46-50
    double tol = dir * SD_ZERO;
    if (dist2 < t->dmin2 - SD_ZERO) {  // :119-139
        r.lower = SD_INF; r.upper = SD_INF; r.accum = SD_INF; r.lowPot = 0; r.upPot = 0;
        if (!(mode == FP_NEIGHBOR && bond_a2 == 0) && !(mode == FP_HB && bond_a1 == 0) && !D.n_tunnel)   /* :127-130: not under tunnels */ atomicAdd((unsigned long long *)&D.ctl->warnings, 1ULL);
        return r;
    }
    r.lower = t->dmin2;
    r.lowPot = 0;
    r.upper = 0; r.upPot = 0; r.accum = 0;
    int ns = t->nstep;
    for (int k = 0; k < ns; k++) {  // :144-169
        double d = t->d2[k];
        if (dist2 < d + tol) {
            r.upper = d;
            r.upPot = -1 * t->e[k];
            r.accum = t->acc[k];
            if (k == 0) { r.cs = 0; return r; }
            if (mode == FP_HB && k == ns - 1 && bond_a1 > 0) { r.cs = 1; return r; }
            if (mode == FP_NEIGHBOR && k == ns - 1 && bond_a2 > 0) { r.upPot = 0; r.cs = 1; return r; }
            return r;
        }
        r.lower = d;
        r.lowPot = t->e[k];
    }
    if (mode == FP_HB && bond_a1 > 0) atomicAdd((unsigned long long *)&D.ctl->warnings, 1ULL);  // :171-178
    if (t->dmax2 > 0) {  // :180-211
        if (dist2 < t->dmax2 + SD_ZERO) {
            r.upper = t->dmax2; r.upPot = 0; r.accum = 0;
            r.cs = 1;
            if (ns == 0 && dir == 1.0) r.cs = 0;
            return r;
        }
        r.lower = SD_INF; r.upper = SD_INF; r.accum = SD_INF; r.lowPot = 0; r.upPot = 0;
        if (!(mode == FP_NEIGHBOR && bond_a2 == 0) && !(mode == FP_HB && bond_a1 == 0)) atomicAdd((unsigned long long *)&D.ctl->warnings, 1ULL);
        return r;
    }
    r.upper = SD_INF; r.upPot = SD_INF; r.accum = SD_INF;  // :212-221
    r.cs = 1;
    return r;
}

// table selection of FindPair, ToolFunctions.c:53-102
__device__ inline FindRes d_find_pair(const DevP &D, int a1, int a2, int mode, double direction, double dist2, int hbPartner,
                                      int typeChange) {
    int idx = -1;
    if (mode == FP_COLLISION) {
        int t1 = D.type[a1];
        if (typeChange == 1) t1 = d_type_change(t1, 1);
        else if (typeChange == -1) t1 = d_type_change(t1, 0);
        idx = D.coll_index[t1 * SD_NT + D.type[a2]];
    } else if (mode == FP_CONSTRAINT) {
        for (int k = D.con_off[a1]; k < D.con_off[a1 + 1]; k++)
            if (D.con_j[k] == a2) { idx = D.con_table[k]; break; }
    } else if (mode == FP_HB) {
        int cls = d_hb_model(D, a1, a2);
        if (cls >= 0) idx = D.hb_index[(cls * SD_NT + 0) * SD_NT + 0];
    } else {
        int cls = hbPartner > 0 ? d_hb_model(D, a2, hbPartner) : -2;
        int t1 = d_type_change(D.type[a1], 0), t2 = d_type_change(D.type[a2], 0);
        if (cls >= 0) idx = D.hb_index[(cls * SD_NT + t1) * SD_NT + t2];
    }
    if (idx < 0 || D.tables[idx].dmin2 == 0) {  // fatal in the reference, :114-117
        D.ctl->error = SDMD_ERR_TABLE;
        FindRes r;
        r.lower = SD_INF; r.upper = SD_INF; r.accum = SD_INF; r.lowPot = 0; r.upPot = 0; r.cs = -1;
        return r;
    }
    return d_find_pair_table(D, &D.tables[idx], mode, direction, dist2, D.hb_bond[a1], D.hb_bond[a2]);
}

// PBCShift, ToolFunctions.c:801-817, from the two cell-index triples
__device__ __forceinline__ double d_shift1(int ci, int cj, int cn, double box) {
    if (ci != cj) {
        if (ci == 0 && cj == cn - 1) return -1 * box;
        if (ci == cn - 1 && cj == 0) return box;
    }
    return 0;
}

// CalculateParameters, TimePrediction.c:1081-1090
__device__ __forceinline__ Pair d_pair(double xi, double yi, double zi, double vxi, double vyi, double vzi, double xj, double yj,
                                       double zj, double vxj, double vyj, double vzj, double shx, double shy, double shz) {
    Pair p;
    double tx = xj + shx, ty = yj + shy, tz = zj + shz;
    p.rx = xi - tx; p.ry = yi - ty; p.rz = zi - tz;
    double wx = vxi - vxj, wy = vyi - vyj, wz = vzi - vzj;
    p.r2 = p.rx * p.rx + p.ry * p.ry + p.rz * p.rz;
    p.b = p.rx * wx + p.ry * wy + p.rz * wz;
    p.v2 = wx * wx + wy * wy + wz * wz;
    return p;
}
__device__ inline Pair d_pair_atoms(const DevP &D, int i, int j) {
    const int *ci = &D.cell_idx[3 * i], *cj = &D.cell_idx[3 * j];
    double shx = d_shift1(ci[0], cj[0], D.cellnum[0], d_box(D, 0));
    double shy = d_shift1(ci[1], cj[1], D.cellnum[1], d_box(D, 1));
    double shz = d_shift1(ci[2], cj[2], D.cellnum[2], d_box(D, 2));
    return d_pair(d_rx(D, i), d_ry(D, i), d_rz(D, i), D.vx[i], D.vy[i], D.vz[i], d_rx(D, j), d_ry(D, j), d_rz(D, j), D.vx[j], D.vy[j], D.vz[j], shx, shy, shz);
}
// CalculateDisc :1092-1094 / CalculateTime :1096-1099
__device__ __forceinline__ double d_disc(const Pair &p, double short2) { return p.b * p.b - p.v2 * (p.r2 - short2); }
__device__ __forceinline__ double d_time(const Pair &p, int s, double d2) {
    return (-1 * p.b + s * sqrt(p.b * p.b - p.v2 * (p.r2 - d2))) / p.v2;
}

// CheckDuplication, TimePrediction.c:1101-1117 (j is eligible only below dup_limit, see predict_all)
__device__ __forceinline__ bool d_check_dup(const DevP &D, const SlotSrc &S, Slot &cur, int i, i64 counter_i, int j, int type,
                                            int dup_limit) {
    if (j < dup_limit && S.partner[j] == i && S.type[j] == type && S.pcounter[j] == counter_i && cur.time > d_stime(D, S, j)) {
        cur.partner = j;
        cur.pot = S.pot[j];
        cur.sub = S.sub[j];
        cur.time = d_stime(D, S, j);
        cur.type = type;
        cur.pcounter = D.ev_counter[j];
        return true;
    }
    return false;
}
// JobAssign, TimePrediction.c:1140-1158 (j == 0: no partner)
__device__ __forceinline__ bool d_job_assign(const DevP &D, Slot &cur, int j, double time, double pot, int sub, int type) {
    if (time > 0 && !(cur.time > time) && time < cur.m2) cur.m2 = time;
    if (time > 0 && j != 0) {
        if (time < cur.bp_time) { cur.sp_time = cur.bp_time; cur.bp_time = time; cur.bp_j = j; }
        else if (time < cur.sp_time) cur.sp_time = time;
    }
    if (time > 0 && cur.time > time) {
        cur.m2 = cur.time;
        cur.pot = pot;
        cur.sub = sub;
        cur.time = time;
        cur.type = type;
        if (j) { cur.partner = j; cur.pcounter = D.ev_counter[j]; }
        else cur.partner = SDMD_INVALID;
        return true;
    }
    return false;
}

__device__ __forceinline__ int d_neighbor_cell(const DevP &D, int cx, int cy, int cz, int k, int *ncx, int *ncy, int *ncz) {
    // DMD.c:118-120 offsets, 2 bits per slot (value + 1) packed per axis, so a warp whose lanes take different k
    // does not serialise on the constant bank
    const unsigned long long PX = 0xa80894054a959ULL, PY = 0x2222160111a665ULL, PZ = 0x28096084059a95ULL;
    int nx = cx + (int)((PX >> (2 * k)) & 3) - 1, ny = cy + (int)((PY >> (2 * k)) & 3) - 1, nz = cz + (int)((PZ >> (2 * k)) & 3) - 1;  // TimePrediction.c:62-72
    if (nx < 0) nx = D.cellnum[0] - 1; else if (nx >= D.cellnum[0]) nx = 0;
    if (ny < 0) ny = D.cellnum[1] - 1; else if (ny >= D.cellnum[1]) ny = 0;
    if (nz < 0) nz = D.cellnum[2] - 1; else if (nz >= D.cellnum[2]) nz = 0;
    *ncx = nx; *ncy = ny; *ncz = nz;
    return nz * D.cellnum[0] * D.cellnum[1] + ny * D.cellnum[0] + nx;  // 0-based
}

// neighbour cell + PBCShift (ToolFunctions.c:801-817) of scan slot k, from the static table when present
__device__ __forceinline__ int d_neighbor(const DevP &D, int cx, int cy, int cz, int k, double *shx, double *shy, double *shz) {
    int ncx, ncy, ncz;
    int cell = d_neighbor_cell(D, cx, cy, cz, k, &ncx, &ncy, &ncz);
    *shx = d_shift1(cx, ncx, D.cellnum[0], d_box(D, 0));
    *shy = d_shift1(cy, ncy, D.cellnum[1], d_box(D, 1));
    *shz = d_shift1(cz, ncz, D.cellnum[2], d_box(D, 2));
    return cell;
}

struct Self {
    int i;
    double x, y, z, vx, vy, vz;
    int cx, cy, cz;
    i64 counter;
    int npair;
    int type, cm0, cm1;
};

// one candidate of CollisionTime's inner loop, TimePrediction.c:82-176, after the duplicate check
__device__ inline void d_collision_candidate(const DevP &D, Self &me, Slot &cur, int j, double xj, double yj, double zj, double vxj,
                                             double vyj, double vzj, double shx, double shy, double shz) {
    int i = me.i;
    int cm = (D.cm_off[i + 1] > D.cm_off[i]) ? d_cm_flags(D, i, j) : 0;
    if ((cm & SDMD_BOND_CONNECT) || D.hb_bond[i] == j || d_find_in_hbn(D, i, j)) return;  // :95-100
    Pair p = d_pair(me.x, me.y, me.z, me.vx, me.vy, me.vz, xj, yj, zj, vxj, vyj, vzj, shx, shy, shz);
    me.npair++;
    if (!(p.r2 <= D.cut2)) return;  // :107
    FindRes f;
    if (cm & SDMD_CONSTRAINT_CONNECT) {
        f = d_find_pair(D, i, j, FP_CONSTRAINT, p.b, p.r2, 0, 0);
    } else {
        f = d_find_pair(D, i, j, FP_COLLISION, p.b, p.r2, 0, 0);
        if (f.lowPot == 0 && f.upPot == 0) return;  // :121-124
    }
    int s, sub;
    double d2, pot;
    if (p.b < 0 && d_disc(p, f.lower) > 0) {  // :130-141
        s = -1; d2 = f.lower; pot = f.lowPot;
        sub = (f.cs == 0) ? SDMD_SUB_CORECOLLI : SDMD_SUB_TBDCAPTR;
    } else {  // :142-164
        s = 1; d2 = f.upper; pot = f.upPot;
        sub = SDMD_SUB_INVALID;
        if (D.method_ding) {
            if (cm & SDMD_CONSTRAINT_CONNECT) sub = (f.cs == 1) ? SDMD_SUB_WELLBOUNC : SDMD_SUB_TBDESCAP;
            else {
                if (f.cs == 1) return;
                sub = SDMD_SUB_TBDESCAP;
            }
        }
    }
    double t = d_time(p, s, d2);
    if (t < 0) atomicAdd((unsigned long long *)&D.ctl->error_prints, 1ULL);  // :168-173 prints and carries on
    d_job_assign(D, cur, j, t, pot, sub, SDMD_EV_COLI);
}

// CollisionTime, TimePrediction.c:44-183, walking the live link list
__device__ inline void d_collision_time(const DevP &D, const SlotSrc &S, Self &me, Slot &cur, int dup_limit) {
    for (int k = 0; k < 27; k++) {
        int ncx, ncy, ncz;
        int cell = d_neighbor_cell(D, me.cx, me.cy, me.cz, k, &ncx, &ncy, &ncz);
        double shx = d_shift1(me.cx, ncx, D.cellnum[0], d_box(D, 0));
        double shy = d_shift1(me.cy, ncy, D.cellnum[1], d_box(D, 1));
        double shz = d_shift1(me.cz, ncz, D.cellnum[2], d_box(D, 2));
        for (int j = D.celllist[D.n + cell + 1]; j; j = D.celllist[j]) {
            if (j == me.i) continue;
            if (d_check_dup(D, S, cur, me.i, me.counter, j, SDMD_EV_COLI, dup_limit)) continue;
            d_collision_candidate(D, me, cur, j, d_rx(D, j), d_ry(D, j), d_rz(D, j), D.vx[j], D.vy[j], D.vz[j], shx, shy, shz);
        }
    }
}

// CollisionTime for the full-system pass, round 0 (no duplicate source): the 27-cell scan is split into
//   stage 1  a uniform loop over the cell-sorted tiles that applies the skip filters (TimePrediction.c:95-100),
//            counts the pair evaluation (:105) and keeps only candidates inside the 0.9*cutoff sphere (:107);
//   stage 2  the expensive step-table walk + quadratic root, run on the compacted survivors only.
// Candidates are still folded in exact scan order, so the result is the sequential one; the split only removes
// the per-candidate divergence that left ~9 of 32 lanes active in the fused loop.
// round-0 evaluation variants (measured in profiles/r1_summary.md, session 4):
//   K2_DIRECT  1 = round 0 evaluates straight from the filter's survivor row (no staging in local arrays), 2 = same with the
//              next record prefetched (slower: 14 more live registers)
//   K2_LEAN    1 = systems without bonds / hydrogen bonds / single-body sources get a kernel instance with those predictors
//              compiled out; it fits 5 blocks per SM without spilling
#ifndef K2_DIRECT
#define K2_DIRECT 1
#endif
#ifndef K2_LEAN
#define K2_LEAN 1
#endif
//   K2_STREAM  1 = the survivor rows (read once) are loaded with the evict-first hint (`ld.global.cs`) so they do not push the
//              64-byte records, which every neighbour re-reads, out of L2: measured 5 % SLOWER (301 vs 287 us), left off
#ifndef K2_STREAM
#define K2_STREAM 0
#endif
#if K2_STREAM
#define K2_LD_ROW(p) __ldcs(p)
#else
#define K2_LD_ROW(p) __ldg(p)
#endif
#define R0_MAXSURV 20
#ifndef R0_SURV_CAP
#define R0_SURV_CAP 32   // most entries per survivor row (128 B; DevP::surv_cap is the row size in use); a longer list marks the row overflowed -> exact in-kernel fallback
#endif
#define R0_DUPFLAG 0x80   // survivor entry whose partner's slot points back: resolved in order in stage 2
// Own-side evaluation of one collision candidate inside the scan sphere: exact distance re-test, step-table walk,
// discriminant, root (CollisionTime, TimePrediction.c:107-175).  Returns false when the candidate produces no event.
__device__ __forceinline__ bool d_eval_pair(const DevP &D, const Self &me, int j, int type_j, double xj, double yj, double zj, double vxj,
                                            double vyj, double vzj, double shx, double shy, double shz, int cm, int type_i, double *t_out,
                                            double *pot_out, int *sub_out) {
    Pair pr = d_pair(me.x, me.y, me.z, me.vx, me.vy, me.vz, xj, yj, zj, vxj, vyj, vzj, shx, shy, shz);
    if (!(pr.r2 <= D.cut2)) return false;
    FindRes f;
    if (cm & SDMD_CONSTRAINT_CONNECT) {
        f = d_find_pair(D, me.i, j, FP_CONSTRAINT, pr.b, pr.r2, 0, 0);
    } else {
        int idx = __ldg(&D.coll_index[type_i * SD_NT + type_j]);
        if (idx < 0) { D.ctl->error = SDMD_ERR_TABLE; return false; }
        f = d_find_pair_table_fast(D, &D.tables[idx], FP_COLLISION, pr.b, pr.r2, 0, 0);
        if (f.lowPot == 0 && f.upPot == 0) return false;
    }
    int sgn, sub;
    double d2, pot;
    if (pr.b < 0 && d_disc(pr, f.lower) > 0) {
        sgn = -1; d2 = f.lower; pot = f.lowPot;
        sub = (f.cs == 0) ? SDMD_SUB_CORECOLLI : SDMD_SUB_TBDCAPTR;
    } else {
        sgn = 1; d2 = f.upper; pot = f.upPot;
        sub = SDMD_SUB_INVALID;
        if (D.method_ding) {
            if (cm & SDMD_CONSTRAINT_CONNECT) sub = (f.cs == 1) ? SDMD_SUB_WELLBOUNC : SDMD_SUB_TBDESCAP;
            else {
                if (f.cs == 1) return false;
                sub = SDMD_SUB_TBDESCAP;
            }
        }
    }
    double t = d_time(pr, sgn, d2);
    if (t < 0) atomicAdd((unsigned long long *)&D.ctl->error_prints, 1ULL);
    *t_out = t; *pot_out = pot; *sub_out = sub;
    return true;
}

__device__ __forceinline__ void d_r0_flush(const DevP &D, const SlotSrc &S, Self &me, Slot &cur, const int *surv_p,
                                           const unsigned char *surv_k, int ns, bool has_cm, int type_i, int hb_i, int h0, int h1,
                                           int h2, int h3) {
    const bool seam = me.cx == 0 || me.cy == 0 || me.cz == 0 || me.cx == D.cellnum[0] - 1 || me.cy == D.cellnum[1] - 1 ||
                      me.cz == D.cellnum[2] - 1;
    for (int s = 0; s < ns; s++) {
        int p = surv_p[s];
        int k = surv_k[s] & 0x1f;
        const double2 *rp = reinterpret_cast<const double2 *>(&D.srec[p]);
        double2 r0 = __ldg(rp), r1 = __ldg(rp + 1), r2v = __ldg(rp + 2);
        int2 ta = __ldg(reinterpret_cast<const int2 *>(rp + 3));
        int j = ta.y;
        double shx = 0, shy = 0, shz = 0;   // PBCShift via the scanned neighbour cell; only atoms in a boundary cell can need one
        if (seam) {
            int ncx, ncy, ncz;
            d_neighbor_cell(D, me.cx, me.cy, me.cz, k, &ncx, &ncy, &ncz);
            shx = d_shift1(me.cx, ncx, D.cellnum[0], d_box(D, 0));
            shy = d_shift1(me.cy, ncy, D.cellnum[1], d_box(D, 1));
            shz = d_shift1(me.cz, ncz, D.cellnum[2], d_box(D, 2));
        }
        int cm = has_cm ? d_cm_flags(D, me.i, j) : 0;
        if (surv_k[s] & R0_DUPFLAG) {
            // CheckDuplication (TimePrediction.c:1101-1117) at this candidate's place in the scan
            if (cur.time > d_stime(D, S, j)) {
                cur.partner = j; cur.pot = S.pot[j]; cur.sub = S.sub[j]; cur.time = d_stime(D, S, j); cur.type = SDMD_EV_COLI;
                cur.pcounter = D.ev_counter[j];
                continue;
            }
            // not copied: the candidate goes down the ordinary path (skip filters, evaluation count, distance test)
            if ((cm & SDMD_BOND_CONNECT) || hb_i == j) continue;
            if ((h0 == j) || (h0 != 0 && (h1 == j || (h1 != 0 && (h2 == j || (h2 != 0 && h3 == j)))))) continue;
            me.npair++;
        }
        double t, pot;
        int sub;
        if (!d_eval_pair(D, me, j, ta.x, r0.x, r0.y, r1.x, r1.y, r2v.x, r2v.y, shx, shy, shz, cm, type_i, &t, &pot, &sub)) continue;
        d_job_assign(D, cur, j, t, pot, sub, SDMD_EV_COLI);
    }
}

template <bool DUP>
__device__ inline void d_collision_time_2stage(const DevP &D, const SlotSrc &S, Self &me, Slot &cur, int dup_limit) {
    int surv_p[R0_MAXSURV];
    unsigned char surv_k[R0_MAXSURV];
    int ns = 0;
    int i = me.i;
    bool has_cm = D.cm_off[i + 1] > D.cm_off[i];
    int hb_i = D.hb_bond[i];
    int h0 = 0, h1 = 0, h2 = 0, h3 = 0;
    if (D.hydrogen_bond) { h0 = D.hbn_partner[5 * i + 1]; h1 = D.hbn_partner[5 * i + 2]; h2 = D.hbn_partner[5 * i + 3]; h3 = D.hbn_partner[5 * i + 4]; }
    int type_i = D.type[i];
    int npair = 0;
    for (int k = 0; k < 27; k++) {
        int ncx, ncy, ncz;
        int cell = d_neighbor_cell(D, me.cx, me.cy, me.cz, k, &ncx, &ncy, &ncz);
        double shx = d_shift1(me.cx, ncx, D.cellnum[0], d_box(D, 0));
        double shy = d_shift1(me.cy, ncy, D.cellnum[1], d_box(D, 1));
        double shz = d_shift1(me.cz, ncz, D.cellnum[2], d_box(D, 2));
        int p1 = D.cell_start[cell + 1];
        for (int p = D.cell_start[cell]; p < p1; p++) {
            int j = D.sorted_atom[p];
            if (j == i) continue;
            bool keep;
            unsigned char tag = (unsigned char)k;
            if (DUP && j < dup_limit && S.partner[j] == i && S.type[j] == SDMD_EV_COLI && S.pcounter[j] == me.counter) {
                keep = true;            // outcome depends on the running best: decide in order, in stage 2
                tag |= R0_DUPFLAG;
            } else {
                if (has_cm && (d_cm_flags(D, i, j) & SDMD_BOND_CONNECT)) continue;
                if (hb_i == j) continue;
                if ((h0 == j) || (h0 != 0 && (h1 == j || (h1 != 0 && (h2 == j || (h2 != 0 && h3 == j)))))) continue;
                npair++;
                double tx = D.srec[p].x + shx, ty = D.srec[p].y + shy, tz = D.srec[p].z + shz;
                double rx = me.x - tx, ry = me.y - ty, rz = me.z - tz;
                double r2 = rx * rx + ry * ry + rz * rz;
                keep = r2 <= D.cut2;
            }
            if (keep) {
                if (ns == R0_MAXSURV) { d_r0_flush(D, S, me, cur, surv_p, surv_k, ns, has_cm, type_i, hb_i, h0, h1, h2, h3); ns = 0; }
                surv_p[ns] = p; surv_k[ns] = tag; ns++;
            }
        }
    }
    d_r0_flush(D, S, me, cur, surv_p, surv_k, ns, has_cm, type_i, hb_i, h0, h1, h2, h3);
    me.npair += npair;
}

// Round-0 stage 1 as its own light kernel body: skip filters, evaluation count, 0.9*cutoff sphere test over the
// cell-sorted tiles; survivors go to HBM in scan order (they are few: ~6 of ~41 candidates at 8 A spacing).
// neighbour-cell offsets of scan slot k (DMD.c:118-120) as floats: the loop counter is warp-uniform, so these are
// constant-bank operands
static __constant__ float c_koff[27][3] = {
    {0, 0, 0},   {1, 0, 0},   {0, 1, 0},   {0, 0, 1},   {0, 1, 1},   {1, 0, 1},  {1, 1, 0},   {1, 1, 1},   {-1, 0, 0},
    {0, -1, 0},  {0, 0, -1},  {0, -1, -1}, {-1, 0, -1}, {-1, -1, 0}, {-1, -1, -1}, {0, -1, 1}, {0, 1, -1},  {1, 0, -1},
    {-1, 0, 1},  {1, -1, 0},  {-1, 1, 0},  {-1, -1, 1}, {-1, 1, -1}, {1, -1, -1}, {1, 1, -1},  {1, -1, 1},  {-1, 1, 1}};

// The filter only has to keep a SUPERSET of the candidates inside the scan sphere (the evaluation stage re-tests every
// survivor in double precision with the reference's arithmetic), so it works in single precision on cell-local
// coordinates: |coordinate| < 2 cells, error of r2 below 1e-3 A^2 against a margin of 1e-4 * cut2 (~7e-3 A^2 at 9 A).
// Relative to the own cell corner a neighbour cell is an offset of whole cells; no PBCShift is needed.
__device__ inline void d_r0_filter(const DevP &D, int p_self) {
    float4 me = __ldg(&D.sloc[p_self]);
    int i = __float_as_int(me.w);
    int c = D.scell[p_self];
    bool has_cm = D.cm_off[i + 1] > D.cm_off[i];
    int hb_i = 0;
    int h0 = 0, h1 = 0, h2 = 0, h3 = 0;
    if (D.hydrogen_bond) { hb_i = D.hb_bond[i]; h0 = D.hbn_partner[5 * i + 1]; h1 = D.hbn_partner[5 * i + 2]; h2 = D.hbn_partner[5 * i + 3]; h3 = D.hbn_partner[5 * i + 4]; }
    bool need_j = has_cm || hb_i != 0 || h0 != 0;
    const int cap = D.surv_cap;
    unsigned int *out = D.surv + (size_t)p_self * cap;
    int ns = 0, npair = 0;
    int cx = 0, cy = 0, cz = 0;
    // neighbour cells by arithmetic (a per-cell table of the 27 ids costs a dependent load per cell and 108 B/cell of
    // traffic).  The loop over the scan slots is fully unrolled so every offset is an immediate: a neighbour cell id is
    // the own id plus a constant plus at most three per-atom periodic-wrap corrections.
    cx = c % D.cellnum[0]; cy = (c / D.cellnum[0]) % D.cellnum[1]; cz = c / (D.cellnum[0] * D.cellnum[1]);
    const int nx = D.cellnum[0], nxy = D.cellnum[0] * D.cellnum[1];
    const int wxm = cx == 0 ? nx : 0, wxp = cx == nx - 1 ? -nx : 0;                                  // wrap when stepping -1 / +1 in x
    const int wym = cy == 0 ? nxy : 0, wyp = cy == D.cellnum[1] - 1 ? -nxy : 0;
    const int wzm = cz == 0 ? nxy * D.cellnum[2] : 0, wzp = cz == D.cellnum[2] - 1 ? -nxy * D.cellnum[2] : 0;
    const float cut2 = D.cut2_f32;
    const float csx = D.cellsize_f[0], csy = D.cellsize_f[1], csz = D.cellsize_f[2];
#pragma unroll
    for (int k = 0; k < 27; k++) {
        constexpr int KX[27] = {0, 1, 0, 0, 0, 1, 1, 1, -1, 0, 0, 0, -1, -1, -1, 0, 0, 1, -1, 1, -1, -1, -1, 1, 1, 1, -1};   // DMD.c:118-120
        constexpr int KY[27] = {0, 0, 1, 0, 1, 0, 1, 1, 0, -1, 0, -1, 0, -1, -1, -1, 1, 0, 0, -1, 1, -1, 1, -1, 1, -1, 1};
        constexpr int KZ[27] = {0, 0, 0, 1, 1, 1, 0, 1, 0, 0, -1, -1, -1, 0, -1, 1, -1, -1, 1, 0, 0, 1, -1, -1, -1, 1, 1};
        const int dx = KX[k], dy = KY[k], dz = KZ[k];
        int cell = c + dx + dy * nx + dz * nxy + (dx < 0 ? wxm : dx > 0 ? wxp : 0) + (dy < 0 ? wym : dy > 0 ? wyp : 0) +
                   (dz < 0 ? wzm : dz > 0 ? wzp : 0);
        // my position as seen from the neighbour cell's corner
        float mx = dx == 0 ? me.x : (dx > 0 ? me.x - csx : me.x + csx);
        float my = dy == 0 ? me.y : (dy > 0 ? me.y - csy : me.y + csy);
        float mz = dz == 0 ? me.z : (dz > 0 ? me.z - csz : me.z + csz);
        int p1 = D.cell_start[cell + 1];
        for (int p = D.cell_start[cell]; p < p1; p++) {
            if (p == p_self) continue;
            float4 q = __ldg(&D.sloc[p]);
            if (need_j) {
                int j = __float_as_int(q.w);
                if (has_cm && (d_cm_flags(D, i, j) & SDMD_BOND_CONNECT)) continue;
                if (hb_i == j) continue;
                if ((h0 == j) || (h0 != 0 && (h1 == j || (h1 != 0 && (h2 == j || (h2 != 0 && h3 == j)))))) continue;
            }
            npair++;
            float rx = mx - q.x, ry = my - q.y, rz = mz - q.z;
            float r2 = fmaf(rz, rz, fmaf(ry, ry, rx * rx));
            if (r2 <= cut2) {
                if (ns < cap) out[ns] = (unsigned)p | ((unsigned)k << 27);
                ns++;
            }
        }
    }
    D.surv_n[p_self] = (ns <= cap) ? ns : -1;
    D.surv_np[p_self] = npair;
}

// Stage 2 from the filter's list.  With DUP, entries whose partner's slot points back at this atom are tagged for
// the in-order CheckDuplication test; such a partner is always in the list (it saw this atom inside the scan
// sphere from its side, and the list is a superset of the sphere).
template <bool DUP>
__device__ inline void d_collision_time_from_list(const DevP &D, const SlotSrc &S, Self &me, Slot &cur, int p_self, int dup_limit) {
    int ns = D.surv_n[p_self];
    const unsigned int *in = D.surv + (size_t)p_self * D.surv_cap;
    int i = me.i;
    bool has_cm = me.cm1 > me.cm0;
    int type_i = me.type;
    int hb_i = 0, h0 = 0, h1 = 0, h2 = 0, h3 = 0;
    if (DUP) {
        hb_i = D.hb_bond[i];
        if (D.hydrogen_bond) { h0 = D.hbn_partner[5 * i + 1]; h1 = D.hbn_partner[5 * i + 2]; h2 = D.hbn_partner[5 * i + 3]; h3 = D.hbn_partner[5 * i + 4]; }
    }
#if K2_DIRECT
    if (!DUP) {
        // round 0 has no tagged entries: evaluate straight from the filter's list (no staging through local arrays),
        // the next entry's record in flight while the current one is evaluated
        const bool seam = me.cx == 0 || me.cy == 0 || me.cz == 0 || me.cx == D.cellnum[0] - 1 || me.cy == D.cellnum[1] - 1 ||
                          me.cz == D.cellnum[2] - 1;
        if (ns > 0) {
            unsigned v = K2_LD_ROW(in);
            const double2 *rp = reinterpret_cast<const double2 *>(&D.srec[v & 0x7ffffffu]);
            double2 r0 = __ldg(rp), r1 = __ldg(rp + 1), r2v = __ldg(rp + 2);
            int2 ta = __ldg(reinterpret_cast<const int2 *>(rp + 3));
            for (int s = 0; s < ns; s++) {
                unsigned vc = v;
                double2 c0 = r0, c1 = r1, c2 = r2v;
                int2 tc = ta;
#if K2_DIRECT == 2
                if (s + 1 < ns) {
                    v = K2_LD_ROW(in + s + 1);
                    rp = reinterpret_cast<const double2 *>(&D.srec[v & 0x7ffffffu]);
                    r0 = __ldg(rp); r1 = __ldg(rp + 1); r2v = __ldg(rp + 2);
                    ta = __ldg(reinterpret_cast<const int2 *>(rp + 3));
                }
#endif
                int j = tc.y;
                double shx = 0, shy = 0, shz = 0;
                if (seam) {
                    int ncx, ncy, ncz;
                    d_neighbor_cell(D, me.cx, me.cy, me.cz, (int)(vc >> 27), &ncx, &ncy, &ncz);
                    shx = d_shift1(me.cx, ncx, D.cellnum[0], d_box(D, 0));
                    shy = d_shift1(me.cy, ncy, D.cellnum[1], d_box(D, 1));
                    shz = d_shift1(me.cz, ncz, D.cellnum[2], d_box(D, 2));
                }
                int cm = has_cm ? d_cm_flags(D, i, j) : 0;
                double t, pot;
                int sub;
                if (d_eval_pair(D, me, j, tc.x, c0.x, c0.y, c1.x, c1.y, c2.x, c2.y, shx, shy, shz, cm, type_i, &t, &pot, &sub))
                    d_job_assign(D, cur, j, t, pot, sub, SDMD_EV_COLI);
#if K2_DIRECT != 2
                if (s + 1 < ns) {
                    v = K2_LD_ROW(in + s + 1);
                    rp = reinterpret_cast<const double2 *>(&D.srec[v & 0x7ffffffu]);
                    r0 = __ldg(rp); r1 = __ldg(rp + 1); r2v = __ldg(rp + 2);
                    ta = __ldg(reinterpret_cast<const int2 *>(rp + 3));
                }
#endif
            }
        }
        me.npair += D.surv_np[p_self];
        return;
    }
#endif
    int surv_p[R0_MAXSURV];
    unsigned char surv_k[R0_MAXSURV];
    int ndup = 0;
    unsigned wl[R0_SURV_CAP];
    if (D.rows_mode) {   // the row filter (sdmd_k2_rows.cu) writes survivors in row order: bring them into visiting order (ascending word)
        for (int s = 0; s < ns; s++) {
            unsigned v = in[s];
            int k = s;
            while (k > 0 && wl[k - 1] > v) { wl[k] = wl[k - 1]; k--; }
            wl[k] = v;
        }
        in = wl;
    }
    for (int s0 = 0; s0 < ns; s0 += R0_MAXSURV) {
        int m = ns - s0 < R0_MAXSURV ? ns - s0 : R0_MAXSURV;
        for (int s = 0; s < m; s++) {
            unsigned v = in[s0 + s];
            int p = (int)(v & 0x7ffffffu);
            unsigned char tag = (unsigned char)(v >> 27);
            if (DUP) {
                int j = D.sorted_atom[p];
                if (j < dup_limit && S.partner[j] == i && S.type[j] == SDMD_EV_COLI && S.pcounter[j] == me.counter) { tag |= R0_DUPFLAG; ndup++; }
            }
            surv_p[s] = p; surv_k[s] = tag;
        }
        d_r0_flush(D, S, me, cur, surv_p, surv_k, m, has_cm, type_i, hb_i, h0, h1, h2, h3);
    }
    // the filter counted every candidate that passes the skip filters; d_r0_flush re-counts tagged entries that were
    // not copied, so take the tagged ones out once
    me.npair += D.surv_np[p_self] - ndup;
}

// BondTime, TimePrediction.c:186-295
__device__ inline void d_bond_time(const DevP &D, const SlotSrc &S, Self &me, Slot &cur, int dup_limit) {
    int i = me.i;
    for (int k = D.bond_off[i]; k < D.bond_off[i + 1]; k++) {
        int j = D.bond_j[k];
        if (j == 0) break;
        if (d_check_dup(D, S, cur, i, me.counter, j, SDMD_EV_BOND, dup_limit)) continue;
        double long2 = D.bond_dmax2[k], short2 = D.bond_dmin2[k];
        Pair p = d_pair_atoms(D, i, j);
        me.npair++;
        if (p.r2 - long2 > SD_ZERO) {  // :240-253
            atomicAdd((unsigned long long *)&D.ctl->error_prints, 1ULL);
            if (p.b >= 0) {
                d_job_assign(D, cur, j, 1e-16 * i * j, 0, SDMD_SUB_WELLBOUNC, SDMD_EV_BOND);
                continue;
            }
        } else if (short2 - p.r2 > SD_ZERO) {  // :255-269
            atomicAdd((unsigned long long *)&D.ctl->error_prints, 1ULL);
            if (p.b < 0) {
                d_job_assign(D, cur, j, 1e-16 * i * j, 0, SDMD_SUB_CORECOLLI, SDMD_EV_BOND);
                continue;
            }
        }
        int s, sub;
        double d2;
        if (p.b < 0 && d_disc(p, short2) > 0) { s = -1; d2 = short2; sub = SDMD_SUB_CORECOLLI; }
        else { s = 1; d2 = long2; sub = SDMD_SUB_WELLBOUNC; }
        double t = d_time(p, s, d2);
        if (t < 0) { atomicAdd((unsigned long long *)&D.ctl->error_prints, 1ULL); t = SD_INF; }
        d_job_assign(D, cur, j, t, 0, sub, SDMD_EV_BOND);
    }
}

// CheckPairMatch, TimePrediction.c:1119-1138
__device__ inline int d_check_pair_match(const DevP &D, int i, int j) {
    int cls = d_hb_model(D, i, j);
    if (cls == -2) { D.ctl->error = SDMD_ERR_TABLE; return -1; }
    if (cls < 0) return -1;
    int ti = D.type[i], tj = D.type[j];
    int a = D.hb_index[(cls * SD_NT + ti) * SD_NT + tj];
    if (a < 0 || D.tables[a].nstep == 0) return -1;
    int ni = d_type_change(D.type[D.hb_neighbor[i]], 0);
    int nj = d_type_change(D.type[D.hb_neighbor[j]], 0);
    int b = D.hb_index[(cls * SD_NT + ti) * SD_NT + nj];
    int c = D.hb_index[(cls * SD_NT + tj) * SD_NT + ni];
    if (b < 0 || D.tables[b].nstep == 0 || c < 0 || D.tables[c].nstep == 0) return -1;
    return 0;
}

// HBTime, TimePrediction.c:298-483
__device__ inline void d_hb_time(const DevP &D, const SlotSrc &S, Self &me, Slot &cur, int dup_limit) {
    int i = me.i;
    if (D.hb_itype[i] == SDMD_HB_SKIP) { D.hb_itype[i] = SDMD_HB_NOEVENT; return; }
    D.hb_itype[i] = SDMD_HB_NOEVENT;
    int bonded = D.hb_bond[i];
    int role = D.hb_role[i];
    if (bonded > 0) {
        int j = bonded;
        if (d_check_dup(D, S, cur, i, me.counter, j, SDMD_EV_HB, dup_limit)) return;
        Pair p = d_pair_atoms(D, i, j);
        me.npair++;
        FindRes f = d_find_pair(D, i, j, FP_HB, p.b, p.r2, 0, 0);
        int s, sub = SDMD_SUB_INVALID;
        double d2, pot;
        if (p.b < 0 && d_disc(p, f.lower) > 0) {
            s = -1; d2 = f.lower; pot = f.lowPot;
            sub = (f.cs == 0) ? SDMD_SUB_CORECOLLI : SDMD_SUB_TBDCAPTR;
        } else {
            s = 1; d2 = f.upper; pot = f.upPot;
            if (D.method_ding) sub = (f.cs == 1) ? SDMD_SUB_TBDHBBRIK : SDMD_SUB_TBDESCAP;
        }
        double t = d_time(p, s, d2);
        if (t < 0) { atomicAdd((unsigned long long *)&D.ctl->error_prints, 1ULL); t = SD_INF; }
        d_job_assign(D, cur, j, t, pot, sub, SDMD_EV_HB);
    } else if ((role == 'A' || role == 'D' || role == 'S') && bonded == 0) {
        for (int k = 0; k < 27; k++) {
            int ncx, ncy, ncz;
            int cell = d_neighbor_cell(D, me.cx, me.cy, me.cz, k, &ncx, &ncy, &ncz);
            for (int j = D.celllist[D.n + cell + 1]; j; j = D.celllist[j]) {
                if (j == i) continue;
                int rj = D.hb_role[j];
                bool ad = (role == 'A' && rj == 'D') || (role == 'D' && rj == 'A');
                bool sd = (role == 'S' && D.type[j] == 2) || (rj == 'S' && D.type[i] == 2);
                bool ss = (role == 'S' && rj == 'S');
                if (!(ad || sd || ss)) continue;
                if (D.hb_bond[j] != 0) continue;
                int d = D.aa_num[i] - D.aa_num[j];
                if (d < 0) d = -d;
                if (!((D.protein[i] == D.protein[j] && d >= 4) || D.protein[i] != D.protein[j])) continue;
                if (d_check_dup(D, S, cur, i, me.counter, j, SDMD_EV_HB, dup_limit)) continue;
                Pair p = d_pair_atoms(D, i, j);
                me.npair++;
                if (!(p.r2 <= D.cut2)) continue;
                if (d_check_pair_match(D, i, j) < 0) continue;
                FindRes f = d_find_pair(D, i, j, FP_HB, p.b, p.r2, 0, 0);
                if (!(p.b < 0 && f.cs == 1 && d_disc(p, f.lower) > 0)) continue;
                double t = d_time(p, -1, f.lower);
                if (t < 0) { atomicAdd((unsigned long long *)&D.ctl->error_prints, 1ULL); t = SD_INF; }
                d_job_assign(D, cur, j, t, f.lowPot, SDMD_SUB_TBDHBFORM, SDMD_EV_HB);
            }
        }
    }
}

// HBNeighborTime, TimePrediction.c:486-579
__device__ inline void d_hbn_time(const DevP &D, const SlotSrc &S, Self &me, Slot &cur, int dup_limit) {
    int i = me.i;
    int flag = 0;
    int ns = D.hbn_status[i];
    for (int n = 1; n <= ns; n++) {
        int j = D.hbn_partner[5 * i + n];
        if (flag == 0) D.hbn_pnum[i] = n;
        if (d_check_dup(D, S, cur, i, me.counter, j, SDMD_EV_HBNE, dup_limit)) { D.hbn_pnum[i] = n; flag = 1; continue; }
        Pair p = d_pair_atoms(D, i, j);
        me.npair++;
        int host, nbr, third;
        int r = D.hb_role[i];
        int bi = D.hb_bond[i];
        if ((r == 'A' || r == 'D' || r == 'S') && (bi && D.hb_neighbor[bi] == j)) { host = i; nbr = j; third = bi; }
        else { host = j; nbr = i; third = D.hb_bond[j]; }
        FindRes f = d_find_pair(D, nbr, host, FP_NEIGHBOR, p.b, p.r2, third, 0);
        int s, sub = SDMD_SUB_INVALID;
        double d2, pot;
        if (p.b < 0 && d_disc(p, f.lower) > 0) {
            s = -1; d2 = f.lower; pot = f.lowPot;
            sub = (f.cs == 0) ? SDMD_SUB_CORECOLLI : SDMD_SUB_TBDCAPTR;
        } else {
            s = 1; d2 = f.upper; pot = f.upPot;
            if (D.method_ding) sub = (f.cs == 1) ? SDMD_SUB_WELLBOUNC : SDMD_SUB_TBDESCAP;
        }
        double t = d_time(p, s, d2);
        if (t < 0) { atomicAdd((unsigned long long *)&D.ctl->error_prints, 1ULL); t = SD_INF; }
        if (d_job_assign(D, cur, j, t, pot, sub, SDMD_EV_HBNE)) { D.hbn_pnum[i] = n; flag = 1; }
    }
}

// PBCandCrossCellTime, TimePrediction.c:598-632
__device__ __forceinline__ void d_pcc_time(const DevP &D, const Self &me, Slot &cur) {
    double best = SD_INF;
    double pos[3] = {me.x, me.y, me.z}, spd[3] = {me.vx, me.vy, me.vz};
    int ci[3] = {me.cx, me.cy, me.cz};
#pragma unroll
    for (int k = 0; k < 3; k++) {
        double lo = ci[k] * D.cellsize[k];
        double hi = (ci[k] + 1) * D.cellsize[k];
        if (spd[k] > 0) {
            double t = (hi - pos[k] + SD_ZERO) / spd[k];
            if (best > t && t > 0) best = t;
        } else if (spd[k] < 0) {
            double t = (lo - pos[k] - SD_ZERO) / spd[k];
            if (best > t && t > 0) best = t;
        }
    }
    d_job_assign(D, cur, 0, best, 0, SDMD_SUB_INVALID, SDMD_EV_PCC);
}

// ThermostatTime, TimePrediction.c:582-595 with a pre-drawn standard normal
__device__ __forceinline__ void d_thermostat_time(const DevP &D, Slot &cur, double g) {
    double t = D.thermoF * ((g < 0) ? (-g) : (g));
    d_job_assign(D, cur, 0, t, 0, SDMD_SUB_INVALID, SDMD_EV_THER);
}

// WallTime, TimePrediction.c:635-737 (smooth wall, wallDyn "no"): the wall is a pseudo-atom at rest on the mid-plane /
// axis / centre (wall_pos); FindPair sees the squared distance from the wall SURFACE, the quadratic is solved in the
// distance from the centre.  Visited last in Predict (SGThread.c:183-185), no partner.
__device__ inline void d_wall_time(const DevP &D, const Self &me, Slot &cur) {
    double px = me.x, py = me.y, pz = me.z, sx = me.vx, sy = me.vy, sz = me.vz;
    if (D.wall_kind == SDMD_WALL_PARALLEL) { px = 0; pz = 0; sx = 0; sz = 0; }      // :654-656
    else if (D.wall_kind == SDMD_WALL_CYLINDER) { px = 0; sx = 0; }                  // :657-660
    Pair p;
    p.rx = px - D.wall_pos[0]; p.ry = py - D.wall_pos[1]; p.rz = pz - D.wall_pos[2];
    double vx = sx - 0.0, vy = sy - 0.0, vz = sz - 0.0;
    p.r2 = p.rx * p.rx + p.ry * p.ry + p.rz * p.rz;
    p.b = p.rx * vx + p.ry * vy + p.rz * vz;
    p.v2 = vx * vx + vy * vy + vz * vz;
    int idx = __ldg(&D.coll_index[me.type * SD_NT + D.wall_type]);
    if (idx < 0) { D.ctl->error = SDMD_ERR_TABLE; return; }
    double by;
    FindRes f;
    for (;;) {                                                                        // ReCal, :641
        by = d_box(D, 1);
        double dist2 = by * by * 0.25 + p.r2 - by * sqrt(p.r2);                       // :667
        f = d_find_pair_table(D, &D.tables[idx], FP_COLLISION, -1 * p.b, dist2, 0, 0);
        if (!(f.lowPot == 0 && f.upPot == 0)) break;
        // outside of the wall: under a moving wall the box is pushed back out by one step and the prediction repeated (:675-697)
        if (!D.wd_active) { D.ctl->error = SDMD_ERR_STATE; return; }                  // exit() :699-701
        Ctl *c = D.ctl;
        c->wd_touch++;
        if (D.wall_kind == SDMD_WALL_PARALLEL) c->dbox[1] += D.wd_step;
        else if (D.wall_kind == SDMD_WALL_CYLINDER) { c->dbox[1] += D.wd_step; c->dbox[2] += D.wd_step; }
        else { c->dbox[0] += D.wd_step; c->dbox[1] += D.wd_step; c->dbox[2] += D.wd_step; }
        if (c->dbox[1] > D.wd_orig_y) { c->error = SDMD_ERR_STATE; return; }          // "cannot be expanded beyond the original": exit() :691-694
    }
    int s, sub;
    double d2, pot;
    if (p.b < 0 && d_disc(p, f.lower) > 0) {                                          // towards the centre :704-714
        if (f.cs == 1) return;
        sub = SDMD_SUB_TBDCAPTR;
        s = -1;
        d2 = by * by * 0.25 + f.upper - by * sqrt(f.upper);
        pot = f.upPot;
    } else {                                                                          // towards the wall :716-726
        sub = (f.cs == 0) ? SDMD_SUB_WELLBOUNC : SDMD_SUB_TBDESCAP;
        s = 1;
        d2 = by * by * 0.25 + f.lower - by * sqrt(f.lower);
        pot = f.lowPot;
    }
    double t = d_time(p, s, d2);
    if (t < 0) t = SD_INF;                                                            // :729-732
    d_job_assign(D, cur, 0, t, pot, sub, SDMD_EV_WALL);
}

// FindPotWellWidth, ToolFunctions.c:954-970: squared range of the pair potential of (type_i, type_j)
__device__ __forceinline__ double d_pot_well_width(const DevP &D, int type_i, int type_j) {
    int idx = __ldg(&D.coll_index[type_i * SD_NT + type_j]);
    if (idx < 0) return 0;
    const sdmd_step_table *t = &D.tables[idx];
    if (t->dmax2 > 0) return t->dmax2;
    return t->nstep > 0 ? t->d2[t->nstep - 1] : 0;
}

// EnterTunnel, ToolFunctions.c:921-952: will the atom be inside a tunnel mouth's cross-section after `time`?
__device__ inline bool d_enter_tunnel(const DevP &D, const Self &me, double time) {
    for (int n = 0; n < D.n_tunnel; n++) {
        double radius2 = D.sb->tun_diam[n] * 0.5;
        radius2 *= radius2;
        double shift2 = d_pot_well_width(D, D.tunnel_type, me.type);
        double py = me.y, pz = me.z;
        py += me.vy * time;
        pz += me.vz * time;
        double r1 = D.sb->tun_pos[n][1] - D.sb->tun_pos[n][1], r2 = py - D.sb->tun_pos[n][2], r3 = pz - D.sb->tun_pos[n][3];
        double r_2 = r1 * r1 + r2 * r2 + r3 * r3;
        double distance2 = radius2 + shift2 - 2 * sqrt(radius2 * shift2);
        if (r_2 <= distance2) return true;
    }
    return false;
}

// TunnelTime, TimePrediction.c:832-930: inside a tunnel the wall distance is taken from the tunnel's axis (y, z); partner
// code -10000 n
__device__ inline void d_tunnel_time(const DevP &D, const Self &me, Slot &cur) {
    int idx = __ldg(&D.coll_index[me.type * SD_NT + D.tunnel_type]);
    if (idx < 0) { D.ctl->error = SDMD_ERR_TABLE; return; }
    const double shift2 = D.tables[idx].dmin2;
    const double extraRange = sqrt(d_pot_well_width(D, D.tunnel_type, me.type));
    for (int n = 0; n < D.n_tunnel; n++) {
        double radius2 = D.sb->tun_diam[n] / 2;
        radius2 *= radius2;
        Pair p;
        p.rx = D.sb->tun_pos[n][1] - D.sb->tun_pos[n][1]; p.ry = me.y - D.sb->tun_pos[n][2]; p.rz = me.z - D.sb->tun_pos[n][3];
        double vx = 0 - 0.0, vy = me.vy - 0.0, vz = me.vz - 0.0;
        p.r2 = p.rx * p.rx + p.ry * p.ry + p.rz * p.rz;
        p.b = p.rx * vx + p.ry * vy + p.rz * vz;
        p.v2 = vx * vx + vy * vy + vz * vz;
        if (p.r2 > radius2 + shift2 - 2 * sqrt(radius2 * shift2) + SD_ZERO) continue;   // outside the tunnel
        double dist2 = radius2 + p.r2 - 2 * sqrt(p.r2 * radius2);
        FindRes f = d_find_pair_table(D, &D.tables[idx], FP_COLLISION, -1 * p.b, dist2, 0, 0);
        if (f.lowPot == 0 && f.upPot == 0) { D.ctl->error = SDMD_ERR_STATE; return; }    // exit(), :880-884
        int s, sub;
        double d2, pot;
        if (p.b < 0 && d_disc(p, f.lower) > 0) {
            if (f.cs == 1) continue;
            sub = SDMD_SUB_TBDCAPTR; s = -1;
            d2 = radius2 + f.upper - 2 * sqrt(f.upper * radius2);
            pot = f.upPot;
        } else {
            sub = (f.cs == 0) ? SDMD_SUB_WELLBOUNC : SDMD_SUB_TBDESCAP; s = 1;
            d2 = radius2 + f.lower - 2 * sqrt(f.lower * radius2);
            pot = f.lowPot;
        }
        double t = d_time(p, s, d2);
        double futurePos = me.x + me.vx * t;                                              // :912-917
        if (futurePos < D.sb->tun_pos[n][0] - extraRange || futurePos > D.sb->tun_pos[n][1] + extraRange) continue;
        if (t < 0) t = SD_INF;
        if (d_job_assign(D, cur, 0, t, pot, sub, SDMD_EV_TUNL)) cur.partner = -10000 * n;
    }
}

// ObstTime, TimePrediction.c:740-829 (obstruction walls): one-dimensional distance from the plane,
// partner code -dim - 10 n
__device__ inline void d_obst_time(const DevP &D, const Self &me, Slot &cur) {
    const double pos[3] = {me.x, me.y, me.z}, spd[3] = {me.vx, me.vy, me.vz};
    for (int n = 0; n < D.n_obst; n++) {
#pragma unroll 1
        for (int dim = 0; dim < 3; dim++) {
            if (D.sb->obst_pos[n][dim] < 0) continue;
            Pair p;
            double r[3], v[3];
#pragma unroll
            for (int k = 0; k < 3; k++) {
                double pi = (k == dim) ? pos[k] : 0, pj = (k == dim) ? D.sb->obst_pos[n][k] : 0, si = (k == dim) ? spd[k] : 0;
                r[k] = pi - pj; v[k] = si - 0.0;
            }
            p.rx = r[0]; p.ry = r[1]; p.rz = r[2];
            p.r2 = r[0] * r[0] + r[1] * r[1] + r[2] * r[2];
            p.b = r[0] * v[0] + r[1] * v[1] + r[2] * v[2];
            p.v2 = v[0] * v[0] + v[1] * v[1] + v[2] * v[2];
            int idx = __ldg(&D.coll_index[me.type * SD_NT + D.sb->obst_type[n]]);
            if (idx < 0) { D.ctl->error = SDMD_ERR_TABLE; return; }
            FindRes f = d_find_pair_table(D, &D.tables[idx], FP_COLLISION, p.b, p.r2, 0, 0);
            if (f.lowPot == 0 && f.upPot == 0) continue;                     // :783-787
            int s, sub;
            double d2, pot;
            if (p.b < 0) {                                                    // approaching :789-798
                sub = (f.cs == 0) ? SDMD_SUB_CORECOLLI : SDMD_SUB_TBDCAPTR;
                s = -1; d2 = f.lower; pot = f.lowPot;
            } else {                                                          // leaving :800-810
                if (f.cs == 1) continue;
                sub = SDMD_SUB_TBDESCAP;
                s = 1; d2 = f.upper; pot = f.upPot;
            }
            double t = d_time(p, s, d2);
            if (t < 0) { D.ctl->error = SDMD_ERR_STATE; return; }             // exit(), :813-816
            // at the next collision the atom would already be in the hole: the prediction is void, :818-824
            if (!(D.sb->obst_hole[n] && d_enter_tunnel(D, me, t)) && d_job_assign(D, cur, 0, t, pot, sub, SDMD_EV_OBST))
                cur.partner = -1 * (dim + 1) - 10 * n;
        }
    }
}

// ChargeTime, TimePrediction.c:933-1017: concentric potential steps of width `gap` around each charge source, partner code -100 n
__device__ inline void d_charge_time(const DevP &D, const Self &me, Slot &cur) {
    if (!D.charge) return;
    const double charge = D.charge[me.i];
    if (!charge) return;
    int flag = 0;
    for (int cn = 0; cn < D.n_charge; cn++) {
        if (me.x > D.sb->chg_pos[cn][0] - SD_ZERO) continue;
        if (D.sb->chg_pos[cn][0] < d_box(D, 0)) flag = 1;                                   // :948-956
        else if (!(flag == 0 && D.sb->chg_pos[cn][0] > d_box(D, 0))) break;
        Pair p;
        p.rx = me.x - D.sb->chg_pos[cn][0]; p.ry = me.y - D.sb->chg_pos[cn][1]; p.rz = me.z - D.sb->chg_pos[cn][2];
        double vx = me.vx - D.sb->chg_vel[cn][0], vy = me.vy - D.sb->chg_vel[cn][1], vz = me.vz - D.sb->chg_vel[cn][2];
        p.r2 = p.rx * p.rx + p.ry * p.ry + p.rz * p.rz;
        p.b = p.rx * vx + p.ry * vy + p.rz * vz;
        p.v2 = vx * vx + vy * vy + vz * vz;
        int gapNum = (int)(sqrt(p.r2) / D.sb->chg_gap[cn]);
        const double gap2 = D.sb->chg_gap[cn] * D.sb->chg_gap[cn];
        const double distance2 = (gapNum * gapNum) * gap2;
        if (p.b < 0 && p.r2 < distance2 + SD_ZERO) gapNum--;
        else if (p.b > 0 && p.r2 > ((gapNum + 1) * (gapNum + 1)) * gap2 - SD_ZERO) gapNum++;
        const double upPot = D.sb->chg_grad[cn] * charge, long2 = ((gapNum + 1) * (gapNum + 1)) * gap2;
        const double lowPot = -1 * D.sb->chg_grad[cn] * charge, short2 = (gapNum * gapNum) * gap2;
        int s, sub;
        double d2, pot;
        if (p.b < 0 && d_disc(p, short2) > 0) { s = -1; d2 = short2; pot = lowPot; sub = SDMD_SUB_TBDCAPTR; }
        else { s = 1; d2 = long2; pot = upPot; sub = SDMD_SUB_TBDESCAP; }
        double t = d_time(p, s, d2);
        if (t < 0) t = SD_INF;
        if (d_job_assign(D, cur, 0, t, pot, sub, SDMD_EV_CHRG)) cur.partner = -100 * cn;
    }
}

// SphObstTime, TimePrediction.c:1020-1071: hard sphere at rest, core collision only, partner code -1000 n
__device__ inline void d_sphobst_time(const DevP &D, const Self &me, Slot &cur) {
    for (int n = 0; n < D.n_sphobst; n++) {
        Pair p;
        p.rx = me.x - D.sb->sph_pos[n][0]; p.ry = me.y - D.sb->sph_pos[n][1]; p.rz = me.z - D.sb->sph_pos[n][2];
        double vx = me.vx - 0.0, vy = me.vy - 0.0, vz = me.vz - 0.0;
        p.r2 = p.rx * p.rx + p.ry * p.ry + p.rz * p.rz;
        p.b = p.rx * vx + p.ry * vy + p.rz * vz;
        p.v2 = vx * vx + vy * vy + vz * vz;
        double short2 = D.sb->sph_r[n] * D.sb->sph_r[n];
        if (!(p.b < 0 && d_disc(p, short2) > 0)) continue;
        double t = d_time(p, -1, short2);
        if (t < 0) { D.ctl->error = SDMD_ERR_STATE; return; }
        if (d_job_assign(D, cur, 0, t, 0, SDMD_SUB_CORECOLLI, SDMD_EV_SPHO)) cur.partner = -1000 * n;
    }
}

__device__ __forceinline__ Self d_load_self(const DevP &D, int i) {
    Self me;
    me.i = i;
    me.x = d_rx(D, i); me.y = d_ry(D, i); me.z = d_rz(D, i);
    me.vx = D.vx[i]; me.vy = D.vy[i]; me.vz = D.vz[i];
    me.cx = D.cell_idx[3 * i]; me.cy = D.cell_idx[3 * i + 1]; me.cz = D.cell_idx[3 * i + 2];
    me.counter = D.ev_counter[i];
    me.npair = 0;
    me.type = D.type[i];
    me.cm0 = D.cm_off[i]; me.cm1 = D.cm_off[i + 1];
    return me;
}

// the same from the cell-sorted record at position p: one coalesced 64-byte read instead of ten gathers by atom id
__device__ __forceinline__ Self d_load_self_sorted(const DevP &D, int p) {
    const double2 *rp = reinterpret_cast<const double2 *>(&D.srec[p]);
    double2 r0 = __ldg(rp), r1 = __ldg(rp + 1), r2 = __ldg(rp + 2);
    int4 q = __ldg(reinterpret_cast<const int4 *>(rp + 3));
    Self me;
    me.i = q.y;
    me.x = r0.x; me.y = r0.y; me.z = r1.x;
    me.vx = r1.y; me.vy = r2.x; me.vz = r2.y;
    me.cx = (int)((unsigned)q.z & 0xffffu); me.cy = (int)((unsigned)q.z >> 16); me.cz = q.w;
    me.counter = D.ev_counter[q.y];
    me.npair = 0;
    me.type = q.x;
    me.cm0 = D.cm_off[q.y]; me.cm1 = D.cm_off[q.y + 1];
    return me;
}

// ResetTarget (ToolFunctions.c:504-514) leaves potential / subEventType / partnerCounter untouched
__device__ __forceinline__ Slot d_reset_slot(double pot, int sub, i64 pcounter) {
    Slot s;
    s.time = SD_INF; s.type = SDMD_EV_INVD; s.partner = SDMD_INVALID; s.m2 = SD_INF; s.bp_time = SD_INF; s.sp_time = SD_INF; s.bp_j = 0;
    s.pot = pot; s.sub = sub; s.pcounter = pcounter;
    return s;
}

// one atom of Predict, SGThread.c:162-208, executed by a single thread in reference order.
// SORTED selects the cell-sorted tiles for the 27-cell scan (fresh link list only).
struct PairBest { double bp_time, sp_time; int bp_j; };
// LEAN: the caller guarantees a system without bonds, hydrogen bonds and single-body sources (gel_mode, !hydrogen_bond,
// !single_body — uniform per system), so those predictors are compiled out of the kernel instance.
template <bool SORTED, bool R0 = false, bool LEAN = false>
__device__ inline Slot d_predict_atom(const DevP &D, const SlotSrc &S, int i, Slot cur, double gauss, int dup_limit, int *npair,
                                      double *second = nullptr, PairBest *pb = nullptr, int surv_pos = -1) {
    if (!D.no_cons && D.is_cons[i]) { cur.time = SD_INF; return cur; }
    Self me = (SORTED && surv_pos >= 0) ? d_load_self_sorted(D, surv_pos) : d_load_self(D, i);
    if (!LEAN && !D.gel_mode) d_bond_time(D, S, me, cur, dup_limit);
    if (R0 && surv_pos >= 0 && D.surv_n[surv_pos] >= 0) d_collision_time_from_list<false>(D, S, me, cur, surv_pos, 0);
    else if (R0) d_collision_time_2stage<false>(D, S, me, cur, 0);
    else if (SORTED && surv_pos >= 0 && D.surv_n[surv_pos] >= 0) d_collision_time_from_list<true>(D, S, me, cur, surv_pos, dup_limit);
    else if (SORTED) d_collision_time_2stage<true>(D, S, me, cur, dup_limit);
    else d_collision_time(D, S, me, cur, dup_limit);
    if (pb) { pb->bp_time = cur.bp_time; pb->sp_time = cur.sp_time; pb->bp_j = cur.bp_j; }   // pair records in front of PCC
    d_pcc_time(D, me, cur);
    if (D.thermostat_on) d_thermostat_time(D, cur, gauss);
    if (D.hydrogen_bond) {
        d_hb_time(D, S, me, cur, dup_limit);
        d_hbn_time(D, S, me, cur, dup_limit);
    }
    if (D.single_body) {   // the single-body predictors come last, SGThread.c:183-207
        if (D.wall_kind) d_wall_time(D, me, cur);
        if (D.n_obst) d_obst_time(D, me, cur);
        if (D.n_tunnel) d_tunnel_time(D, me, cur);
        if (D.flow_mark == 3) d_charge_time(D, me, cur);
        if (D.n_sphobst) d_sphobst_time(D, me, cur);
    }
    *npair = me.npair;
    if (second) *second = cur.m2;
    return cur;
}

__device__ __forceinline__ SlotSrc d_live_slots(const DevP &D) {
    SlotSrc S;
    S.time = D.ev_time; S.pot = D.ev_pot; S.partner = D.ev_partner; S.pcounter = D.ev_pcounter; S.type = D.ev_type; S.sub = D.ev_sub;
    return S;
}
__device__ __forceinline__ void d_store_slot(const DevP &D, int i, const Slot &s) {
    D.ev_time[i] = s.time; D.ev_pot[i] = s.pot; D.ev_partner[i] = s.partner; D.ev_pcounter[i] = s.pcounter;
    D.ev_type[i] = (uint8_t)s.type; D.ev_sub[i] = (uint8_t)s.sub;
}
__device__ __forceinline__ double d_next_gauss(const DevP &D) {
    Ctl *c = D.ctl;
    if (c->g_pos >= c->g_filled) { c->error = SDMD_ERR_STATE; return 1.0; }
    double g = D.gauss[c->g_pos % D.g_cap];
    c->g_pos++;
    return g;
}

// Predict(renewList) for the event loop, SGThread.c:158-211: reset all listed slots first, then predict
// them in list order (the partner sees the target's fresh slot, never the reverse)
__device__ inline void d_predict_list_serial(const DevP &D, int a0, int a1) {
    int list[2] = {a0, a1};
    SlotSrc S = d_live_slots(D);
    for (int k = 0; k < 2; k++) {
        int i = list[k];
        if (i <= 0) continue;
        D.ev_type[i] = SDMD_EV_INVD; D.ev_time[i] = SD_INF; D.ev_partner[i] = SDMD_INVALID;
    }
    for (int k = 0; k < 2; k++) {
        int i = list[k];
        if (i <= 0) continue;
        Slot cur = d_reset_slot(D.ev_pot[i], D.ev_sub[i], D.ev_pcounter[i]);
        double g = 0;
        if (D.thermostat_on && !D.is_cons[i]) g = d_next_gauss(D);
        int np = 0;
        cur = d_predict_atom<false>(D, S, i, cur, g, 0x7fffffff, &np);
        d_store_slot(D, i, cur);
        D.ctl->pair_predictions += np;
    }
}

// ---------------------------------------------------------------------------------------------------
// UpdateCBTRun, CBT.c:56-83
__device__ inline void d_cbt_update_run(const DevP &D, int a) {
    int *node = D.cbt_node;
    int father, left, right, old;
    for (father = D.cbt_leaf[a] / 2; father > 0; father /= 2) {
        if (node[father] != a) break;
        left = node[father * 2]; right = node[father * 2 + 1];
        node[father] = (d_rt(D, left) < d_rt(D, right)) ? left : right;
    }
    for (; father > 0; father /= 2) {
        old = node[father];
        left = node[father * 2]; right = node[father * 2 + 1];
        int w = (d_rt(D, left) < d_rt(D, right)) ? left : right;
        node[father] = w;
        if (w == old) return;
    }
}

// ---------------------------------------------------------------------------------------------------
// event execution (Events.c), serial device versions

// HBBarrier, Events.c:1097-1119
__device__ inline double d_hb_barrier(const DevP &D, int a1, int a2) {
    if (D.hb_role[a1] == 'A' || D.hb_role[a1] == 'S') { int s = a1; a1 = a2; a2 = s; }
    int o1 = d_type_change(D.type[a1], 0), o2 = d_type_change(D.type[a2], 0);
    if (o1 == 2 && o2 == 26) return D.hb_barrier[0];
    if (o1 != 2 && o2 != 26) return D.hb_barrier[1];
    return D.hb_barrier[2];
}

// HBEnergyChange, Events.c:1027-1095 (hbType 0 "_nonHB", 1 "_HB"); `tmp` and `accumPoten` carry over
// between pairs exactly like the reference's aliased locals
__device__ inline double d_hb_energy_change(const DevP &D, int i, int skipAtom, int hbType) {
    double energyChange = 0, tmp = 0, accumPoten = 0;
    int cx = D.cell_idx[3 * i], cy = D.cell_idx[3 * i + 1], cz = D.cell_idx[3 * i + 2];
    double xi = d_rx(D, i), yi = d_ry(D, i), zi = d_rz(D, i);
    int skipNeighbor = D.hb_neighbor[skipAtom];
    for (int k = 0; k < 27; k++) {
        int ncx, ncy, ncz;
        int cell = d_neighbor_cell(D, cx, cy, cz, k, &ncx, &ncy, &ncz);
        double shx = d_shift1(cx, ncx, D.cellnum[0], d_box(D, 0));
        double shy = d_shift1(cy, ncy, D.cellnum[1], d_box(D, 1));
        double shz = d_shift1(cz, ncz, D.cellnum[2], d_box(D, 2));
        for (int j = D.celllist[D.n + cell + 1]; j; j = D.celllist[j]) {
            if (j == i || j == skipAtom) continue;
            int cm = d_cm_flags(D, i, j);
            if (cm & (SDMD_BOND_CONNECT | SDMD_CONSTRAINT_CONNECT)) continue;
            double tx = d_rx(D, j) + shx, ty = d_ry(D, j) + shy, tz = d_rz(D, j) + shz;
            double rx = xi - tx, ry = yi - ty, rz = zi - tz;
            double r2 = rx * rx + ry * ry + rz * rz;
            if (r2 <= D.cutfull2) {
                if (hbType == 0) {
                    FindRes f = d_find_pair(D, i, j, FP_COLLISION, tmp, r2, 0, -1);
                    tmp = f.upPot; accumPoten = f.accum;
                } else if (j != skipNeighbor) {
                    FindRes f = d_find_pair(D, i, j, FP_COLLISION, tmp, r2, 0, 1);
                    tmp = f.upPot; accumPoten = f.accum;
                }
                if (accumPoten == SD_INF) accumPoten = 0;
                energyChange += accumPoten;
            }
        }
    }
    return energyChange;
}

// position of b as seen from a under the cell-index minimum image, Events.c:815-823
__device__ inline void d_image_of(const DevP &D, int a, int b, double pos[3]) {
    pos[0] = d_rx(D, b); pos[1] = d_ry(D, b); pos[2] = d_rz(D, b);
#pragma unroll
    for (int k = 0; k < 3; k++) {
        int ca = D.cell_idx[3 * a + k], cb = D.cell_idx[3 * b + k];
        if (ca != cb) {
            if (ca == 0 && cb == D.cellnum[k] - 1) pos[k] -= d_box(D, k);
            else if (ca == D.cellnum[k] - 1 && cb == 0) pos[k] += d_box(D, k);
        }
    }
}

// InteractionEvent, Events.c:798-1024
__device__ inline void d_interaction_event(const DevP &D, int i, int j, int hbEvent) {
    double pi[3] = {d_rx(D, i), d_ry(D, i), d_rz(D, i)}, pj[3];
    d_image_of(D, i, j, pj);
    double rx = pi[0] - pj[0], ry = pi[1] - pj[1], rz = pi[2] - pj[2];
    double wx = D.vx[i] - D.vx[j], wy = D.vy[i] - D.vy[j], wz = D.vz[i] - D.vz[j];
    double d2 = rx * rx + ry * ry + rz * rz;
    double b = rx * wx + ry * wy + rz * wz;
    double mi = D.mass[i], mj = D.mass[j];
    double reduced = mi * mj / (mi + mj);
    double potential = D.ev_pot[i];
    int sub = D.ev_sub[i];
    int s_vel = 0;

    if (hbEvent && (sub == SDMD_SUB_TBDHBFORM || sub == SDMD_SUB_TBDHBBRIK)) {  // :832-943
        int ni = D.hb_neighbor[i], nj = D.hb_neighbor[j];
        if ((d_neighbor_connected(D, i, nj) || d_neighbor_connected(D, j, ni)) && sub == SDMD_SUB_TBDHBFORM) {
            D.hb_itype[i] = SDMD_HB_SKIP;
            return;
        }
        double pni[3], pnj[3];
        d_image_of(D, i, ni, pni);
        d_image_of(D, i, nj, pnj);
        double ax = pi[0] - pnj[0], ay = pi[1] - pnj[1], az = pi[2] - pnj[2];
        double bx = pj[0] - pni[0], by = pj[1] - pni[1], bz = pj[2] - pni[2];
        double rni2 = ax * ax + ay * ay + az * az;
        double rnj2 = bx * bx + by * by + bz * bz;
        double acc0 = d_hb_energy_change(D, i, j, 0);
        double acc1 = d_hb_energy_change(D, j, i, 0);
        double acc2 = d_hb_energy_change(D, i, j, 1);
        double acc3 = d_hb_energy_change(D, j, i, 1);
        double acc4 = d_find_pair(D, nj, i, FP_NEIGHBOR, 0.0, rni2, j, 0).accum;
        double acc5 = d_find_pair(D, ni, j, FP_NEIGHBOR, 0.0, rnj2, i, 0).accum;
        if (sub == SDMD_SUB_TBDHBFORM) {
            potential -= d_hb_barrier(D, i, j);
            potential += (acc2 - acc0 + acc3 - acc1 + acc4 + acc5);
        } else {
            potential += d_hb_barrier(D, i, j);
            potential += (acc0 - acc2 + acc1 - acc3 - acc4 - acc5);
        }
        if (potential < -0.5 * SD_INF && sub == SDMD_SUB_TBDHBBRIK) { D.hb_itype[i] = SDMD_HB_SKIP; return; }
        else if (potential > 0.5 * SD_INF && sub == SDMD_SUB_TBDHBFORM) { D.hb_itype[i] = SDMD_HB_SKIP; return; }
        if (b * b - 2 * d2 * potential / reduced > 0) {
            if (sub == SDMD_SUB_TBDHBFORM) { sub = SDMD_SUB_WELLCAPTR; D.hb_itype[i] = SDMD_HB_FORM; }
            else { sub = SDMD_SUB_WELLESCAP; D.hb_itype[i] = SDMD_HB_BREAK; }
        } else {
            if (sub == SDMD_SUB_TBDHBFORM) { D.hb_itype[i] = SDMD_HB_SKIP; return; }
            potential = 0;
            sub = SDMD_SUB_WELLBOUNC;
            D.hb_itype[i] = SDMD_HB_NOEVENT;
        }
        D.ev_sub[i] = (uint8_t)sub;
    }

    switch (sub) {  // :945-1003
        case SDMD_SUB_CORECOLLI: s_vel = 1; potential = 0; break;
        case SDMD_SUB_WELLBOUNC: s_vel = (b < 0) ? 1 : -1; potential = 0; break;
        case SDMD_SUB_WELLCAPTR: s_vel = -1; break;
        case SDMD_SUB_WELLESCAP: s_vel = (b < 0) ? -1 : 1; break;
        case SDMD_SUB_TBDCAPTR:
            if (b * b - 2 * d2 * potential / reduced <= 0) { potential = 0; s_vel = 1; }
            else s_vel = -1;
            break;
        case SDMD_SUB_TBDESCAP:
            if (b * b - 2 * d2 * potential / reduced <= 0) { potential = 0; s_vel = (b < 0) ? 1 : -1; }
            else s_vel = (b < 0) ? -1 : 1;
            break;
        default: D.ctl->error = SDMD_ERR_STATE; return;
    }
    double phi = (-1 * b + s_vel * sqrt((b * b - 2 * d2 * potential / reduced))) / ((mi + mj) * d2);  // :1005-1006
    if (phi != phi) { D.ctl->error = SDMD_ERR_STATE; return; }
    if (!D.is_cons[i]) {
        double f = mj * phi;
        D.vx[i] = D.vx[i] + f * rx; D.vy[i] = D.vy[i] + f * ry; D.vz[i] = D.vz[i] + f * rz;
    }
    if (!D.is_cons[j]) {
        double f = mi * phi;
        D.vx[j] = D.vx[j] - f * rx; D.vy[j] = D.vy[j] - f * ry; D.vz[j] = D.vz[j] - f * rz;
    }
}

// ListRefresh(target, list, 1, 4), ToolFunctions.c:572-597
__device__ inline void d_list_refresh(const DevP &D, int a, int target) {
    int *l = &D.hbn_partner[5 * a];
    int i = d_find_in_hbn(D, a, target);
    if (i) {
        int n = i + 1;
        while (n <= 4 && l[n] != 0) { l[i] = l[n]; i++; n++; }
        l[i] = 0;
    }
}

// HBEvent, Events.c:275-393
// HB form / break changes what cached candidate lists recorded about atoms i and j (type, bonded state): every list that
// can contain them lies within one cell of theirs
__device__ __forceinline__ void d_touch_cells(const DevP &D, int i, int j) {
    i64 t = ++D.ctl->cc_epoch;
    D.cell_epoch[(D.cell_idx[3 * i + 2] * D.cellnum[1] + D.cell_idx[3 * i + 1]) * D.cellnum[0] + D.cell_idx[3 * i]] = t;
    D.cell_epoch[(D.cell_idx[3 * j + 2] * D.cellnum[1] + D.cell_idx[3 * j + 1]) * D.cellnum[0] + D.cell_idx[3 * j]] = t;
}
// CheckAlphaHB, Events.c:1121-1128: backbone H and O of one chain, the O four residues in front of the H
__device__ __forceinline__ bool d_check_alpha_hb(const DevP &D, int i, int j) {
    if (!D.hb_name || D.protein[i] != D.protein[j]) return false;
    int ni = D.hb_name[i], nj = D.hb_name[j];
    return (ni == 1 && nj == 2 && D.aa_num[j] == D.aa_num[i] - 4) || (ni == 2 && nj == 1 && D.aa_num[i] == D.aa_num[j] - 4);
}
__device__ inline void d_hb_event(const DevP &D, int i, int j, int ni, int nj) {
    if (D.hb_itype[i] == SDMD_HB_FORM) {
        D.hb_bond[i] = j; D.hb_bond[j] = i;
        D.type[i] = d_type_change(D.type[i], 1);
        D.type[j] = d_type_change(D.type[j], 1);
        int t;
        t = ++D.hbn_status[i];  D.hbn_partner[5 * i + t] = D.hb_neighbor[j];
        t = ++D.hbn_status[j];  D.hbn_partner[5 * j + t] = D.hb_neighbor[i];
        t = ++D.hbn_status[ni]; D.hbn_partner[5 * ni + t] = j;
        t = ++D.hbn_status[nj]; D.hbn_partner[5 * nj + t] = i;
        if (d_check_alpha_hb(D, i, j)) D.ctl->alphaformed++;
        D.ctl->hbnumformed++;
        d_touch_cells(D, i, j);
    } else if (D.hb_itype[i] == SDMD_HB_BREAK) {
        D.hb_bond[i] = 0; D.hb_bond[j] = 0;
        D.type[i] = d_type_change(D.type[i], 0);
        D.type[j] = d_type_change(D.type[j], 0);
        --D.hbn_status[i];  d_list_refresh(D, i, D.hb_neighbor[j]);
        --D.hbn_status[j];  d_list_refresh(D, j, D.hb_neighbor[i]);
        --D.hbn_status[ni]; d_list_refresh(D, ni, j);
        --D.hbn_status[nj]; d_list_refresh(D, nj, i);
        if (d_check_alpha_hb(D, i, j)) D.ctl->alphaformed--;
        D.ctl->hbnumformed--;
        if (D.ctl->hbnumformed < D.ctl->alphaformed) atomicAdd((unsigned long long *)&D.ctl->warnings, 1ULL);   // Events.c:386-389
        d_touch_cells(D, i, j);
    }
}

// WallEvent, Events.c:423-499 (smooth wall): radial velocity kick of the atom alone
__device__ inline void d_wall_event(const DevP &D, int i) {
    double px = D.x[i], py = D.y[i], pz = D.z[i], sx = D.vx[i], sy = D.vy[i], sz = D.vz[i];
    double tvx = sx, tvy = sy, tvz = sz;
    if (D.wall_kind == SDMD_WALL_PARALLEL) { px = 0; pz = 0; sx = 0; sz = 0; }
    else if (D.wall_kind == SDMD_WALL_CYLINDER) { px = 0; sx = 0; }
    double rx = px - D.wall_pos[0], ry = py - D.wall_pos[1], rz = pz - D.wall_pos[2];
    double vx = sx - 0.0, vy = sy - 0.0, vz = sz - 0.0;
    double d2 = rx * rx + ry * ry + rz * rz;
    double b = rx * vx + ry * vy + rz * vz;
    double mass = D.mass[i];
    double pot = D.ev_pot[i];
    int s_vel = 0;
    switch (D.ev_sub[i]) {
        case SDMD_SUB_WELLBOUNC: s_vel = -1; pot = 0; break;
        case SDMD_SUB_TBDCAPTR:
            if (b * b - 2 * d2 * pot / mass <= 0) { pot = 0; s_vel = 1; } else s_vel = -1;
            break;
        case SDMD_SUB_TBDESCAP:
            if (b * b - 2 * d2 * pot / mass <= 0) { pot = 0; s_vel = -1; } else s_vel = 1;
            break;
        default: D.ctl->error = SDMD_ERR_STATE; return;
    }
    double phi = (-1 * b + s_vel * sqrt(b * b - 2 * d2 * pot / mass)) / d2;
    if (phi != phi) { D.ctl->error = SDMD_ERR_STATE; return; }
    D.vx[i] = tvx + phi * rx; D.vy[i] = tvy + phi * ry; D.vz[i] = tvz + phi * rz;
}

// ObstEvent, Events.c:543-619 / SphObstEvent, Events.c:752-796: kick along the one-dimensional / radial separation
__device__ inline void d_obst_event(const DevP &D, int i, bool sph) {
    const double pos[3] = {D.x[i], D.y[i], D.z[i]}, spd[3] = {D.vx[i], D.vy[i], D.vz[i]};
    double r[3], v[3];
    int partner = D.ev_partner[i];
    if (sph) {
        int n = partner / -1000;
        if (n < 0 || n >= D.n_sphobst) { D.ctl->error = SDMD_ERR_STATE; return; }
        for (int k = 0; k < 3; k++) { r[k] = pos[k] - D.sb->sph_pos[n][k]; v[k] = spd[k] - 0.0; }
    } else {
        int n = partner / -10, dim = (partner * -1) % 10 - 1;
        if (n < 0 || n >= D.n_obst || dim < 0 || dim > 2) { D.ctl->error = SDMD_ERR_STATE; return; }
        for (int k = 0; k < 3; k++) {
            double pi = (k == dim) ? pos[k] : 0, pj = (k == dim) ? D.sb->obst_pos[n][k] : 0, si = (k == dim) ? spd[k] : 0;
            r[k] = pi - pj; v[k] = si - 0.0;
        }
    }
    double d2 = r[0] * r[0] + r[1] * r[1] + r[2] * r[2];
    double b = r[0] * v[0] + r[1] * v[1] + r[2] * v[2];
    double mass = D.mass[i];
    double pot = D.ev_pot[i];
    int s_vel = 0;
    int sub = D.ev_sub[i];
    if (sub == SDMD_SUB_CORECOLLI) { s_vel = 1; pot = 0; }
    else if (!sph && sub == SDMD_SUB_TBDCAPTR) { if (b * b - 2 * d2 * pot / mass <= 0) { pot = 0; s_vel = 1; } else s_vel = -1; }
    else if (!sph && sub == SDMD_SUB_TBDESCAP) { if (b * b - 2 * d2 * pot / mass <= 0) { pot = 0; s_vel = -1; } else s_vel = 1; }
    else { D.ctl->error = SDMD_ERR_STATE; return; }
    double phi = (-1 * b + s_vel * sqrt(b * b - 2 * d2 * pot / mass)) / d2;
    if (phi != phi) { D.ctl->error = SDMD_ERR_STATE; return; }
    D.vx[i] = spd[0] + phi * r[0]; D.vy[i] = spd[1] + phi * r[1]; D.vz[i] = spd[2] + phi * r[2];
}

// ChargeEvent, Events.c:685-749
__device__ inline void d_charge_event(const DevP &D, int i) {
    int cn = D.ev_partner[i] / -100;
    if (cn < 0 || cn >= D.n_charge) { D.ctl->error = SDMD_ERR_STATE; return; }
    double tvx = D.vx[i], tvy = D.vy[i], tvz = D.vz[i];
    double rx = D.x[i] - D.sb->chg_pos[cn][0], ry = D.y[i] - D.sb->chg_pos[cn][1], rz = D.z[i] - D.sb->chg_pos[cn][2];
    double vx = tvx - D.sb->chg_vel[cn][0], vy = tvy - D.sb->chg_vel[cn][1], vz = tvz - D.sb->chg_vel[cn][2];
    double d2 = rx * rx + ry * ry + rz * rz;
    double b = rx * vx + ry * vy + rz * vz;
    double mass = D.mass[i];
    double pot = D.ev_pot[i];
    int s_vel = 0;
    switch (D.ev_sub[i]) {
        case SDMD_SUB_TBDCAPTR:
            if (b * b - 2 * d2 * pot / mass <= 0) { pot = 0; s_vel = 1; } else s_vel = -1;
            break;
        case SDMD_SUB_TBDESCAP:
            if (b * b - 2 * d2 * pot / mass <= 0) { pot = 0; s_vel = -1; } else s_vel = 1;
            break;
        default: D.ctl->error = SDMD_ERR_STATE; return;
    }
    double phi = (-1 * b + s_vel * sqrt(b * b - 2 * d2 * pot / mass)) / d2;
    if (phi != phi) { D.ctl->error = SDMD_ERR_STATE; return; }
    D.vx[i] = tvx + phi * rx; D.vy[i] = tvy + phi * ry; D.vz[i] = tvz + phi * rz;
}

// TunnelEvent, Events.c:621-682: WallEvent's kick with the separation taken from the tunnel's axis
__device__ inline void d_tunnel_event(const DevP &D, int i) {
    int n = D.ev_partner[i] / -10000;
    if (n < 0 || n >= D.n_tunnel) { D.ctl->error = SDMD_ERR_STATE; return; }
    double tvx = D.vx[i], tvy = D.vy[i], tvz = D.vz[i];
    double rx = D.sb->tun_pos[n][1] - D.sb->tun_pos[n][1], ry = D.y[i] - D.sb->tun_pos[n][2], rz = D.z[i] - D.sb->tun_pos[n][3];
    double vx = 0 - 0.0, vy = tvy - 0.0, vz = tvz - 0.0;
    double d2 = rx * rx + ry * ry + rz * rz;
    double b = rx * vx + ry * vy + rz * vz;
    double mass = D.mass[i];
    double pot = D.ev_pot[i];
    int s_vel = 0;
    switch (D.ev_sub[i]) {
        case SDMD_SUB_WELLBOUNC: s_vel = -1; pot = 0; break;
        case SDMD_SUB_TBDCAPTR:
            if (b * b - 2 * d2 * pot / mass <= 0) { pot = 0; s_vel = 1; } else s_vel = -1;
            break;
        case SDMD_SUB_TBDESCAP:
            if (b * b - 2 * d2 * pot / mass <= 0) { pot = 0; s_vel = -1; } else s_vel = 1;
            break;
        default: D.ctl->error = SDMD_ERR_STATE; return;
    }
    double phi = (-1 * b + s_vel * sqrt(b * b - 2 * d2 * pot / mass)) / d2;
    if (phi != phi) { D.ctl->error = SDMD_ERR_STATE; return; }
    D.vx[i] = tvx + phi * rx; D.vy[i] = tvy + phi * ry; D.vz[i] = tvz + phi * rz;
}

// ThermostatEvent (Andersen), Events.c:396-420; RandomVelocity ToolFunctions.c:600-610
__device__ inline void d_thermostat_event(const DevP &D, int i) {
    double speed[3];
    double variance = D.ctl->temperature / D.mass[i];
    double sd = sqrt(variance);
    double netV[3] = {0, 0, 0};
    if (D.charge_com && D.flow_mark == 3) {   // CalCOMV of the atom's chain from the raw velocities, Events.c:400-402
        int a = i, b = i;
        while (a > 1 && D.protein[a - 1] == D.protein[i]) a--;
        while (b < D.n && D.protein[b + 1] == D.protein[i]) b++;
        for (int n = a; n <= b; n++) { netV[0] += D.vx[n] * D.mass[n]; netV[1] += D.vy[n] * D.mass[n]; netV[2] += D.vz[n] * D.mass[n]; }
        for (int k = 0; k < 3; k++) netV[k] /= D.protein_mass[i];
    }
#pragma unroll 1
    for (int k = 0; k < 3; k++) {
        // flow terms, Events.c:408-410
        const double a = D.flow_a ? D.flow_a[3 * i + k] : 0.0, rec = D.flow_time_rec ? D.flow_time_rec[i] : 0.0;
        do {
            speed[k] = d_next_gauss(D) * sd + 0;
            speed[k] += D.flow_v[k] + netV[k] + a * (D.ctl->currenttime - rec);
        } while (speed[k] == 0 && !D.ctl->error);
    }
    D.vx[i] = speed[0]; D.vy[i] = speed[1]; D.vz[i] = speed[2];
    if (D.flow_time_rec) D.flow_time_rec[i] = D.ctl->currenttime;   // :418
}

__device__ __forceinline__ int d_cell_of(const DevP &D, double x, double y, double z, int idx[3]) {
    double co[3] = {x, y, z};
#pragma unroll
    for (int k = 0; k < 3; k++) {
        idx[k] = (int)(co[k] / D.cellsize[k]);  // Events.c:132-133
        if (idx[k] == D.cellnum[k]) idx[k]--;   // Events.c:143-145
    }
    return idx[2] * D.cellnum[0] * D.cellnum[1] + idx[1] * D.cellnum[0] + idx[0];  // 0-based
}

// PBCandCrossCellEvent, Events.c:248-272 = PBC("part") :230-238 + LinkList("part") :158-199
__device__ inline void d_pcc_event(const DevP &D, int i) {
    int old[3] = {D.cell_idx[3 * i], D.cell_idx[3 * i + 1], D.cell_idx[3 * i + 2]};
    double co[3] = {D.x[i], D.y[i], D.z[i]};
    bool edge = false;
#pragma unroll
    for (int k = 0; k < 3; k++) edge = edge || old[k] == 0 || old[k] == D.cellnum[k] - 1;
    if (edge) {
#pragma unroll
        for (int k = 0; k < 3; k++) {
            if (!(D.pbc_mask >> k & 1)) continue;   // dimensions closed by a wall do not wrap, Events.c:210-216
            if (co[k] >= d_box(D, k)) co[k] -= d_box(D, k);
            else if (co[k] < 0) co[k] += d_box(D, k);
        }
        D.x[i] = co[0]; D.y[i] = co[1]; D.z[i] = co[2];
    }
    int idx[3];
    int newCell = d_cell_of(D, co[0], co[1], co[2], idx);
    int oldCell = old[2] * D.cellnum[0] * D.cellnum[1] + old[1] * D.cellnum[0] + old[0];
    D.cell_idx[3 * i] = idx[0]; D.cell_idx[3 * i + 1] = idx[1]; D.cell_idx[3 * i + 2] = idx[2];
    int *cl = D.celllist;
    int next = oldCell + 1 + D.n;
    if (cl[next] == 0) { D.ctl->error = SDMD_ERR_STATE; return; }
    while (cl[next] != i) {
        next = cl[next];
        if (next == 0) { D.ctl->error = SDMD_ERR_STATE; return; }
    }
    cl[next] = cl[i];
    next = newCell + 1 + D.n;
    cl[i] = cl[next];
    cl[next] = i;
    D.cell_cnt[oldCell]--;
    D.cell_cnt[newCell]++;
    i64 t = ++D.ctl->cc_epoch;   // only lists that look into one of the two cells are stale
    D.cell_epoch[oldCell] = t;
    D.cell_epoch[newCell] = t;
}

// DoEvent, Events.c:25-114
__device__ inline void d_do_event(const DevP &D, int i, int j) {
    Ctl *c = D.ctl;
    switch (D.ev_type[i]) {
        case SDMD_EV_COLI: d_interaction_event(D, i, j, 0); c->n_coli++; break;
        case SDMD_EV_BOND: d_interaction_event(D, i, j, 0); c->n_bond++; break;
        case SDMD_EV_HBNE: d_interaction_event(D, i, j, 0); c->n_hbne++; break;
        case SDMD_EV_HB: {
            int ni = D.hb_neighbor[i], nj = D.hb_neighbor[j];
            d_interaction_event(D, i, j, 1);
            d_hb_event(D, i, j, ni, nj);
            c->n_hb++;
            break;
        }
        case SDMD_EV_THER: d_thermostat_event(D, i); c->n_ther++; break;
        case SDMD_EV_PCC: d_pcc_event(D, i); c->n_pcc++; break;
        case SDMD_EV_WALL: d_wall_event(D, i); c->n_wall++; break;
        case SDMD_EV_OBST: d_obst_event(D, i, false); c->n_wall++; break;
        case SDMD_EV_TUNL: d_tunnel_event(D, i); c->n_wall++; break;
        case SDMD_EV_CHRG: d_charge_event(D, i); break;   // no counter, Events.c:93-95
        case SDMD_EV_SPHO: d_obst_event(D, i, true); c->n_wall++; break;
        default: c->error = SDMD_ERR_STATE; break;
    }
}

// ===================================================================================================
// CTA-parallel re-prediction of the dirty atoms (K3).
//
// Predict(renewList) (SGThread.c:158-211) for the 1-2 atoms of an event, executed by one thread block:
// every candidate of every predictor (bond partners, the 27-cell scan of CollisionTime, PCC, thermostat,
// bonded / free HB partners, HB neighbours) becomes one entry of a candidate array laid out in the exact
// order the reference visits them; all threads evaluate entries in parallel (own-side time, potential,
// sub-event), and the per-atom earliest-event fold (JobAssign, strict '<', first wins) is then taken over
// that array.  CheckDuplication couples the fold to the running best, so the array is cut at the entries
// whose partner's slot points back at this atom; the pieces in between collapse to their first minimum
// (warp REDUX), and one lane replays the few pieces in order.
//
// The loop is latency bound (one event at a time): the 27-cell candidate list of an atom is cached between link-list
// changes, the minimum-image shifts of its 27 neighbour cells and its HB lists are staged in shared memory by the
// warp that lays the array out, and the static connectionMap row is one load from a dense matrix (small systems).
// ===================================================================================================
#define PC_MAXC 448          // candidates per atom held in shared memory; larger neighbourhoods use the serial path
enum { CK_BOND = 0, CK_COLL = 1, CK_PCC = 2, CK_THER = 3, CK_HB = 4, CK_HBFREE = 5, CK_HBNE = 6, CK_NONE = 7 };
#define CF_VALID 1           // own-side event exists and its time is > 0
#define CF_COUNTED 2         // a CalculateParameters-equivalent evaluation happened on the own-side path
#define CF_ELIGIBLE 4        // partner's slot points back (CheckDuplication's first three conditions)
#define PC_MAXE 3            // eligible entries handled by the segmented fold (more: serial replay)

struct PredShared {
    int natoms;
    int atom[2];
    double gauss[2];
    int gathered;             // candidate arrays already laid out (overlapped with the event)
    int use_serial[2];
    int ncand[2];
    int hb_free[2];
    int n_bond[2], n_coll[2], n_hbne[2];
    int off_coll[2], off_pcc[2], off_ther[2], off_hb[2], off_hbne[2];
    Slot result[2];
    int winner_pos[2];
    int npair[2];
    int hbn[2][4];               // neighborPartner[1..4]
    int hb_bond[2];
    int cc_row[2];
    int n_pass[2];
    double shift[2][27][3];      // minimum-image shift of neighbour-cell slot k (PBCShift, ToolFunctions.c:801-817)
    int elig_n[2];
    int elig_pos[2][8];
    // candidate arrays: atom a occupies [a * PC_MAXC, (a + 1) * PC_MAXC)
    int cj[2 * PC_MAXC];
    unsigned char ckind[2 * PC_MAXC];
    unsigned char caux[2 * PC_MAXC];   // neighbour-cell slot k / bond slot / HB-neighbour index n
    unsigned char csub[2 * PC_MAXC];
    unsigned char cflag[2 * PC_MAXC];
    double ctime[2 * PC_MAXC];
    double cpot[2 * PC_MAXC];
    // tournament-tree repair scratch
    int path_old[40], path_sib[40];
    double path_tsib[40];
};

// clock64() that cannot issue before a shared-memory word is read back: BAR.SYNC is deferred-blocking, so a bare
// clock read right after __syncthreads() samples the barrier's ISSUE time, not its release
__device__ __forceinline__ long long d_clock_after(const int *smem_word) {
    int v = *(volatile const int *)smem_word;
    long long t;
    asm volatile("{ .reg .b32 d; mov.b32 d, %1; mov.u64 %0, %%clock64; }" : "=l"(t) : "r"(v) : "memory");
    return t;
}

__device__ __forceinline__ int d_kind_event(int kind) {
    return kind == CK_BOND ? SDMD_EV_BOND : kind == CK_COLL ? SDMD_EV_COLI : kind == CK_PCC ? SDMD_EV_PCC
         : kind == CK_THER ? SDMD_EV_THER : kind == CK_HBNE ? SDMD_EV_HBNE : SDMD_EV_HB;
}

// FindPair's step walk with every table load issued up front (same comparisons, same first match)
__device__ __forceinline__ FindRes d_find_pair_table_fast(const DevP &D, const sdmd_step_table *t, int mode, double direction,
                                                          double dist2, int bond_a1, int bond_a2) {
    FindRes r;
    r.cs = -1;
    double dir = (direction < 0) ? 1.0 : -1.0;
    double tol = dir * SD_ZERO;
    double dmin2 = __ldg(&t->dmin2), dmax2 = __ldg(&t->dmax2);
    int ns = __ldg(&t->nstep);
    if (dmin2 == 0) D.ctl->error = SDMD_ERR_TABLE;   // fatal in the reference, ToolFunctions.c:114-117
    if (dist2 < dmin2 - SD_ZERO) {
        r.lower = SD_INF; r.upper = SD_INF; r.accum = SD_INF; r.lowPot = 0; r.upPot = 0;
        if (!(mode == FP_NEIGHBOR && bond_a2 == 0) && !(mode == FP_HB && bond_a1 == 0) && !D.n_tunnel)   /* :127-130: not under tunnels */ atomicAdd((unsigned long long *)&D.ctl->warnings, 1ULL);
        return r;
    }
    int kk = ns;
    for (int k = ns - 1; k >= 0; k--) {   // first step with dist2 < d + tol (steps ascend, ToolFunctions.c:144-145)
        double d = __ldg(&t->d2[k]);
        if (dist2 < d + tol) kk = k;
    }
    r.lower = dmin2; r.lowPot = 0; r.upper = 0; r.upPot = 0; r.accum = 0;
    if (kk > 0) { r.lower = __ldg(&t->d2[kk - 1]); r.lowPot = __ldg(&t->e[kk - 1]); }
    if (kk < ns) {
        r.upper = __ldg(&t->d2[kk]);
        r.upPot = -1 * __ldg(&t->e[kk]);
        r.accum = __ldg(&t->acc[kk]);
        if (kk == 0) { r.cs = 0; return r; }
        if (mode == FP_HB && kk == ns - 1 && bond_a1 > 0) { r.cs = 1; return r; }
        if (mode == FP_NEIGHBOR && kk == ns - 1 && bond_a2 > 0) { r.upPot = 0; r.cs = 1; return r; }
        return r;
    }
    if (mode == FP_HB && bond_a1 > 0) atomicAdd((unsigned long long *)&D.ctl->warnings, 1ULL);
    if (dmax2 > 0) {
        if (dist2 < dmax2 + SD_ZERO) {
            r.upper = dmax2; r.upPot = 0; r.accum = 0;
            r.cs = 1;
            if (ns == 0 && dir == 1.0) r.cs = 0;
            return r;
        }
        r.lower = SD_INF; r.upper = SD_INF; r.accum = SD_INF; r.lowPot = 0; r.upPot = 0;
        if (!(mode == FP_NEIGHBOR && bond_a2 == 0) && !(mode == FP_HB && bond_a1 == 0)) atomicAdd((unsigned long long *)&D.ctl->warnings, 1ULL);
        return r;
    }
    r.upper = SD_INF; r.upPot = SD_INF; r.accum = SD_INF;
    r.cs = 1;
    return r;
}

__device__ inline FindRes d_find_pair_fast(const DevP &D, int a1, int a2, int mode, double direction, double dist2, int hbPartner) {
    int idx = -1;
    if (mode == FP_COLLISION) {
        idx = __ldg(&D.coll_index[D.type[a1] * SD_NT + D.type[a2]]);
    } else if (mode == FP_CONSTRAINT) {
        for (int k = D.con_off[a1]; k < D.con_off[a1 + 1]; k++)
            if (D.con_j[k] == a2) { idx = D.con_table[k]; break; }
    } else if (mode == FP_HB) {
        int cls = d_hb_model(D, a1, a2);
        if (cls >= 0) idx = __ldg(&D.hb_index[(cls * SD_NT + 0) * SD_NT + 0]);
    } else {
        int cls = hbPartner > 0 ? d_hb_model(D, a2, hbPartner) : -2;
        int t1 = d_type_change(D.type[a1], 0), t2 = d_type_change(D.type[a2], 0);
        if (cls >= 0) idx = __ldg(&D.hb_index[(cls * SD_NT + t1) * SD_NT + t2]);
    }
    if (idx < 0) {
        D.ctl->error = SDMD_ERR_TABLE;
        FindRes r;
        r.lower = SD_INF; r.upper = SD_INF; r.accum = SD_INF; r.lowPot = 0; r.upPot = 0; r.cs = -1;
        return r;
    }
    return d_find_pair_table_fast(D, &D.tables[idx], mode, direction, dist2, D.hb_bond[a1], D.hb_bond[a2]);
}

// ---- phase A: lay out the candidate array of predict-atom `a` (executed by ONE warp per atom) ----
// The 27-cell candidate list of an atom only changes when the link list or the HB state changes (PCC / HB
// events, < 1 % of events on proteins), so it is cached per atom in HBM (L1/L2 resident for small systems) and
// re-walked only when its epoch stamp is stale.  On a miss, per-cell occupancy counts give every neighbour
// cell's offset with one warp scan, so each list is walked exactly once.
__device__ inline void d_pc_gather_warp(const DevP &D, PredShared &sh, int a) {
    int lane = threadIdx.x & 31;
    int i = sh.atom[a];
    int base = a * PC_MAXC;
    int cx = D.cell_idx[3 * i], cy = D.cell_idx[3 * i + 1], cz = D.cell_idx[3 * i + 2];
    int cap = D.cc_cap;
    i64 epoch = D.ctl->cc_epoch;
    int row = D.cc_per_atom ? i : a;
    i64 stamp = D.cc_per_atom ? D.cc_stamp[i] : -1;
    int role = 0, bonded = 0, itype = 0, nhbne = 0;
    if (D.hydrogen_bond) {
        role = D.hb_role[i];
        bonded = D.hb_bond[i];
        itype = D.hb_itype[i];
        nhbne = D.hbn_status[i];
    }
    bool can_hbfree = D.hydrogen_bond && (role == 'A' || role == 'D' || role == 'S');
    int cell = 0;
    if (lane < 27) {
        int ncx, ncy, ncz;
        cell = d_neighbor_cell(D, cx, cy, cz, lane, &ncx, &ncy, &ncz);
        sh.shift[a][lane][0] = d_shift1(cx, ncx, D.cellnum[0], d_box(D, 0));
        sh.shift[a][lane][1] = d_shift1(cy, ncy, D.cellnum[1], d_box(D, 1));
        sh.shift[a][lane][2] = d_shift1(cz, ncz, D.cellnum[2], d_box(D, 2));
    }
    // valid while nothing in the 27-cell neighbourhood changed after the list was cached (and no new link list since)
    bool fresh = lane >= 27 || D.cell_epoch[cell] <= stamp;
    bool hit = stamp >= D.ctl->cc_floor && __all_sync(0xffffffffu, fresh);
    int ncoll;
    if (hit) {
        ncoll = D.cc_n[i];
    } else {
        int cnt = (lane < 27) ? D.cell_cnt[cell] - (lane == 0 ? 1 : 0) : 0;   // offset 0 is the atom's own cell
        int incl = cnt;
        for (int o = 1; o < 32; o <<= 1) {
            int v = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += v;
        }
        ncoll = __shfl_sync(0xffffffffu, incl, 31);
        int off = incl - cnt;
        if (ncoll <= cap) {
            if (lane < 27) {
                int w = row * cap + off;
                int ti = D.type[i], pi = D.protein[i], ai = D.aa_num[i];
                for (int j = D.celllist[D.n + cell + 1]; j; j = D.celllist[j]) {
                    if (j == i) continue;
                    unsigned char pass = 0;
                    if (can_hbfree) {
                        // role / bonded / sequence-separation filters of the free-HB scan, TimePrediction.c:418-429
                        int rj = D.hb_role[j];
                        bool ad = (role == 'A' && rj == 'D') || (role == 'D' && rj == 'A');
                        bool sd = (role == 'S' && D.type[j] == 2) || (rj == 'S' && ti == 2);
                        bool ss = (role == 'S' && rj == 'S');
                        int d = ai - D.aa_num[j];
                        if (d < 0) d = -d;
                        pass = (ad || sd || ss) && D.hb_bond[j] == 0 && ((pi == D.protein[j] && d >= 4) || pi != D.protein[j]);
                    }
                    D.cc_j[w] = j; D.cc_k[w] = (unsigned char)lane; D.cc_pass[w] = pass;
                    w++;
#if SDMD_LZ
                    d_prefetch_atom(D, j);      // (large systems: see the end of this function)
#endif
                }
            }
            __syncwarp();
            // compact list of the entries that pass the free-HB filters (positions in scan order)
            int npass = 0;
            if (can_hbfree) {
                for (int p0 = 0; p0 < ncoll; p0 += 32) {
                    int p = p0 + lane;
                    bool ps = p < ncoll && D.cc_pass[row * cap + p];
                    unsigned m = __ballot_sync(0xffffffffu, ps);
                    if (ps) {
                        int slot = npass + __popc(m & ((1u << lane) - 1));
                        if (slot < 32) D.cc_passidx[row * 32 + slot] = (unsigned short)p;
                    }
                    npass += __popc(m);
                }
            }
            if (lane == 0) {
                D.cc_npass[row] = npass;
                if (D.cc_per_atom) { D.cc_n[i] = ncoll; D.cc_stamp[i] = epoch; }
            }
        }
    }
    __syncwarp();
    // section layout (every lane computes the same small thing)
    int nb = D.gel_mode ? 0 : (D.bond_off[i + 1] - D.bond_off[i]);
    int nhb = 0, hbfree = 0;
    if (D.hydrogen_bond && itype != SDMD_HB_SKIP) {   // TimePrediction.c:309-327
        if (bonded > 0) nhb = 1;
        else if (can_hbfree && bonded == 0) { hbfree = 1; nhb = ncoll; }
    }
    int off_coll = nb;
    int off_pcc = off_coll + ncoll;
    int off_ther = off_pcc + 1;
    int off_hb = off_ther + (D.thermostat_on ? 1 : 0);
    int off_hbne = off_hb + nhb;
    int total = off_hbne + nhbne;
    int npass_row = hbfree ? D.cc_npass[row] : 0;
    bool serial = total > PC_MAXC || nb > 27 || ncoll > cap || npass_row > 32;
    if (lane == 0) {
        sh.n_bond[a] = nb; sh.n_coll[a] = ncoll; sh.n_hbne[a] = nhbne; sh.hb_free[a] = hbfree;
        sh.off_coll[a] = off_coll; sh.off_pcc[a] = off_pcc; sh.off_ther[a] = off_ther; sh.off_hb[a] = off_hb;
        sh.off_hbne[a] = off_hbne; sh.ncand[a] = total; sh.use_serial[a] = serial;
        sh.hb_bond[a] = bonded;
        sh.cc_row[a] = row;
        sh.n_pass[a] = npass_row;
    }
    if (serial) return;
    if (lane < nb) {
        int p = base + lane;
        sh.cj[p] = D.bond_j[D.bond_off[i] + lane]; sh.ckind[p] = CK_BOND; sh.caux[p] = (unsigned char)lane;
    }
    if (lane == 31) {
        int p = base + off_pcc;
        sh.cj[p] = 0; sh.ckind[p] = CK_PCC; sh.caux[p] = 0;
        if (D.thermostat_on) { p = base + off_ther; sh.cj[p] = 0; sh.ckind[p] = CK_THER; sh.caux[p] = 0; }
        if (!hbfree && nhb == 1) { p = base + off_hb; sh.cj[p] = bonded; sh.ckind[p] = CK_HB; sh.caux[p] = 0; }
    }
    if (lane >= 27 && lane < 31) {
        int n = lane - 27 + 1;
        int hp = D.hydrogen_bond ? D.hbn_partner[5 * i + n] : 0;
        sh.hbn[a][n - 1] = hp;
        if (n <= nhbne) {
            int p = base + off_hbne + n - 1;
            sh.cj[p] = hp; sh.ckind[p] = CK_HBNE; sh.caux[p] = (unsigned char)n;
        }
    }
#if SDMD_LZ
    // Large systems: the state of the candidates sits in DRAM and the evaluation phase would meet it as two dependent round trips
    // (list entry -> the partner's coordinates, velocities and slot).  The layout warp asks for all of it now, while thread 0 is still
    // executing the event, so the evaluating threads find it in this SM's L1.
    // (a list that had to be rebuilt above has asked for them entry by entry already)
    if (hit)
        for (int p = lane; p < ncoll; p += 32) d_prefetch_atom(D, D.cc_j[row * cap + p]);
#endif
}

// ---- phase B: own-side evaluation of one candidate entry (any thread) ----
// lean path for the cell-scan entries of CollisionTime (TimePrediction.c:82-176)
__device__ __forceinline__ void d_pc_eval_coll(const DevP &D, PredShared &sh, int a, int p, const Self &me) {
    int q = a * PC_MAXC + p;
    int i = me.i;
    int ce = sh.cc_row[a] * D.cc_cap + (p - sh.off_coll[a]);
    int j = D.cc_j[ce];
    int k = D.cc_k[ce];
    sh.cj[q] = j; sh.ckind[q] = CK_COLL; sh.caux[q] = (unsigned char)k;
    int flag = 0, sub = SDMD_SUB_INVALID;
    double time = SD_INF, pot = 0;
    // independent loads first
    double xj = d_rx(D, j), yj = d_ry(D, j), zj = d_rz(D, j), vxj = D.vx[j], vyj = D.vy[j], vzj = D.vz[j];
    int pj = D.ev_partner[j];
    int tyj = D.ev_type[j];
    i64 pcj = D.ev_pcounter[j];
    int typej = D.type[j];
    double shx = sh.shift[a][k][0], shy = sh.shift[a][k][1], shz = sh.shift[a][k][2];
    int cm = 0, con_tbl = 0;
    if (D.cm_dense) {   // static connectionMap entry (BOND / CONSTRAINT bits + constraint table), dense for small systems
        unsigned v = __ldg(&D.cm_dense[(size_t)i * (D.n + 1) + j]);
        cm = v & 0xff; con_tbl = v >> 8;
    } else {
        for (int e = me.cm0; e < me.cm1; e++)
            if (D.cm_j[e] == j) { cm = D.cm_flags[e]; con_tbl = D.cm_table[e]; }
    }
    bool skip = (cm & SDMD_BOND_CONNECT) || sh.hb_bond[a] == j;
    // FindElemInList over neighborPartner stops at the first 0 (ToolFunctions.c:1078-1084)
    {
        int h0 = sh.hbn[a][0], h1 = sh.hbn[a][1], h2 = sh.hbn[a][2], h3 = sh.hbn[a][3];
        bool in = (h0 == j) || (h0 != 0 && (h1 == j || (h1 != 0 && (h2 == j || (h2 != 0 && h3 == j)))));
        skip = skip || in;
    }
    if (pj == i && tyj == SDMD_EV_COLI && pcj == me.counter) flag |= CF_ELIGIBLE;
    if (!skip) {
        Pair pr = d_pair(me.x, me.y, me.z, me.vx, me.vy, me.vz, xj, yj, zj, vxj, vyj, vzj, shx, shy, shz);
        flag |= CF_COUNTED;
        if (pr.r2 <= D.cut2) {
            FindRes f;
            bool none = false;
            if (cm & SDMD_CONSTRAINT_CONNECT) {
                f = d_find_pair_table_fast(D, &D.tables[con_tbl], FP_CONSTRAINT, pr.b, pr.r2, 0, 0);
            } else {
                int idx = __ldg(&D.coll_index[me.type * SD_NT + typej]);
                if (idx < 0) { D.ctl->error = SDMD_ERR_TABLE; idx = 0; }
                f = d_find_pair_table_fast(D, &D.tables[idx], FP_COLLISION, pr.b, pr.r2, 0, 0);
                if (f.lowPot == 0 && f.upPot == 0) none = true;
            }
            if (!none) {
                int s; double d2;
                if (pr.b < 0 && d_disc(pr, f.lower) > 0) {
                    s = -1; d2 = f.lower; pot = f.lowPot;
                    sub = (f.cs == 0) ? SDMD_SUB_CORECOLLI : SDMD_SUB_TBDCAPTR;
                } else {
                    s = 1; d2 = f.upper; pot = f.upPot;
                    if (D.method_ding) {
                        if (cm & SDMD_CONSTRAINT_CONNECT) sub = (f.cs == 1) ? SDMD_SUB_WELLBOUNC : SDMD_SUB_TBDESCAP;
                        else { if (f.cs == 1) none = true; sub = SDMD_SUB_TBDESCAP; }
                    }
                }
                if (!none) {
                    time = d_time(pr, s, d2);
                    if (time < 0) atomicAdd((unsigned long long *)&D.ctl->error_prints, 1ULL);
                    if (time > 0) flag |= CF_VALID;
                }
            }
        }
    }
    sh.ctime[q] = time; sh.cpot[q] = pot; sh.csub[q] = (unsigned char)sub; sh.cflag[q] = (unsigned char)flag;
}

// every other predictor's entries
__device__ inline void d_pc_eval(const DevP &D, PredShared &sh, int a, int p, const Self &me) {
    int q = a * PC_MAXC + p;
    int i = me.i;
    if (sh.hb_free[a] && p >= sh.off_hb[a] && p < sh.off_hbne[a]) {
        // free-HB scan entry (passed the role filters): same candidate list as the collision scan
        int ce = sh.cc_row[a] * D.cc_cap + (p - sh.off_hb[a]);
        sh.cj[q] = D.cc_j[ce]; sh.caux[q] = D.cc_k[ce];
        sh.ckind[q] = CK_HBFREE;
    }
    int kind = sh.ckind[q];
    if (kind == CK_NONE) return;
    int j = sh.cj[q];
    int flag = 0, sub = SDMD_SUB_INVALID;
    double time = SD_INF, pot = 0;
    if (kind == CK_PCC) {
        Slot tmp = d_reset_slot(0, 0, 0);
        d_pcc_time(D, me, tmp);
        if (tmp.type == SDMD_EV_PCC) { time = tmp.time; flag = CF_VALID; }
    } else if (kind == CK_THER) {
        double g = sh.gauss[a];
        time = D.thermoF * ((g < 0) ? (-g) : (g));
        if (time > 0) flag = CF_VALID;
    } else {
        // eligibility for CheckDuplication (TimePrediction.c:1102-1104)
        if (D.ev_partner[j] == i && D.ev_type[j] == d_kind_event(kind) && D.ev_pcounter[j] == me.counter) flag |= CF_ELIGIBLE;
        Pair pr = d_pair_atoms(D, i, j);
        flag |= CF_COUNTED;
        if (kind == CK_BOND) {
            int k = D.bond_off[i] + sh.caux[q];
            double long2 = D.bond_dmax2[k], short2 = D.bond_dmin2[k];
            bool done = false;
            if (pr.r2 - long2 > SD_ZERO) {
                atomicAdd((unsigned long long *)&D.ctl->error_prints, 1ULL);
                if (pr.b >= 0) { time = 1e-16 * i * j; sub = SDMD_SUB_WELLBOUNC; done = true; }
            } else if (short2 - pr.r2 > SD_ZERO) {
                atomicAdd((unsigned long long *)&D.ctl->error_prints, 1ULL);
                if (pr.b < 0) { time = 1e-16 * i * j; sub = SDMD_SUB_CORECOLLI; done = true; }
            }
            if (!done) {
                int s; double d2;
                if (pr.b < 0 && d_disc(pr, short2) > 0) { s = -1; d2 = short2; sub = SDMD_SUB_CORECOLLI; }
                else { s = 1; d2 = long2; sub = SDMD_SUB_WELLBOUNC; }
                time = d_time(pr, s, d2);
                if (time < 0) { atomicAdd((unsigned long long *)&D.ctl->error_prints, 1ULL); time = SD_INF; }
            }
            pot = 0;
            if (time > 0) flag |= CF_VALID;
        } else if (kind == CK_HB) {
            FindRes f = d_find_pair_fast(D, i, j, FP_HB, pr.b, pr.r2, 0);
            int s; double d2;
            if (pr.b < 0 && d_disc(pr, f.lower) > 0) {
                s = -1; d2 = f.lower; pot = f.lowPot;
                sub = (f.cs == 0) ? SDMD_SUB_CORECOLLI : SDMD_SUB_TBDCAPTR;
            } else {
                s = 1; d2 = f.upper; pot = f.upPot;
                if (D.method_ding) sub = (f.cs == 1) ? SDMD_SUB_TBDHBBRIK : SDMD_SUB_TBDESCAP;
            }
            time = d_time(pr, s, d2);
            if (time < 0) { atomicAdd((unsigned long long *)&D.ctl->error_prints, 1ULL); time = SD_INF; }
            if (time > 0) flag |= CF_VALID;
        } else if (kind == CK_HBFREE) {
            if (pr.r2 <= D.cut2 && d_check_pair_match(D, i, j) >= 0) {
                FindRes f = d_find_pair_fast(D, i, j, FP_HB, pr.b, pr.r2, 0);
                if (pr.b < 0 && f.cs == 1 && d_disc(pr, f.lower) > 0) {
                    time = d_time(pr, -1, f.lower);
                    pot = f.lowPot; sub = SDMD_SUB_TBDHBFORM;
                    if (time < 0) { atomicAdd((unsigned long long *)&D.ctl->error_prints, 1ULL); time = SD_INF; }
                    if (time > 0) flag |= CF_VALID;
                }
            }
        } else {  // CK_HBNE
            int host, nbr, third;
            int r = D.hb_role[i];
            int bi = D.hb_bond[i];
            if ((r == 'A' || r == 'D' || r == 'S') && (bi && D.hb_neighbor[bi] == j)) { host = i; nbr = j; third = bi; }
            else { host = j; nbr = i; third = D.hb_bond[j]; }
            FindRes f = d_find_pair_fast(D, nbr, host, FP_NEIGHBOR, pr.b, pr.r2, third);
            int s; double d2;
            if (pr.b < 0 && d_disc(pr, f.lower) > 0) {
                s = -1; d2 = f.lower; pot = f.lowPot;
                sub = (f.cs == 0) ? SDMD_SUB_CORECOLLI : SDMD_SUB_TBDCAPTR;
            } else {
                s = 1; d2 = f.upper; pot = f.upPot;
                if (D.method_ding) sub = (f.cs == 1) ? SDMD_SUB_WELLBOUNC : SDMD_SUB_TBDESCAP;
            }
            time = d_time(pr, s, d2);
            if (time < 0) { atomicAdd((unsigned long long *)&D.ctl->error_prints, 1ULL); time = SD_INF; }
            if (time > 0) flag |= CF_VALID;
        }
    }
    sh.ctime[q] = time; sh.cpot[q] = pot; sh.csub[q] = (unsigned char)sub; sh.cflag[q] = (unsigned char)flag;
}

// warp-wide lexicographic (time, position) minimum with three REDUX steps; times are positive doubles, so
// their bit patterns order like unsigned integers.  Result is uniform across the warp; pos = INT_MAX if empty.
__device__ __forceinline__ void d_warp_firstmin(double &t, int &p) {
    unsigned long long bits = (p == 0x7fffffff) ? 0xffffffffffffffffULL : (unsigned long long)__double_as_longlong(t);
    unsigned hi = (unsigned)(bits >> 32), lo = (unsigned)bits;
    unsigned mh = __reduce_min_sync(0xffffffffu, hi);
    unsigned ml = __reduce_min_sync(0xffffffffu, hi == mh ? lo : 0xffffffffu);
    unsigned mp = __reduce_min_sync(0xffffffffu, (hi == mh && lo == ml) ? (unsigned)p : 0x7fffffffu);
    p = (int)mp;
    t = __longlong_as_double((long long)(((unsigned long long)mh << 32) | ml));
}

// ---- phase C: the ordered fold of atom `a` (warp 0) ----
__device__ inline void d_pc_fold(const DevP &D, PredShared &sh, int a) {
    int lane = threadIdx.x & 31;
    int i = sh.atom[a];
    int base = a * PC_MAXC;
    int nc = sh.ncand[a];
    i64 counter_i = D.ev_counter[i];
    int other = sh.atom[a ^ 1];
    int npre = sh.off_pcc[a];
    // pass 1: refresh eligibility of entries whose partner is the co-predicted atom (its slot may just have been
    // written), collect eligible positions, count pair evaluations
    int cnt = 0;
    for (int p0 = 0; p0 < nc; p0 += 32) {
        int p = p0 + lane;
        bool el = false;
        if (p < nc) {
            int q = base + p;
            int f = sh.cflag[q];
            int j = sh.cj[q];
            int kind = sh.ckind[q];
            if (j == other && j != 0 && kind != CK_NONE) {
                bool e2 = D.ev_partner[j] == i && D.ev_type[j] == d_kind_event(kind) && D.ev_pcounter[j] == counter_i;
                f = e2 ? (f | CF_ELIGIBLE) : (f & ~CF_ELIGIBLE);
                sh.cflag[q] = (unsigned char)f;
            }
            el = (f & CF_ELIGIBLE) != 0;
            cnt += (f & CF_COUNTED) ? 1 : 0;
        }
        unsigned m = __ballot_sync(0xffffffffu, el);
        if (m) {   // append eligible positions in scan order
            int n_before = sh.elig_n[a];
            if (el) {
                int slot = n_before + __popc(m & ((1u << lane) - 1));
                if (slot < 8) sh.elig_pos[a][slot] = p;
            }
            __syncwarp();
            if (lane == 0) sh.elig_n[a] = n_before + __popc(m);
            __syncwarp();
        }
    }
    cnt = __reduce_add_sync(0xffffffffu, cnt);
    int ne = sh.elig_n[a];
    Slot cur = d_reset_slot(D.ev_pot[i], D.ev_sub[i], D.ev_pcounter[i]);
    int win = -1, npair = cnt;
    SlotSrc S = d_live_slots(D);
    bool need_serial = ne > PC_MAXE;
    if (!need_serial) {
        int e0 = ne > 0 ? sh.elig_pos[a][0] : 0x7fffffff;
        int e1 = ne > 1 ? sh.elig_pos[a][1] : 0x7fffffff;
        int e2 = ne > 2 ? sh.elig_pos[a][2] : 0x7fffffff;
        // segments between eligible entries: first minimum of own-side values; plus the first minimum of the
        // pair sections in front of PCC (for the partnerCounter carry-over when a partner-less entry wins)
        double t0 = SD_INF, t1 = SD_INF, t2 = SD_INF, t3 = SD_INF, tp = SD_INF;
        int p0_ = 0x7fffffff, p1_ = 0x7fffffff, p2_ = 0x7fffffff, p3_ = 0x7fffffff, pp = 0x7fffffff;
        for (int p = lane; p < nc; p += 32) {
            int q = base + p;
            if (!(sh.cflag[q] & CF_VALID)) continue;
            if (p == e0 || p == e1 || p == e2) continue;
            double t = sh.ctime[q];
            if (p < e0) { if (t < t0) { t0 = t; p0_ = p; } }
            else if (p < e1) { if (t < t1) { t1 = t; p1_ = p; } }
            else if (p < e2) { if (t < t2) { t2 = t; p2_ = p; } }
            else { if (t < t3) { t3 = t; p3_ = p; } }
            if (p < npre && t < tp) { tp = t; pp = p; }
        }
        d_warp_firstmin(t0, p0_);
        if (ne > 0) d_warp_firstmin(t1, p1_);
        if (ne > 1) d_warp_firstmin(t2, p2_);
        if (ne > 2) d_warp_firstmin(t3, p3_);
        if (ne == 0) d_warp_firstmin(tp, pp);
        if (lane == 0) {
            double st[4] = {t0, t1, t2, t3};
            int spos[4] = {p0_, p1_, p2_, p3_};
            int epos[3] = {e0, e1, e2};
#pragma unroll
            for (int k = 0; k < 4; k++) {
                if (k <= ne && spos[k] != 0x7fffffff) {
                    int q = base + spos[k];
                    if (d_job_assign(D, cur, sh.cj[q], st[k], sh.cpot[q], sh.csub[q], d_kind_event(sh.ckind[q]))) win = spos[k];
                }
                if (k < 3 && k < ne) {
                    int p = epos[k];
                    int q = base + p;
                    int f = sh.cflag[q];
                    int j = sh.cj[q];
                    int ev = d_kind_event(sh.ckind[q]);
                    if (cur.time > d_stime(D, S, j)) {   // CheckDuplication copies the partner's slot
                        cur.partner = j; cur.pot = S.pot[j]; cur.sub = S.sub[j]; cur.time = d_stime(D, S, j); cur.type = ev;
                        cur.pcounter = D.ev_counter[j];
                        win = p;
                        if (f & CF_COUNTED) npair--;
                    } else if ((f & CF_VALID) && d_job_assign(D, cur, j, sh.ctime[q], sh.cpot[q], sh.csub[q], ev)) win = p;
                }
            }
            // partner-less winner: partnerCounter keeps the counter of the last pair entry that was a running best
            if (win >= 0 && sh.cj[base + win] == 0) {
                if (ne == 0) {
                    if (pp != 0x7fffffff && tp < SD_INF) cur.pcounter = D.ev_counter[sh.cj[base + pp]];
                } else need_serial = true;
            }
        }
        need_serial = __shfl_sync(0xffffffffu, (int)need_serial, 0) != 0;
    }
    if (need_serial && lane == 0) {
        // exact in-order replay (rare: many eligible entries, or eligible entries with a partner-less winner)
        cur = d_reset_slot(D.ev_pot[i], D.ev_sub[i], D.ev_pcounter[i]);
        win = -1;
        npair = 0;
        for (int p = 0; p < nc; p++) {
            int q = base + p;
            int f = sh.cflag[q];
            if (sh.ckind[q] == CK_NONE) continue;
            int j = sh.cj[q];
            int ev = d_kind_event(sh.ckind[q]);
            if ((f & CF_ELIGIBLE) && cur.time > d_stime(D, S, j)) {
                cur.partner = j; cur.pot = S.pot[j]; cur.sub = S.sub[j]; cur.time = d_stime(D, S, j); cur.type = ev;
                cur.pcounter = D.ev_counter[j];
                win = p;
                continue;
            }
            if (f & CF_COUNTED) npair++;
            if ((f & CF_VALID) && d_job_assign(D, cur, j, sh.ctime[q], sh.cpot[q], sh.csub[q], ev)) win = p;
        }
    }
    if (lane == 0) {
        sh.result[a] = cur;
        sh.winner_pos[a] = win;
        sh.npair[a] = npair;
    }
}

// evaluate the entries [lo, hi) of both atoms with the thread group [g0, g0 + gsize)
template <bool COLL>
__device__ __forceinline__ void d_pc_eval_group(const DevP &D, PredShared &sh, const Self &me0, const Self &me1, int lo0, int hi0,
                                                int lo1, int hi1, int g0, int gsize) {
    int gt = (int)threadIdx.x - g0;
    if (gt < 0 || gt >= gsize) return;
    int n0 = hi0 - lo0, n1 = hi1 - lo1;
    for (int w = gt; w < n0 + n1; w += gsize) {
        int a = (w < n0) ? 0 : 1;
        int p = a ? lo1 + (w - n0) : lo0 + w;
        if (COLL) d_pc_eval_coll(D, sh, a, p, a ? me1 : me0);
        else d_pc_eval(D, sh, a, p, a ? me1 : me0);
    }
}

// candidate layout for all listed atoms: warp w lays out atom w (call with all threads, no barrier inside)
__device__ __forceinline__ void d_predict_gather(const DevP &D, PredShared &sh, int first_warp, int natoms) {
    int w = (int)(threadIdx.x >> 5) - first_warp;
    if (w >= 0 && w < natoms) {
        if ((threadIdx.x & 31) == 0) sh.elig_n[w] = 0;
        d_pc_gather_warp(D, sh, w);
    }
}

// Predict(renewList) by the whole CTA.  sh.natoms / sh.atom[] / sh.gauss[] are set and the candidate arrays laid
// out (d_predict_gather) before the barrier that precedes this call; returns with warp 0 holding the stored slots
// (other warps run ahead to the caller's next barrier).
__device__ inline void d_predict_cta(const DevP &D, PredShared &sh) {
    int tid = threadIdx.x;
    int na = sh.natoms;
    long long pc1 = d_clock_after(&sh.natoms);
    {
        bool on0 = na > 0 && !sh.use_serial[0] && !D.is_cons[sh.atom[0]];
        bool on1 = na > 1 && !sh.use_serial[1] && !D.is_cons[sh.atom[1]];
        if (on0 || on1) {
            Self me0 = d_load_self(D, sh.atom[0]);
            Self me1 = (na > 1) ? d_load_self(D, sh.atom[1]) : me0;
            int nthr = blockDim.x;
            // kind-uniform thread groups keep warps from diverging across predictors
            int gc = nthr - 96;   // cell-scan group
            d_pc_eval_group<true>(D, sh, me0, me1, on0 ? sh.off_coll[0] : 0, on0 ? sh.off_pcc[0] : 0, on1 ? sh.off_coll[1] : 0,
                                  on1 ? sh.off_pcc[1] : 0, 0, gc);
            d_pc_eval_group<false>(D, sh, me0, me1, 0, on0 ? sh.n_bond[0] : 0, 0, on1 ? sh.n_bond[1] : 0, gc, 32);
            // PCC, thermostat, bonded HB partner, HB neighbours
            {
                int m0 = on0 ? sh.off_pcc[0] : 0, m0e = on0 ? (sh.hb_free[0] ? sh.off_hb[0] : sh.off_hbne[0]) : 0;
                int m1 = on1 ? sh.off_pcc[1] : 0, m1e = on1 ? (sh.hb_free[1] ? sh.off_hb[1] : sh.off_hbne[1]) : 0;
                d_pc_eval_group<false>(D, sh, me0, me1, m0, m0e, m1, m1e, gc + 32, 16);
                d_pc_eval_group<false>(D, sh, me0, me1, on0 ? sh.off_hbne[0] : 0, on0 ? sh.ncand[0] : 0, on1 ? sh.off_hbne[1] : 0,
                                       on1 ? sh.ncand[1] : 0, gc + 48, 16);
            }
            // free-HB scan: one warp blanks the section, then evaluates only the entries that passed the role filters
            if (tid >= gc + 64 && tid < gc + 96) {
                int lane = tid & 31;
                for (int a = 0; a < na; a++) {
                    bool on = (a == 0) ? on0 : on1;
                    if (!on || !sh.hb_free[a]) continue;
                    int b0 = a * PC_MAXC;
                    for (int p = sh.off_hb[a] + lane; p < sh.off_hbne[a]; p += 32) {
                        sh.ckind[b0 + p] = CK_NONE; sh.cflag[b0 + p] = 0; sh.ctime[b0 + p] = SD_INF; sh.cj[b0 + p] = 0;
                    }
                    __syncwarp();
                    if (lane < sh.n_pass[a]) {
                        int p = sh.off_hb[a] + D.cc_passidx[sh.cc_row[a] * 32 + lane];
                        d_pc_eval(D, sh, a, p, a ? me1 : me0);
                    }
                    __syncwarp();
                }
            }
        }
    }
    __syncthreads();
    long long pc2 = d_clock_after(&sh.natoms);
    if (tid == 0) D.ctl->phase_cycles[4] += pc2 - pc1;
    if (tid < 32) {
        for (int a = 0; a < na; a++) {
            int i = sh.atom[a];
            if (D.is_cons[i]) {
                if (tid == 0) { D.ev_time[i] = SD_INF; }
            } else if (sh.use_serial[a]) {
                if (tid == 0) {
                    SlotSrc S = d_live_slots(D);
                    Slot cur = d_reset_slot(D.ev_pot[i], D.ev_sub[i], D.ev_pcounter[i]);
                    int np = 0;
                    cur = d_predict_atom<false>(D, S, i, cur, sh.gauss[a], 0x7fffffff, &np);
                    d_store_slot(D, i, cur);
                    D.ctl->pair_predictions += np;
                }
            } else {
                d_pc_fold(D, sh, a);
                if (tid == 0) {
                    if (D.single_body) {   // wall / obstruction / sphere predictors come after every other one: a few more
                        Slot w = sh.result[a];   // JobAssigns on the folded slot
                        Self me = d_load_self(D, i);
                        if (D.wall_kind) d_wall_time(D, me, w);
                        if (D.n_obst) d_obst_time(D, me, w);
                        if (D.n_tunnel) d_tunnel_time(D, me, w);
                        if (D.flow_mark == 3) d_charge_time(D, me, w);
                        if (D.n_sphobst) d_sphobst_time(D, me, w);
                        sh.result[a] = w;
                    }
                    d_store_slot(D, i, sh.result[a]);
                    D.ctl->pair_predictions += sh.npair[a];
                    if (D.hydrogen_bond) {
                        D.hb_itype[i] = SDMD_HB_NOEVENT;                       // TimePrediction.c:325,329
                        int ns = sh.n_hbne[a];
                        if (ns > 0) {                                           // partnerNum, TimePrediction.c:504-511,574-577
                            int w = sh.winner_pos[a];
                            D.hbn_pnum[i] = (w >= sh.off_hbne[a]) ? (w - sh.off_hbne[a] + 1) : ns;
                        }
                    }
                }
            }
            __syncwarp();
        }
        if (tid == 0) D.ctl->phase_cycles[5] += clock64() - pc2;
    }
}

// UpdateCBTRun (CBT.c:56-83) with the root path's nodes, siblings and sibling times fetched by parallel lanes
// first; the two-phase walk itself (which decides where to stop) is replayed in order by lane 0.
__device__ inline void d_cbt_update_warp(const DevP &D, PredShared &sh, int a) {
    int lane = threadIdx.x & 31;
    int leaf = D.cbt_leaf[a];
    int depth = 0;
    for (int f = leaf / 2; f > 0; f /= 2) depth++;
    for (int k = lane; k < depth; k += 32) {
        int child = leaf >> k;          // node on the path at level k (k = 0: the leaf)
        int f = child >> 1;
        int sib = D.cbt_node[child ^ 1];
        sh.path_old[k] = D.cbt_node[f];
        sh.path_sib[k] = sib;
        sh.path_tsib[k] = d_rt(D, sib);
    }
    __syncwarp();
    if (lane == 0) {
        double ta = d_rt(D, a);
        int w = a;              // current winner coming up the path
        double tw = ta;
        bool phase1 = true;
        for (int k = 0; k < depth; k++) {
            int child = leaf >> k;
            int f = child >> 1;
            int old = sh.path_old[k];
            if (phase1 && old != a) phase1 = false;
            int sib = sh.path_sib[k];
            double ts = sh.path_tsib[k];
            // left child = node[2f], right = node[2f+1]; winner = t[left] < t[right] ? left : right
            bool w_is_left = (child & 1) == 0;
            int nw;
            double ntw;
            if (w_is_left) { if (tw < ts) { nw = w; ntw = tw; } else { nw = sib; ntw = ts; } }
            else { if (ts < tw) { nw = sib; ntw = ts; } else { nw = w; ntw = tw; } }
            D.cbt_node[f] = nw;
            if (!phase1 && nw == old) break;
            w = nw; tw = ntw;
        }
    }
    __syncwarp();
}

// ---------------------------------------------------------------------------------------------------
// the event loop's per-iteration pieces (SGThread.c:75-152), shared by every loop driver
__device__ inline void d_loop_head(const DevP &D, double t_until, i64 iter_limit) {
    Ctl *c = D.ctl;
    if (c->stop) return;
    if (c->error) { c->stop = 1 + SDMD_STOP_ERROR; return; }
    if (!(c->currenttime <= c->timestep)) { c->stop = 1 + SDMD_STOP_TIME_END; return; }
    if (c->iterations >= iter_limit) { c->stop = 1 + SDMD_STOP_MAX_ITER; return; }
    if (D.thermostat_on && c->g_filled - c->g_pos < 16) { c->stop = 1 + SDMD_STOP_NEED_RNG; return; }
    int target = D.cbt_node[1];
    c->target = target;
    c->dt = d_rt(D, target);
    int p = D.ev_partner[target];
    c->partner = p > 0 ? p : 0;
}

// publish the AssignJob result to the block (thread 0, before a barrier)
__device__ __forceinline__ void d_publish_job(const DevP &D, PredShared &sh) {
    Ctl *c = D.ctl;
    int target = c->target, partner = c->partner;
    sh.atom[0] = target;
    sh.atom[1] = partner;
    int type = D.ev_type[target];
    // candidate gathering may overlap the event unless the event itself edits what gathering reads:
    // the link list (PCC) or the HB lists / types (HB)
    sh.gathered = (type != SDMD_EV_PCC && type != SDMD_EV_HB);
}

// loop body after the advance, executed by the whole CTA: HazardCheck, DoEvent, counters (thread 0) overlapped
// with the candidate layout of the dirty atoms (warps 1-2), then the CTA-parallel Predict and the warp-assisted
// UpdateCBT (SGThread.c:80-151).  Precondition: d_publish_job + __syncthreads.
__device__ inline void d_loop_body_cta(const DevP &D, PredShared &sh, double t_until, sdmd_trace_rec *trace, i64 trace_cap) {
    Ctl *c = D.ctl;
    long long t0 = clock64();
    int spec = sh.gathered;
    if (threadIdx.x == 0) {
        int target = sh.atom[0], partner = sh.atom[1];
        c->currenttime += c->dt;
        int hazard = (partner > 0 && D.ev_pcounter[target] < D.ev_counter[partner]) ? -1 : 0;  // SGThread.c:224-227
        if (trace && c->trace_pos < trace_cap) {
            sdmd_trace_rec r;
            r.frame = c->frame; r.target = target; r.partner = D.ev_partner[target];
            r.type = D.ev_type[target]; r.sub = D.ev_sub[target]; r.executed = !hazard; r.pad_ = 0; r.dt = c->dt;
            trace[c->trace_pos] = r;
        }
        c->trace_pos++;
        c->iterations++;
        int na;
        if (!hazard) {
            c->frame++;
            d_do_event(D, target, partner);
            D.ev_counter[target]++;
            if (partner && !D.is_cons[partner]) D.ev_counter[partner]++;
            na = partner ? 2 : 1;
            if (c->currenttime >= t_until) c->stop = 1 + SDMD_STOP_T_UNTIL;
        } else {
            na = 1;
            c->count_recal++;
        }
        sh.natoms = na;
        for (int a = 0; a < na; a++) {
            int i = sh.atom[a];
            // ResetTarget for every listed atom first (ToolFunctions.c:504-514)
            D.ev_type[i] = SDMD_EV_INVD; D.ev_time[i] = SD_INF; D.ev_partner[i] = SDMD_INVALID;
            // one thermostat draw per predicted atom, in list order
            sh.gauss[a] = 0;
            if (D.thermostat_on && !D.is_cons[i]) sh.gauss[a] = d_next_gauss(D);
        }
        c->phase_cycles[8] += clock64() - t0;
    } else if (spec) {
        d_predict_gather(D, sh, 1, sh.atom[1] > 0 ? 2 : 1);
    }
    __syncthreads();
    long long t1 = d_clock_after(&sh.natoms);
    if (!spec) {
        d_predict_gather(D, sh, 1, sh.natoms);
        __syncthreads();
    }
    long long t1b = d_clock_after(&sh.natoms);
    d_predict_cta(D, sh);
    long long t2 = clock64();
    if (threadIdx.x < 32) {
        d_cbt_update_warp(D, sh, sh.atom[0]);
        if (sh.natoms > 1) d_cbt_update_warp(D, sh, sh.atom[1]);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        c->phase_cycles[2] += t1 - t0; c->phase_cycles[3] += t1b - t1; c->phase_cycles[6] += d_clock_after(&sh.natoms) - t2; c->phase_cycles[7] += 1;
    }
}

