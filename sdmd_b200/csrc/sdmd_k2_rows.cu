// sdmd_k2_rows.cu — round 0 of the full-system Predict (SGThread.c:47-52 -> Predict, SGThread.c:158-211 -> CollisionTime,
// TimePrediction.c:44-183) for systems without bonds / hydrogen bonds / single-body sources / constrained atoms (the GEL
// boxes of BASELINE config 5), rebuilt around memory-level parallelism.  Both round-0 kernels of sdmd_cuda.cu were bound by
// chains of dependent loads (ncu: long scoreboard + wait ~ 7 of 9 stall cycles per issue), not by DRAM or instruction count:
//
//   k_r0_filter_rows   thread per cell-sorted atom.  The cell-sorted arrays keep an x-row of cells contiguous, so the 27-cell
//                      neighbourhood is 9 rows of three cells = 9 (nearly always contiguous) position ranges.  All bounds of a
//                      row are loaded side by side, then the row's candidates are fetched EIGHT AT A TIME with predicated,
//                      mutually independent 16-byte loads and tested in single precision (superset sphere, cell-local
//                      coordinates: see d_r0_filter in sdmd_dev_body.cuh).  One load round trip per row instead of one per
//                      cell bound + one per candidate.  Survivors go to the HBM row `position | scan slot << 27` in ROW order,
//                      i.e. NOT in the reference's visiting order.  (Marking hits in a per-row bit mask and writing them out after the
//                      row — the store path once per hit instead of predicated at every candidate — was measured slower: 106 vs 99 us.)
//   k_r0_eval_rows     thread per cell-sorted atom.  Survivor records (64 bytes each) are gathered FOUR AT A TIME by cp.async
//                      (LDGSTS) into a thread-private strip of shared memory — no registers held across the gather, one L2 / DRAM
//                      round trip per four survivors instead of one each — then evaluated in double precision with the
//                      reference's arithmetic.  Everything read by atom id (stale slot fields, the partner's event counter)
//                      is loaded once, after the fold, and only where the winner needs it.
//
// Order independence.  JobAssign (TimePrediction.c:1140-1158) replaces only on strict `<`: the sequential result is the
// minimum time with ties going to the candidate visited FIRST.  A survivor word is exactly the visiting rank (scan slot of
// the neighbour cell, DMD.c:118-120, then position inside the cell = link-list order of a fresh LinkList("full")), so the
// fold takes the lexicographic minimum of (time, word); the runner-up the resolve stage wants is the second smallest time,
// which no order changes.  Re-evaluation rounds (CheckDuplication in visiting order) sort the row first
// (d_collision_time_from_list, DevP::rows_mode).
#include <cstdio>
#include <cstdlib>
#include "sdmd_host.h"

#ifndef RF_NT
#define RF_NT 256
#endif
#ifndef RF_MINB
#define RF_MINB 4
#endif
#ifndef RF_B
#define RF_B 2            // candidates fetched per round trip (measured at 10^6 atoms, 8 A lattice: 8 -> 161 us, 6 -> 139, 4 -> 105, 2 -> 99; 80^3 at 5 A: 4 -> 166, 2 -> 161; cell-by-cell filter: 140)
#endif
#ifndef RE_NT
#define RE_NT 128
#endif
#ifndef RE_MINB
#define RE_MINB 5
#endif
#ifndef RE_NB
#define RE_NB 4           // survivor records gathered per round trip (RE_NB * 64 B of shared memory per thread)
#endif

// Streaming L2 prefetch.  Both kernels walk the cell-sorted arrays front to back, one atom per thread, and every warp starts
// with a chain of first-touch DRAM loads (ncu: ~45 % of the evaluation kernel's stall samples sit on its first three loads, with
// only ~18 warps per SM to hide them).  Each thread therefore asks L2 for the lines the thread PF atoms further on will open with;
// PF is about twice the atoms in flight on the whole GPU.  The neighbour records an atom gathers lie within one cell plane
// (~n^(2/3) positions) of it, so the same stream makes those gathers L2 hits.
#ifndef RF_PF
#define RF_PF 0
#endif
#ifndef RE_PF
#define RE_PF 0
#endif
__device__ __forceinline__ void r0_prefetch_l2(const void *p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

__global__ void __launch_bounds__(RF_NT, RF_MINB) k_r0_filter_rows(DevP D, int *pcount, int *zero, int nzero) {
    SDMD_PDL_ENTRY();
    const int p_self = blockIdx.x * RF_NT + threadIdx.x;
    if (p_self < nzero) zero[p_self] = 0;
    if (p_self == 0) pcount[0] = 0;
    if (p_self >= D.n) return;
    pcount[p_self + 1] = 0;             // (indexed by atom id 1..n; the evaluation kernel counts into it)
#if RF_PF
    if (p_self + RF_PF < D.n) {
        r0_prefetch_l2(&D.sloc[p_self + RF_PF]);
        if ((threadIdx.x & 7) == 0) r0_prefetch_l2(&D.scell[p_self + RF_PF]);
    }
#endif
    const float4 me = __ldg(&D.sloc[p_self]);
    const int c = __ldg(&D.scell[p_self]);
    const int nx = D.cellnum[0], ny = D.cellnum[1], nz = D.cellnum[2], nxy = nx * ny;
    const int cz = c / nxy, rem = c - cz * nxy, cy = rem / nx, cx = rem - cy * nx;
    const bool wm = cx == 0, wp = cx == nx - 1;                 // the row wraps on my -x / +x side: that cell is not adjacent in memory
    const int xm = wm ? nx - 1 : cx - 1, xp = wp ? 0 : cx + 1;
    const float csx = D.cellsize_f[0], csy = D.cellsize_f[1], csz = D.cellsize_f[2];
    const float mxA = me.x + csx, mxC = me.x - csx;             // my position as seen from the -x / +x neighbour cell's corner
    const float cut2 = D.cut2_f32;
    const int cap = D.surv_cap;
    unsigned int *out = D.surv + (size_t)p_self * cap;
    int ns = 0, ncand = 0;
#pragma unroll
    for (int r = 0; r < 9; r++) {
        const int dy = r % 3 - 1, dz = r / 3 - 1;
        int yy = cy + dy, zz = cz + dz;
        if (dy < 0 && yy < 0) yy = ny - 1;
        if (dy > 0 && yy == ny) yy = 0;
        if (dz < 0 && zz < 0) zz = nz - 1;
        if (dz > 0 && zz == nz) zz = 0;
        const float my = dy == 0 ? me.y : (dy > 0 ? me.y - csy : me.y + csy);
        const float mz = dz == 0 ? me.z : (dz > 0 ? me.z - csz : me.z + csz);
        const int *cs = D.cell_start + (zz * nxy + yy * nx);
        const int sB = __ldg(cs + cx), eB = __ldg(cs + cx + 1);
        const int sA = __ldg(cs + xm), eC = __ldg(cs + xp + 1);
        int eA = sB, sC = eB;
        if (wm) eA = __ldg(cs + xm + 1);
        if (wp) sC = __ldg(cs + xp);
        const int t1 = eA - sA, t2 = t1 + (eB - sB), tot = t2 + (eC - sC);
        const int offB = sB - t1, offC = sC - t2;
        ncand += tot;
        // scan slot (0..26) of the neighbour-cell offset (dx, dy, dz) at [(dz+1)*9 + (dy+1)*3 + dx+1]: inverse of the DMD.c:118-120 tables
        constexpr unsigned KINV[27] = {14, 11, 23, 12, 10, 17, 22, 16, 24, 13, 9, 19, 8, 0, 1, 20, 2, 6, 21, 15, 25, 18, 3, 5, 26, 4, 7};
        const unsigned kA = KINV[r * 3] << 27, kB = KINV[r * 3 + 1] << 27, kC = KINV[r * 3 + 2] << 27;
        for (int base = 0; base < tot; base += RF_B) {
            float4 q[RF_B];
            int pp[RF_B];
#pragma unroll
            for (int u = 0; u < RF_B; u++) {
                const int s = base + u;
                pp[u] = s + (s < t1 ? sA : (s < t2 ? offB : offC));
                if (s < tot) q[u] = __ldg(&D.sloc[pp[u]]);
            }
#pragma unroll
            for (int u = 0; u < RF_B; u++) {
                const int s = base + u;
                if (s < tot) {
                    const float mx = s < t1 ? mxA : (s < t2 ? me.x : mxC);
                    const float rx = mx - q[u].x, ry = my - q[u].y, rz = mz - q[u].z;
                    const float r2 = fmaf(rz, rz, fmaf(ry, ry, rx * rx));
                    bool hit = r2 <= cut2;
                    if (r == 4) hit = hit && pp[u] != p_self;
                    if (hit) {
                        const unsigned k = s < t1 ? kA : (s < t2 ? kB : kC);
                        if (ns < cap) out[ns] = (unsigned)pp[u] | k;
                        ns++;
                    }
                }
            }
        }
    }
    D.surv_n[p_self] = (ns <= cap) ? ns : -1;
    D.surv_np[p_self] = ncand - 1;      // every candidate but myself is a counted pair evaluation (TimePrediction.c:105): no skip filter applies here
}

__device__ __forceinline__ unsigned r0_smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void r0_cp_async16(unsigned dst, const void *src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void r0_cp_async_wait_all() {
    asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
}

#define R0_NOWORD 0xffffffffu
#ifndef RE_COOP
#define RE_COOP 0
#endif
// One gather = the records of RE_NB survivors per atom, copied by cp.async into the atoms' strips.
// RE_COOP 0 (shipped): every lane copies its own records, four 16-byte parts each.
// RE_COOP 1: the warp copies together — four neighbouring lanes take the four parts of ONE record, so a copy instruction touches
//   8 cache lines instead of 32 (lane l copies, in step t, part l % 4 of entry t / 4 of the atom of lane (t % 4) * 8 + l / 4).
//   Tried because L1TEX wavefronts are the busiest unit of the kernel (ncu: 48 %); measured 127.9 us against 122.9 us: the
//   shuffles and the two warp barriers per batch cost more than the wavefronts saved.
__device__ __forceinline__ void r0_gather_batch(const DevP &D, const unsigned ww[4], unsigned sbase_self, int lane) {
#if RE_COOP
    const int part = lane & 3;
#pragma unroll
    for (int t = 0; t < 4 * RE_NB; t++) {
        const int e = t >> 2, owner = ((t & 3) << 3) + (lane >> 2);
        const unsigned w = __shfl_sync(0xffffffffu, ww[e], owner);
        if (w != R0_NOWORD)
            r0_cp_async16(sbase_self + (unsigned)((owner - lane) * 16 + (e * 4 + part) * RE_NT * 16),
                          reinterpret_cast<const char *>(&D.srec[w & 0x7ffffffu]) + 16 * part);
    }
#else
#pragma unroll
    for (int e = 0; e < RE_NB; e++)
        if (ww[e] != R0_NOWORD) {
            const char *src = reinterpret_cast<const char *>(&D.srec[ww[e] & 0x7ffffffu]);
#pragma unroll
            for (int q = 0; q < 4; q++) r0_cp_async16(sbase_self + (unsigned)((e * 4 + q) * RE_NT * 16), src + 16 * q);
        }
#endif
}

// FAST: every atom has the same type and its collision table has at most two steps (DevP::uni_table; the GEL boxes: one type,
// hard core + one step).  The table then lives in registers and FindPair (ToolFunctions.c:104-224) + the branch of CollisionTime
// (TimePrediction.c:121-164) become selects: no table loads on the path of a candidate.
#ifndef RE_ILP
#define RE_ILP 1
#endif
template <bool FAST>
__global__ void __launch_bounds__(RE_NT, RE_MINB) k_r0_eval_rows(DevP D, SlotBuf dst, Round0Out out, i64 g_base) {
    // thread-private strips: entry e, 16-byte part q of thread t at [(e * 4 + q) * RE_NT + t] (conflict-free 16-byte accesses)
    __shared__ double2 strip[RE_NB * 4 * RE_NT];
    SDMD_PDL_ENTRY();
    const int tid = threadIdx.x, lane = tid & 31;
    const int p = blockIdx.x * RE_NT + tid;
    const bool live = p < D.n;                       // (every lane of a warp takes part in the gathers)
#if RE_PF
    if (p + RE_PF < D.n) {
        const char *rec = reinterpret_cast<const char *>(&D.srec[p + RE_PF]);
        r0_prefetch_l2(rec); r0_prefetch_l2(rec + 32);
        const char *rw = reinterpret_cast<const char *>(D.surv + (size_t)(p + RE_PF) * D.surv_cap);
        r0_prefetch_l2(rw); r0_prefetch_l2(rw + 32);
        if ((tid & 7) == 0) { r0_prefetch_l2(&D.surv_n[p + RE_PF]); r0_prefetch_l2(&D.surv_np[p + RE_PF]); }
    }
#endif
    const int pc = live ? p : D.n - 1;
    const uint4 *row = reinterpret_cast<const uint4 *>(D.surv + (size_t)pc * D.surv_cap);
    const int ns_raw = __ldg(&D.surv_n[pc]);
    const int ns = (live && ns_raw > 0) ? ns_raw : 0;
    uint4 w4 = __ldg(row);          // (entries behind ns are stale words of earlier passes: masked below)
    const unsigned sbase = r0_smem_u32(&strip[tid]);
#if RE_COOP
    const int nmax = __reduce_max_sync(0xffffffffu, ns);    // warp-uniform trip count: every lane takes part in every gather
#else
    const int nmax = ns;
#endif
    Self me;
    {
        const double2 *rp = reinterpret_cast<const double2 *>(&D.srec[pc]);
        const double2 r0 = __ldg(rp), r1 = __ldg(rp + 1), r2 = __ldg(rp + 2);
        const int4 q = __ldg(reinterpret_cast<const int4 *>(rp + 3));
        me.i = q.y;
        me.x = r0.x; me.y = r0.y; me.z = r1.x; me.vx = r1.y; me.vy = r2.x; me.vz = r2.y;
        me.cx = (int)((unsigned)q.z & 0xffffu); me.cy = (int)((unsigned)q.z >> 16); me.cz = q.w;
        me.counter = 0; me.npair = 0; me.type = q.x; me.cm0 = 0; me.cm1 = 0;
    }
    double tb_dmin2 = 0, tb_d0 = 0, tb_d1 = 0, tb_e0 = 0, tb_e1 = 0;
    int tb_ns = 0;
    if (FAST) {
        const sdmd_step_table *tb = &D.tables[D.uni_table];
        tb_dmin2 = __ldg(&tb->dmin2); tb_ns = __ldg(&tb->nstep);
        tb_d0 = __ldg(&tb->d2[0]); tb_e0 = __ldg(&tb->e[0]); tb_d1 = __ldg(&tb->d2[1]); tb_e1 = __ldg(&tb->e[1]);
    }
    const int np = __ldg(&D.surv_np[pc]);
    // every load above is independent of the others and in flight by now; the first gather needs the row's words
    if (nmax > 0) {
        const unsigned ww[4] = {0 < ns ? w4.x : R0_NOWORD, 1 < ns ? w4.y : R0_NOWORD, 2 < ns ? w4.z : R0_NOWORD, 3 < ns ? w4.w : R0_NOWORD};
        r0_gather_batch(D, ww, sbase, lane);
    }
    const int i = me.i;
    double g = 0;
    if (D.thermostat_on) { i64 gi = g_base + i - 1; g = __ldg(&D.gauss[gi >= D.g_cap ? gi - D.g_cap : gi]); }
    // partnerCounter survives ResetTarget (ToolFunctions.c:504-514): the stale value stays in the slot of a partner-less winner.
    // Loads by atom id are issued as early as their address is known and consumed after the fold.
    const i64 pc_self = D.ev_pcounter[i];
    i64 pc1 = 0;
    const bool seam = me.cx == 0 || me.cy == 0 || me.cz == 0 || me.cx == D.cellnum[0] - 1 || me.cy == D.cellnum[1] - 1 ||
                      me.cz == D.cellnum[2] - 1;
    // fold state: lexicographic minimum of (time, visiting rank) and the second smallest time
    double t1 = SD_INF, m2 = SD_INF, pot1 = 0;
    unsigned w1 = 0xffffffffu;
    int sub1 = 0, j1 = 0;
    for (int s0 = 0; s0 < nmax; s0 += RE_NB) {
        const uint4 wc = w4;
        const int nb = ns - s0 < RE_NB ? ns - s0 : RE_NB;      // may be <= 0: this lane is done, it only helps to copy
        if (s0 + RE_NB < ns) w4 = __ldg(row + (s0 + RE_NB) / 4);     // next batch's words travel under this batch's work
        r0_cp_async_wait_all();
#if RE_COOP
        __syncwarp();                                  // the parts of my records were copied by three other lanes
#endif
        const unsigned ww[4] = {wc.x, wc.y, wc.z, wc.w};
        if (FAST) {
#pragma unroll
            for (int e0 = 0; e0 < RE_NB; e0 += RE_ILP) {
                if (e0 < nb) {
                    double T[RE_ILP], POT[RE_ILP], SX[RE_ILP], SY[RE_ILP], SZ[RE_ILP];
                    int SUB[RE_ILP], J[RE_ILP];
                    bool OK[RE_ILP], CORE[RE_ILP];
#pragma unroll
                    for (int u = 0; u < RE_ILP; u++) { SX[u] = 0; SY[u] = 0; SZ[u] = 0; }
                    if (seam) {     // PBCShift via the scanned neighbour cell; only atoms in a boundary cell can need one
#pragma unroll
                        for (int u = 0; u < RE_ILP; u++) {
                            int ncx, ncy, ncz;
                            d_neighbor_cell(D, me.cx, me.cy, me.cz, (int)(ww[e0 + u] >> 27), &ncx, &ncy, &ncz);
                            SX[u] = d_shift1(me.cx, ncx, D.cellnum[0], D.box[0]);
                            SY[u] = d_shift1(me.cy, ncy, D.cellnum[1], D.box[1]);
                            SZ[u] = d_shift1(me.cz, ncz, D.cellnum[2], D.box[2]);
                        }
                    }
#pragma unroll
                    for (int u = 0; u < RE_ILP; u++) {
                        const int e = e0 + u;
                        const double2 c0 = strip[(e * 4 + 0) * RE_NT + tid], c1 = strip[(e * 4 + 1) * RE_NT + tid], c2 = strip[(e * 4 + 2) * RE_NT + tid];
                        const int4 tq = *reinterpret_cast<const int4 *>(&strip[(e * 4 + 3) * RE_NT + tid]);
                        J[u] = tq.y;
                        const Pair pr = d_pair(me.x, me.y, me.z, me.vx, me.vy, me.vz, c0.x, c0.y, c1.x, c1.y, c2.x, c2.y, SX[u], SY[u], SZ[u]);
                        const bool in = e < nb && pr.r2 <= D.cut2;                         // TimePrediction.c:107
                        const double tol = (pr.b < 0) ? SD_ZERO : -SD_ZERO;               // ToolFunctions.c:46-50: dir * ZERO
                        CORE[u] = in && pr.r2 < tb_dmin2 - SD_ZERO;                        // :119-139 (warning, no event)
                        const bool s0c = pr.r2 < tb_d0 + tol, s1c = pr.r2 < tb_d1 + tol;
                        const int kk = (tb_ns > 0 && s0c) ? 0 : ((tb_ns > 1 && s1c) ? 1 : tb_ns);   // first step with dist2 < d + tol, :144-169
                        const double lower = kk == 0 ? tb_dmin2 : (kk == 1 ? tb_d0 : tb_d1);
                        const double lowPot = kk == 0 ? 0.0 : (kk == 1 ? tb_e0 : tb_e1);
                        const bool has_up = kk < tb_ns;
                        const double upper = has_up ? (kk == 0 ? tb_d0 : tb_d1) : SD_INF;        // collision tables have no dmax2: :212-221
                        const double upPot = has_up ? -1 * (kk == 0 ? tb_e0 : tb_e1) : SD_INF;
                        const int cs = has_up ? (kk == 0 ? 0 : -1) : 1;
                        const bool none = lowPot == 0 && upPot == 0;                                 // TimePrediction.c:121-124
                        const bool inner = pr.b < 0 && d_disc(pr, lower) > 0;                       // :130-141
                        const double d2 = inner ? lower : upper;
                        POT[u] = inner ? lowPot : upPot;
                        SUB[u] = inner ? (cs == 0 ? SDMD_SUB_CORECOLLI : SDMD_SUB_TBDCAPTR) : (D.method_ding ? SDMD_SUB_TBDESCAP : SDMD_SUB_INVALID);
                        const bool skip = !inner && D.method_ding && cs == 1;                      // :152-155
                        T[u] = d_time(pr, inner ? -1 : 1, d2);
                        OK[u] = in && !CORE[u] && !none && !skip;
                    }
#pragma unroll
                    for (int u = 0; u < RE_ILP; u++) {
                        if (CORE[u]) atomicAdd((unsigned long long *)&D.ctl->warnings, 1ULL);
                        if (OK[u]) {
                            const double t = T[u];
                            const unsigned w = ww[e0 + u];
                            if (t < 0) atomicAdd((unsigned long long *)&D.ctl->error_prints, 1ULL);   // :168-173 prints and carries on
                            if (t > 0) {
                                if (t < t1 || (t == t1 && w < w1)) { m2 = t1; t1 = t; w1 = w; pot1 = POT[u]; sub1 = SUB[u]; j1 = J[u]; pc1 = D.ev_counter[J[u]]; }
                                else if (t < m2) m2 = t;
                            }
                        }
                    }
                }
            }
        } else {
#pragma unroll
            for (int e = 0; e < RE_NB; e++) {
                if (e < nb) {
                    const double2 c0 = strip[(e * 4 + 0) * RE_NT + tid], c1 = strip[(e * 4 + 1) * RE_NT + tid], c2 = strip[(e * 4 + 2) * RE_NT + tid];
                    const int4 tq = *reinterpret_cast<const int4 *>(&strip[(e * 4 + 3) * RE_NT + tid]);
                    const unsigned w = ww[e];
                    double shx = 0, shy = 0, shz = 0;
                    if (seam) {
                        int ncx, ncy, ncz;
                        d_neighbor_cell(D, me.cx, me.cy, me.cz, (int)(w >> 27), &ncx, &ncy, &ncz);
                        shx = d_shift1(me.cx, ncx, D.cellnum[0], D.box[0]);
                        shy = d_shift1(me.cy, ncy, D.cellnum[1], D.box[1]);
                        shz = d_shift1(me.cz, ncz, D.cellnum[2], D.box[2]);
                    }
                    double t, pot;
                    int sub;
                    if (d_eval_pair(D, me, tq.y, tq.x, c0.x, c0.y, c1.x, c1.y, c2.x, c2.y, shx, shy, shz, 0, me.type, &t, &pot, &sub) && t > 0) {
                        if (t < t1 || (t == t1 && w < w1)) { m2 = t1; t1 = t; w1 = w; pot1 = pot; sub1 = sub; j1 = tq.y; pc1 = D.ev_counter[tq.y]; }
                        else if (t < m2) m2 = t;
                    }
                }
            }
        }
        if (s0 + RE_NB < nmax) {
#if RE_COOP
            __syncwarp();           // every lane has read its records of this batch: the strips are free for the next gather
#endif
            const int left = ns - (s0 + RE_NB);
            const unsigned wn[4] = {0 < left ? w4.x : R0_NOWORD, 1 < left ? w4.y : R0_NOWORD, 2 < left ? w4.z : R0_NOWORD, 3 < left ? w4.w : R0_NOWORD};
            r0_gather_batch(D, wn, sbase, lane);
        }
    }
    if (!live) return;
    if (ns_raw < 0) {   // the row overflowed: exact in-kernel re-scan of the 27 cells (generic round-0 path)
        Slot cur = d_reset_slot(D.ev_pot[i], D.ev_sub[i], D.ev_pcounter[i]);
        int npf = 0;
        double m2f = SD_INF;
        PairBest pb;
        pb.bp_time = SD_INF; pb.sp_time = SD_INF; pb.bp_j = 0;
        SlotSrc none = {};
        cur = d_predict_atom<true, true, true>(D, none, i, cur, g, 0, &npf, &m2f, &pb, p);
        dst.time[i] = cur.time; dst.pot[i] = cur.pot; dst.partner[i] = cur.partner; dst.pcounter[i] = cur.pcounter;
        dst.type[i] = (uint8_t)cur.type; dst.sub[i] = (uint8_t)cur.sub;
        out.npair[i] = npf;
        out.second[i] = cur.partner > 0 ? m2f : pb.sp_time;
        out.bp_time[i] = pb.bp_time;
        out.bp_j[i] = pb.bp_j;
        if (cur.partner > i) { atomicAdd(&out.pcount[cur.partner], 1); out.pwho[cur.partner] = i; }
        return;
    }
    // the collision stage as JobAssign would have left it (all candidates so far are pair candidates)
    Slot cur;
    cur.time = t1; cur.m2 = m2; cur.bp_time = t1; cur.sp_time = m2; cur.bp_j = j1;
    cur.pot = pot1; cur.sub = sub1; cur.pcounter = 0;
    cur.type = j1 ? SDMD_EV_COLI : SDMD_EV_INVD;
    cur.partner = j1 ? j1 : SDMD_INVALID;
    const double bp_time = t1, sp_time = m2;
    d_pcc_time(D, me, cur);
    if (D.thermostat_on) d_thermostat_time(D, cur, g);
    const i64 pcounter = cur.partner > 0 ? pc1 : pc_self;
    if (cur.type == SDMD_EV_INVD) { cur.pot = D.ev_pot[i]; cur.sub = D.ev_sub[i]; }   // ResetTarget leaves them untouched
    dst.time[i] = cur.time; dst.pot[i] = cur.pot; dst.partner[i] = cur.partner; dst.pcounter[i] = pcounter;
    dst.type[i] = (uint8_t)cur.type; dst.sub[i] = (uint8_t)cur.sub;
    out.npair[i] = np;
    out.second[i] = cur.partner > 0 ? cur.m2 : sp_time;
    out.bp_time[i] = bp_time;
    out.bp_j[i] = j1;
    if (cur.partner > i) { atomicAdd(&out.pcount[cur.partner], 1); out.pwho[cur.partner] = i; }
}

int sdmd_k2_rows_configure(DevP &P) {
    P.rows_mode = 0;
    const char *env = getenv("SDMD_K2_ROWS");      // SDMD_K2_ROWS=0: the cell-by-cell filter + ordered evaluation of sdmd_cuda.cu (kept, parity-tested)
    if (env && env[0] == '0') return 1;
    if (!(P.gel_mode && !P.hydrogen_bond && !P.single_body && P.no_cons && P.cm_empty)) return 1;
    if (P.n >= (1 << 27) || !P.sloc || P.tile_mode) return 1;
    if (P.surv_cap % 4) return 1;
    P.rows_mode = 1;
    return 0;
}

int sdmd_k2_rows_filter_launch(const DevP &P, int *pcount, int *zero, int nzero, cudaStream_t stream) {
    if (nzero > P.n) return -1;
    return sdmd_launch_pdl(k_r0_filter_rows, dim3((P.n + RF_NT - 1) / RF_NT), dim3(RF_NT), stream, P, pcount, zero, nzero) == cudaSuccess ? 0 : -1;
}

int sdmd_k2_rows_eval_launch(const DevP &P, SlotBuf dst, Round0Out out, long long g_base, cudaStream_t stream) {
    const dim3 grid((P.n + RE_NT - 1) / RE_NT), block(RE_NT);
    const i64 gb = g_base;
    cudaError_t e = P.uni_table >= 0 ? sdmd_launch_pdl(k_r0_eval_rows<true>, grid, block, stream, P, dst, out, gb)
                                     : sdmd_launch_pdl(k_r0_eval_rows<false>, grid, block, stream, P, dst, out, gb);
    return e == cudaSuccess ? 0 : -1;
}
