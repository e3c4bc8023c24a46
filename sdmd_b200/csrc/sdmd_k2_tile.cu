// sdmd_k2_tile.cu — round 0 of the full-system Predict (SGThread.c:47-52 -> Predict, SGThread.c:158-211 -> CollisionTime,
// TimePrediction.c:44-183) as ONE kernel over cell tiles, for systems without bonds / hydrogen bonds / single-body sources
// (the GEL boxes of BASELINE config 5).
//
// One thread block owns a box of TX x TY x TZ cells.  The cell-sorted arrays keep every x-row of cells contiguous, so the
// block's halo — (TX+2) x (TY+2) x (TZ+2) cells — is (TY+2)(TZ+2) contiguous runs (a run splits where the row wraps around
// the periodic box).  Each run is brought into shared memory by the TMA engine (1-D cp.async.bulk + mbarrier, UBLKCP in
// SASS): the 64-byte records (x, y, z, vx, vy, vz, type, atom, cell) and the 16-byte single-precision cell-local positions.
// Nothing else is read from HBM for the pair work.
//
//   phase A  thread per own atom: walk the 9 x-rows of its 27-cell neighbourhood (3 cells of a row are one contiguous range
//            of the staged arrays) with a single-precision superset test on tile-relative coordinates; candidates inside the
//            superset sphere are appended to a per-warp pool by ballot (no atomics, no divergence on the expensive branch);
//   phase B  when the pool fills (and at the end) the warp evaluates its entries, one per lane: exact double-precision
//            distance re-test with the reference's arithmetic, step-table walk, discriminant, root; events are linked into
//            the owner's list; each owner folds its list;
//   phase C  thread per own atom: PCC, thermostat, slot / runner-up / pointer outputs — same outputs as k_predict_all_round.
//
// The fold is order independent: JobAssign (TimePrediction.c:1140-1158) replaces only on strict `<`, so the result is the
// minimum time with ties going to the candidate visited first; every pool entry carries its visiting rank (scan slot of the
// neighbour cell, DMD.c:118-120, then position inside the cell = link-list order of a fresh LinkList("full")), and the fold
// takes the lexicographic minimum of (time, rank).  The runner-up values the resolve stage needs are the second smallest
// time, again independent of order.
#include <cstdio>
#include "sdmd_host.h"

#define TILE_NT 256
#define TILE_WARPS (TILE_NT / 32)
#define TILE_LCAP 8           // private hit-list entries per lane between two evaluation rounds
#define TILE_ECAP 96          // event entries per warp between two folds
#ifndef TILE_MINBLOCKS
#define TILE_MINBLOCKS 2
#endif

struct WarpPool {
    unsigned priv[TILE_LCAP * 32];            // [k][lane]: staged index of the k-th hit of the lane's atom | x-row (0..8) << 16
    unsigned ent[TILE_LCAP * 32];             // the same, compacted over the warp: | owner lane << 20
    double time[TILE_ECAP], pot[TILE_ECAP];   // event entries: predicted time, potential step
    unsigned short key[TILE_ECAP];            // visiting rank: scan slot << 11 | position inside the cell
    unsigned short eq[TILE_ECAP];             // staged index of the partner
    short next[TILE_ECAP];
    unsigned char sub[TILE_ECAP];
    int head[32];                             // per owner lane: its list of event entries
    int ownp[32], ownh[32];                   // per owner lane: staged index of the own atom, halo cell (hx-1, hy-1, hz-1)
};

// scan slot (0..26) of neighbour cell offset (dx,dy,dz), inverse of the DMD.c:118-120 tables
static __constant__ unsigned char c_tile_kinv[27] = {14, 11, 23, 12, 10, 17, 22, 16, 24, 13, 9, 19, 8, 0, 1, 20, 2, 6, 21, 15, 25, 18, 3, 5, 26, 4, 7};

__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long *bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned parity) {
    unsigned ok;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    } while (!ok);
}
// 1-D bulk copy global -> shared through the TMA engine; completion is counted in bytes on the mbarrier
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, unsigned bytes, unsigned long long *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ Self d_self_from_rec(const DevP &D, const SRec &r) {
    Self me;
    me.i = r.atom;
    me.x = r.x; me.y = r.y; me.z = r.z; me.vx = r.vx; me.vy = r.vy; me.vz = r.vz;
    me.cx = (int)(r.cxy & 0xffffu); me.cy = (int)(r.cxy >> 16); me.cz = (int)r.cz;
    me.counter = 0;
    me.npair = 0;
    me.type = r.type;
    me.cm0 = 0; me.cm1 = 0;
    return me;
}
// PBCShift (ToolFunctions.c:801-817) along one axis from the two cell indices, without branches (cellnum >= 3, so either
// condition already implies ci != cj)
__device__ __forceinline__ double d_shift1_sel(int ci, int cj, int cn, double box) {
    double s = 0;
    s = (ci == 0 && cj == cn - 1) ? -1 * box : s;
    s = (ci == cn - 1 && cj == 0) ? box : s;
    return s;
}

// slot, runner-up records and "who points at whom" of one atom: what k_predict_all_round writes in round 0
__device__ __forceinline__ void d_round0_store(const DevP &D, const SlotBuf &dst, const Round0Out &o, int i, const Slot &cur,
                                               const PairBest &pb, int np) {
    dst.time[i] = cur.time; dst.pot[i] = cur.pot; dst.partner[i] = cur.partner; dst.pcounter[i] = cur.pcounter;
    dst.type[i] = (uint8_t)cur.type; dst.sub[i] = (uint8_t)cur.sub;
    o.npair[i] = np;
    bool pairwin = cur.partner > 0;
    o.second[i] = pairwin ? cur.m2 : pb.sp_time;
    o.bp_time[i] = pb.bp_time;
    o.bp_j[i] = pb.bp_j;
    if (cur.partner > i) { atomicAdd(&o.pcount[cur.partner], 1); o.pwho[cur.partner] = i; }
}

struct Fold {
    double t, m2, pot;
    int key, q, sub;
};
// every owner folds its list: lexicographic minimum of (time, visiting rank) and the second smallest time
__device__ __forceinline__ void d_pool_fold(WarpPool *wp, int lane, Fold &f) {
    for (int e = wp->head[lane]; e >= 0; e = wp->next[e]) {
        double t = wp->time[e];
        int key = wp->key[e];
        if (t < f.t || (t == f.t && key < f.key)) {
            f.m2 = f.t;
            f.t = t; f.key = key; f.pot = wp->pot[e]; f.sub = wp->sub[e]; f.q = wp->eq[e];
        } else if (t < f.m2) f.m2 = t;
    }
    wp->head[lane] = -1;
}
// phase B: the warp's private hit lists are compacted (so every lane gets an entry whatever its own atom found), each entry
// is evaluated from its owner's side — exact re-test and event prediction, CollisionTime, TimePrediction.c:105-175 — events
// are linked into their owners' lists together with their visiting rank, and every owner folds its list
__device__ __forceinline__ void d_pool_flush(const DevP &D, WarpPool *wp, const SRec *rec, const int *lbeg, int HX, int HXY, bool seam,
                                             int lane, int cnt, Fold &f) {
    const unsigned lt_mask = (1u << lane) - 1u;
    int inc = cnt;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        int t = __shfl_up_sync(0xffffffffu, inc, off);
        if (lane >= off) inc += t;
    }
    const int T = __shfl_sync(0xffffffffu, inc, 31);
    const int base = inc - cnt;
    const int maxcnt = __reduce_max_sync(0xffffffffu, cnt);
    for (int k = 0; k < maxcnt; k++)
        if (k < cnt) wp->ent[base + k] = wp->priv[k * 32 + lane] | ((unsigned)lane << 20);
    __syncwarp();
    int nev = 0;
    for (int e0 = 0; e0 < T; e0 += 32) {
        const int e = e0 + lane;
        bool ev = false;
        double t = 0, pot = 0;
        int sub = 0, q = 0, o = 0, r9 = 0;
        if (e < T) {
            const unsigned w = wp->ent[e];
            q = (int)(w & 0xffffu); r9 = (int)((w >> 16) & 15u); o = (int)(w >> 20);
            const SRec &ro = rec[wp->ownp[o]];
            const SRec &rq = rec[q];
            Self me = d_self_from_rec(D, ro);
            double shx = 0, shy = 0, shz = 0;
            if (seam) {      // only tiles on the box boundary hold pairs across a periodic seam
                const int cjx = (int)(rq.cxy & 0xffffu), cjy = (int)(rq.cxy >> 16), cjz = (int)rq.cz;
                shx = d_shift1_sel(me.cx, cjx, D.cellnum[0], D.box[0]);
                shy = d_shift1_sel(me.cy, cjy, D.cellnum[1], D.box[1]);
                shz = d_shift1_sel(me.cz, cjz, D.cellnum[2], D.box[2]);
            }
            ev = d_eval_pair(D, me, rq.atom, rq.type, rq.x, rq.y, rq.z, rq.vx, rq.vy, rq.vz, shx, shy, shz, 0, me.type, &t, &pot, &sub) &&
                 t > 0 && t < SD_INF;      // JobAssign ignores anything else (time > 0, strict < against a running best <= INFINIT)
        }
        const unsigned m = __ballot_sync(0xffffffffu, ev);
        if (ev) {
            const int slot = nev + __popc(m & lt_mask);
            const int dz = r9 / 3, dy = r9 - 3 * dz;
            const int hrow = wp->ownh[o] + dy * HX + dz * HXY;
            const int b0 = lbeg[hrow], b1 = lbeg[hrow + 1], b2 = lbeg[hrow + 2];
            const int dx = (q >= b1) + (q >= b2);
            const int cb = q >= b2 ? b2 : (q >= b1 ? b1 : b0);
            const int k = c_tile_kinv[dx + 3 * r9];
            wp->time[slot] = t; wp->pot[slot] = pot; wp->sub[slot] = (unsigned char)sub;
            wp->key[slot] = (unsigned short)((k << 11) | (q - cb));
            wp->eq[slot] = (unsigned short)q;
            wp->next[slot] = (short)atomicExch(&wp->head[o], slot);
        }
        nev += __popc(m);
        if (nev > TILE_ECAP - 32) {
            __syncwarp();
            d_pool_fold(wp, lane, f);
            nev = 0;
            __syncwarp();
        }
    }
    __syncwarp();
    d_pool_fold(wp, lane, f);
    __syncwarp();
}

__global__ void __launch_bounds__(TILE_NT, TILE_MINBLOCKS) k_predict_all_tile(DevP D, SlotBuf dst, Round0Out out, i64 g_base) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int TX = D.tile_tx, TY = D.tile_ty, TZ = D.tile_tz;
    const int HX = TX + 2, HY = TY + 2, HZ = TZ + 2, HXY = HX * HY, HC = HXY * HZ;
    const int HCAP = D.tile_hcap;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // shared-memory layout (sdmd_k2_tile_configure sizes it)
    unsigned long long *bar = reinterpret_cast<unsigned long long *>(smem);
    int *misc = reinterpret_cast<int *>(smem + 8);                           // [0] halo atoms (or HCAP + 1: not staged), [1] own atoms
    SRec *rec = reinterpret_cast<SRec *>(smem + 16);
    float4 *loc = reinterpret_cast<float4 *>(rec + HCAP);
    WarpPool *pools = reinterpret_cast<WarpPool *>(loc + HCAP);
    int *lbeg = reinterpret_cast<int *>(pools + TILE_WARPS);                 // [HC + 1] first staged atom of halo cell h (h: x fastest)
    int *gbeg = lbeg + HC + 1;                                               // [HC] its first sorted position in HBM
    int *rowoff = gbeg + HC;                                                 // [TY*TZ + 1] own atoms in front of own row r
    int *rowbeg = rowoff + TY * TZ + 1;                                      // [TY*TZ] first staged atom of own row r
    int *rowh = rowbeg + TY * TZ;                                            // [TY*TZ] halo cell (0, hy-1, hz-1) of own row r

    const int nx = D.cellnum[0], ny = D.cellnum[1], nz = D.cellnum[2];
    const int ntx = (nx + TX - 1) / TX, nty = (ny + TY - 1) / TY;
    const int bid = blockIdx.x;
    const int x0 = (bid % ntx) * TX, y0 = ((bid / ntx) % nty) * TY, z0 = (bid / (ntx * nty)) * TZ;
    const int TXe = min(TX, nx - x0), TYe = min(TY, ny - y0), TZe = min(TZ, nz - z0);
    const bool seam = x0 == 0 || y0 == 0 || z0 == 0 || x0 + TXe == nx || y0 + TYe == ny || z0 + TZe == nz;

    if (tid == 0) mbar_init(bar, 1);
    // ---- halo cells: sorted-position range of each (warp per x-row of halo cells) ---------------------------------------
    for (int row = warp; row < HY * HZ; row += TILE_WARPS) {
        const int hz = row / HY, hy = row - hz * HY;
        int gy = y0 - 1 + hy, gz = z0 - 1 + hz;
        if (gy < 0) gy += ny; else if (gy >= ny) gy -= ny;
        if (gz < 0) gz += nz; else if (gz >= nz) gz -= nz;
        const bool rowok = hy <= TYe + 1 && hz <= TZe + 1;
        for (int hx = lane; hx < HX; hx += 32) {
            int c0 = 0, c1 = 0;
            if (rowok && hx <= TXe + 1) {
                int gx = x0 - 1 + hx;
                if (gx < 0) gx += nx; else if (gx >= nx) gx -= nx;
                int g = (gz * ny + gy) * nx + gx;
                c0 = __ldg(&D.cell_start[g]);
                c1 = __ldg(&D.cell_start[g + 1]);
            }
            gbeg[row * HX + hx] = c0;
            lbeg[row * HX + hx] = c1 - c0;
        }
    }
    __syncthreads();
    if (warp == 0) {   // exclusive scan of the counts, in place
        int per = (HC + 31) / 32, h0 = lane * per, h1 = min(HC, h0 + per);
        int s = 0, big = 0;
        for (int h = h0; h < h1; h++) { s += lbeg[h]; big |= lbeg[h] >= 2048; }   // position inside a cell is an 11-bit field of the visiting rank
        big = __any_sync(0xffffffffu, big);
        int inc = s;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            int t = __shfl_up_sync(0xffffffffu, inc, off);
            if (lane >= off) inc += t;
        }
        int run = inc - s;
        for (int h = h0; h < h1; h++) { int c = lbeg[h]; lbeg[h] = run; run += c; }
        int total = __shfl_sync(0xffffffffu, inc, 31);
        if (lane == 0) {
            lbeg[HC] = total;
            misc[0] = big ? HCAP + 1 : total;
            if (total <= HCAP && !big) mbar_expect_tx(bar, (unsigned)total * 80u);
        }
    }
    __syncthreads();
    const int H = misc[0];
    const bool staged = H <= HCAP;
    // ---- TMA: one bulk copy pair per contiguous run of a halo row ----------------------------------------------------
    if (staged) {
        for (int row = tid; row < HY * HZ; row += TILE_NT) {
            const int hz = row / HY, hy = row - hz * HY;
            if (hy > TYe + 1 || hz > TZe + 1) continue;
            int hrow = row * HX;
            int hs = hrow;
            for (int h = hrow + 1; h <= hrow + TXe + 2; h++) {
                bool brk = (h == hrow + TXe + 2) || (gbeg[h] != gbeg[hs] + (lbeg[h] - lbeg[hs]));
                if (brk) {
                    int cnt = lbeg[h] - lbeg[hs];
                    if (cnt > 0) {
                        bulk_g2s(rec + lbeg[hs], D.srec + gbeg[hs], (unsigned)cnt * 64u, bar);
                        bulk_g2s(loc + lbeg[hs], D.sloc + gbeg[hs], (unsigned)cnt * 16u, bar);
                    }
                    hs = h;
                }
            }
        }
    }
    if (tid == TILE_NT - 1) {   // own rows (meanwhile)
        int acc = 0, r = 0;
        for (int hz = 1; hz <= TZe; hz++)
            for (int hy = 1; hy <= TYe; hy++, r++) {
                int h1 = (hz * HY + hy) * HX + 1;
                rowoff[r] = acc;
                rowbeg[r] = staged ? lbeg[h1] : gbeg[h1];   // unstaged tiles walk the sorted arrays in HBM
                rowh[r] = ((hz - 1) * HY + (hy - 1)) * HX;
                acc += lbeg[h1 + TXe] - lbeg[h1];
            }
        rowoff[r] = acc;
        misc[1] = acc;
    }
    for (int k = tid; k < TILE_WARPS * 32; k += TILE_NT) pools[k >> 5].head[k & 31] = -1;
    if (staged) mbar_wait(bar, 0);
    __syncthreads();
    const int nown = misc[1];
    const int nrow = TYe * TZe;

    if (!staged) {
        // a tile whose halo does not fit the staging buffers (a very uneven box): exact thread-per-atom scan from HBM
        for (int a = tid; a < nown; a += TILE_NT) {
            int r = 0;
            while (r + 1 < nrow && rowoff[r + 1] <= a) r++;
            int p = rowbeg[r] + (a - rowoff[r]);
            int i = D.sorted_atom[p];
            Slot cur = d_reset_slot(D.ev_pot[i], D.ev_sub[i], D.ev_pcounter[i]);
            double g = 0;
            if (D.thermostat_on) { i64 gi = g_base + (i - 1); g = D.gauss[gi >= D.g_cap ? gi - D.g_cap : gi]; }
            SlotSrc none = {};
            int np = 0;
            double m2 = SD_INF;
            PairBest pb;
            cur = d_predict_atom<true, true, true>(D, none, i, cur, g, 0, &np, &m2, &pb, -1);
            d_round0_store(D, dst, out, i, cur, pb, np);
        }
        return;
    }

    // ---- tile-relative single-precision positions: cell-local + whole cells (lane per halo cell of a row) ---------------
    const float csx = D.cellsize_f[0], csy = D.cellsize_f[1], csz = D.cellsize_f[2];
    for (int row = warp; row < HY * HZ; row += TILE_WARPS) {
        const int hz = row / HY, hy = row - hz * HY;
        const float oy = (float)hy * csy, oz = (float)hz * csz;
        for (int hx = lane; hx < HX; hx += 32) {
            const float ox = (float)hx * csx;
            const int q1 = lbeg[row * HX + hx + 1];
            for (int q = lbeg[row * HX + hx]; q < q1; q++) {
                float4 v = loc[q];
                v.x += ox; v.y += oy; v.z += oz;
                loc[q] = v;
            }
        }
    }
    __syncthreads();

    WarpPool *wp = pools + warp;
    const float cut2f = D.tile_cut2f;
    for (int base = 0; base < nown; base += TILE_NT) {
        const int a = base + tid;
        const bool active = a < nown;
        int p_self = 0, ncand = 0, hrow0 = 0;
        float mx = 0, my = 0, mz = 0;
        float ex0 = 3e38f, ex2 = 3e38f, ey0 = 3e38f, ey2 = 3e38f, ez0 = 3e38f, ez2 = 3e38f;   // inactive lanes: every row pruned
        if (active) {
            int r = 0;
            while (r + 1 < nrow && rowoff[r + 1] <= a) r++;
            p_self = rowbeg[r] + (a - rowoff[r]);
            const int hx = (int)(rec[p_self].cxy & 0xffffu) - x0 + 1;
            const int hq = rowh[r];                       // halo cell (0, hy-1, hz-1)
            hrow0 = hq + (hx - 1);                        // halo cell (hx-1, hy-1, hz-1): first cell of the first of the 9 x-rows
            const int hz1 = hq / HXY, hy1 = (hq - hz1 * HXY) / HX;      // hy-1, hz-1
            float4 m = loc[p_self];
            mx = m.x; my = m.y; mz = m.z;
            // squared distance from the atom to the nearest face of the lower / upper neighbour cell along each axis: no atom
            // of a cell can be closer than the sum over the axes on which the cell is offset
            float lx = mx - (float)hx * csx, ly = my - (float)(hy1 + 1) * csy, lz = mz - (float)(hz1 + 1) * csz;
            lx = fmaxf(lx, 0.f); ly = fmaxf(ly, 0.f); lz = fmaxf(lz, 0.f);
            float ux = fmaxf(csx - lx, 0.f), uy = fmaxf(csy - ly, 0.f), uz = fmaxf(csz - lz, 0.f);
            ex0 = lx * lx; ex2 = ux * ux; ey0 = ly * ly; ey2 = uy * uy; ez0 = lz * lz; ez2 = uz * uz;
        }
        wp->ownp[lane] = p_self;
        wp->ownh[lane] = hrow0;
        Fold f;
        f.t = SD_INF; f.m2 = SD_INF; f.pot = 0; f.key = 0x7fffffff; f.q = -1; f.sub = 0;
        int cnt = 0;        // private hit list fill
        __syncwarp();
        // ---- phase A: 9 x-rows of 3 cells; rows and end cells no atom of which can be inside the superset sphere are skipped ----
#pragma unroll 1
        for (int r9 = 0; r9 < 9; r9++) {      // (not unrolled: one copy of the evaluation code)
            const int dzi = r9 / 3, dyi = r9 - 3 * dzi;
            const float eyz = (dyi == 0 ? ey0 : dyi == 2 ? ey2 : 0.f) + (dzi == 0 ? ez0 : dzi == 2 ? ez2 : 0.f);
            int start = 0, len = 0;
            if (active) {
                const int hrow = hrow0 + dyi * HX + dzi * HXY;
                const int b0 = lbeg[hrow], b3 = lbeg[hrow + 3];
                ncand += b3 - b0;                         // the evaluation count takes every candidate (TimePrediction.c:105)
                if (eyz <= cut2f) {
                    start = (eyz + ex0 <= cut2f) ? b0 : lbeg[hrow + 1];
                    len = ((eyz + ex2 <= cut2f) ? b3 : lbeg[hrow + 2]) - start;
                }
            }
            const int maxlen = __reduce_max_sync(0xffffffffu, len);
            for (int k = 0; k < maxlen; k++) {
                const int q = start + k;
                const float4 c = loc[q];
                const float rx = mx - c.x, ry = my - c.y, rz = mz - c.z;
                const float r2 = fmaf(rz, rz, fmaf(ry, ry, rx * rx));
                const bool hit = (k < len) & (r2 <= cut2f) & (q != p_self);
                if (hit) wp->priv[cnt * 32 + lane] = (unsigned)q | ((unsigned)r9 << 16);
                cnt += hit;
                if (__any_sync(0xffffffffu, cnt == TILE_LCAP)) {
                    d_pool_flush(D, wp, rec, lbeg, HX, HXY, seam, lane, cnt, f);
                    cnt = 0;
                }
            }
        }
        if (__any_sync(0xffffffffu, cnt > 0)) d_pool_flush(D, wp, rec, lbeg, HX, HXY, seam, lane, cnt, f);
        // ---- phase C ---------------------------------------------------------------------------------------------------
        if (active) {
            const SRec &rs = rec[p_self];
            Self me = d_self_from_rec(D, rs);
            int i = me.i;
            Slot cur = d_reset_slot(D.ev_pot[i], D.ev_sub[i], D.ev_pcounter[i]);
            if (f.t < SD_INF) {
                int j = rec[f.q].atom;
                cur.time = f.t; cur.pot = f.pot; cur.sub = f.sub; cur.type = SDMD_EV_COLI; cur.partner = j;
                cur.pcounter = D.ev_counter[j];
                cur.m2 = f.m2; cur.bp_time = f.t; cur.bp_j = j; cur.sp_time = f.m2;
            }
            PairBest pb;
            pb.bp_time = cur.bp_time; pb.sp_time = cur.sp_time; pb.bp_j = cur.bp_j;
            d_pcc_time(D, me, cur);
            if (D.thermostat_on) {
                i64 gi = g_base + (i - 1);
                d_thermostat_time(D, cur, D.gauss[gi >= D.g_cap ? gi - D.g_cap : gi]);
            }
            d_round0_store(D, dst, out, i, cur, pb, ncand - 1);
        }
        __syncwarp();
    }
}

// ----------------------------------------------------------------------------------------------------------------------
// Variant 2 (SDMD_K2_TILE=2): the neighbour filter stays the stand-alone kernel (k_predict_all_filter: survivor rows in HBM, in
// visiting order), only the EVALUATION becomes a tile kernel.  The 64-byte records of a block's halo are staged by TMA exactly
// as above; a thread then walks its atom's survivor row in order and evaluates every survivor from shared memory — no gather
// from HBM/L2 per survivor, which is what the stand-alone evaluation kernel waits on (long-scoreboard stalls).  The fold is the
// sequential one (JobAssign per survivor in visiting order), so no visiting ranks are needed.
__global__ void __launch_bounds__(TILE_NT, TILE_MINBLOCKS) k_predict_all_evaltile(DevP D, SlotBuf dst, Round0Out out, i64 g_base) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int TX = D.tile_tx, TY = D.tile_ty, TZ = D.tile_tz;
    const int HX = TX + 2, HY = TY + 2, HZ = TZ + 2, HXY = HX * HY, HC = HXY * HZ;
    const int HCAP = D.tile_hcap;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    unsigned long long *bar = reinterpret_cast<unsigned long long *>(smem);
    int *misc = reinterpret_cast<int *>(smem + 8);
    SRec *rec = reinterpret_cast<SRec *>(smem + 16);
    int *lbeg = reinterpret_cast<int *>(rec + HCAP);
    int *gbeg = lbeg + HC + 1;
    int *rowoff = gbeg + HC;
    int *rowbeg = rowoff + TY * TZ + 1;
    int *rowh = rowbeg + TY * TZ;

    const int nx = D.cellnum[0], ny = D.cellnum[1], nz = D.cellnum[2];
    const int ntx = (nx + TX - 1) / TX, nty = (ny + TY - 1) / TY;
    const int bid = blockIdx.x;
    const int x0 = (bid % ntx) * TX, y0 = ((bid / ntx) % nty) * TY, z0 = (bid / (ntx * nty)) * TZ;
    const int TXe = min(TX, nx - x0), TYe = min(TY, ny - y0), TZe = min(TZ, nz - z0);
    const bool seam = x0 == 0 || y0 == 0 || z0 == 0 || x0 + TXe == nx || y0 + TYe == ny || z0 + TZe == nz;

    if (tid == 0) mbar_init(bar, 1);
    for (int row = warp; row < HY * HZ; row += TILE_WARPS) {
        const int hz = row / HY, hy = row - hz * HY;
        int gy = y0 - 1 + hy, gz = z0 - 1 + hz;
        if (gy < 0) gy += ny; else if (gy >= ny) gy -= ny;
        if (gz < 0) gz += nz; else if (gz >= nz) gz -= nz;
        const bool rowok = hy <= TYe + 1 && hz <= TZe + 1;
        for (int hx = lane; hx < HX; hx += 32) {
            int c0 = 0, c1 = 0;
            if (rowok && hx <= TXe + 1) {
                int gx = x0 - 1 + hx;
                if (gx < 0) gx += nx; else if (gx >= nx) gx -= nx;
                int g = (gz * ny + gy) * nx + gx;
                c0 = __ldg(&D.cell_start[g]);
                c1 = __ldg(&D.cell_start[g + 1]);
            }
            gbeg[row * HX + hx] = c0;
            lbeg[row * HX + hx] = c1 - c0;
        }
    }
    __syncthreads();
    if (warp == 0) {
        int per = (HC + 31) / 32, h0 = lane * per, h1 = min(HC, h0 + per);
        int s = 0;
        for (int h = h0; h < h1; h++) s += lbeg[h];
        int inc = s;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            int t = __shfl_up_sync(0xffffffffu, inc, off);
            if (lane >= off) inc += t;
        }
        int run = inc - s;
        for (int h = h0; h < h1; h++) { int c = lbeg[h]; lbeg[h] = run; run += c; }
        int total = __shfl_sync(0xffffffffu, inc, 31);
        if (lane == 0) {
            lbeg[HC] = total;
            misc[0] = total;
            if (total <= HCAP) mbar_expect_tx(bar, (unsigned)total * 64u);
        }
    }
    __syncthreads();
    const bool staged = misc[0] <= HCAP;
    if (staged) {
        for (int row = tid; row < HY * HZ; row += TILE_NT) {
            const int hz = row / HY, hy = row - hz * HY;
            if (hy > TYe + 1 || hz > TZe + 1) continue;
            int hrow = row * HX;
            int hs = hrow;
            for (int h = hrow + 1; h <= hrow + TXe + 2; h++) {
                bool brk = (h == hrow + TXe + 2) || (gbeg[h] != gbeg[hs] + (lbeg[h] - lbeg[hs]));
                if (brk) {
                    int cnt = lbeg[h] - lbeg[hs];
                    if (cnt > 0) bulk_g2s(rec + lbeg[hs], D.srec + gbeg[hs], (unsigned)cnt * 64u, bar);
                    hs = h;
                }
            }
        }
    }
    if (tid == TILE_NT - 1) {
        int acc = 0, r = 0;
        for (int hz = 1; hz <= TZe; hz++)
            for (int hy = 1; hy <= TYe; hy++, r++) {
                int h1 = (hz * HY + hy) * HX + 1;
                rowoff[r] = acc;
                rowbeg[r] = h1;                              // first own halo cell of the row
                rowh[r] = h1 - 1;
                acc += lbeg[h1 + TXe] - lbeg[h1];
            }
        rowoff[r] = acc;
        misc[1] = acc;
    }
    if (staged) mbar_wait(bar, 0);
    __syncthreads();
    const int nown = misc[1];
    const int nrow = TYe * TZe;
    const int cap = D.surv_cap;
    SlotSrc none = {};
    for (int a = tid; a < nown; a += TILE_NT) {
        int r = 0;
        while (r + 1 < nrow && rowoff[r + 1] <= a) r++;
        const int h1 = rowbeg[r];
        const int pos_in_row = a - rowoff[r];
        const int p_glob = gbeg[h1] + pos_in_row;             // own cells of a row are contiguous in the sorted arrays
        const int ns = D.surv_n[p_glob];
        double g = 0;
        int i;
        Slot cur;
        PairBest pb;
        int np = 0;
        if (!staged || ns < 0) {
            // halo too large for the staging buffer, or a survivor row that overflowed: the exact scan from HBM
            i = D.sorted_atom[p_glob];
            cur = d_reset_slot(D.ev_pot[i], D.ev_sub[i], D.ev_pcounter[i]);
            if (D.thermostat_on) { i64 gi = g_base + (i - 1); g = D.gauss[gi >= D.g_cap ? gi - D.g_cap : gi]; }
            double m2 = SD_INF;
            cur = d_predict_atom<true, true, true>(D, none, i, cur, g, 0, &np, &m2, &pb, p_glob);
            d_round0_store(D, dst, out, i, cur, pb, np);
            continue;
        }
        const int p_self = lbeg[h1] + pos_in_row;
        const SRec &rs = rec[p_self];
        Self me = d_self_from_rec(D, rs);
        i = me.i;
        const int h_self = h1 + ((int)(rs.cxy & 0xffffu) - x0);
        cur = d_reset_slot(D.ev_pot[i], D.ev_sub[i], D.ev_pcounter[i]);
        const unsigned *row = D.surv + (size_t)p_glob * cap;
        for (int s = 0; s < ns; s++) {
            const unsigned w = __ldg(&row[s]);
            const int k = (int)(w >> 27);
            // DMD.c:118-120 offsets of scan slot k, 2 bits per slot (value + 1)
            const unsigned long long PX = 0xa80894054a959ULL, PY = 0x2222160111a665ULL, PZ = 0x28096084059a95ULL;
            const int dx = (int)((PX >> (2 * k)) & 3) - 1, dy = (int)((PY >> (2 * k)) & 3) - 1, dz = (int)((PZ >> (2 * k)) & 3) - 1;
            const int hj = h_self + dx + dy * HX + dz * HXY;
            const int q = lbeg[hj] + ((int)(w & 0x7ffffffu) - gbeg[hj]);
            const SRec &rq = rec[q];
            double shx = 0, shy = 0, shz = 0;
            if (seam) {
                const int cjx = (int)(rq.cxy & 0xffffu), cjy = (int)(rq.cxy >> 16), cjz = (int)rq.cz;
                shx = d_shift1_sel(me.cx, cjx, D.cellnum[0], D.box[0]);
                shy = d_shift1_sel(me.cy, cjy, D.cellnum[1], D.box[1]);
                shz = d_shift1_sel(me.cz, cjz, D.cellnum[2], D.box[2]);
            }
            double t, pot;
            int sub;
            if (d_eval_pair(D, me, rq.atom, rq.type, rq.x, rq.y, rq.z, rq.vx, rq.vy, rq.vz, shx, shy, shz, 0, me.type, &t, &pot, &sub))
                d_job_assign(D, cur, rq.atom, t, pot, sub, SDMD_EV_COLI);
        }
        pb.bp_time = cur.bp_time; pb.sp_time = cur.sp_time; pb.bp_j = cur.bp_j;
        d_pcc_time(D, me, cur);
        if (D.thermostat_on) {
            i64 gi = g_base + (i - 1);
            d_thermostat_time(D, cur, D.gauss[gi >= D.g_cap ? gi - D.g_cap : gi]);
        }
        d_round0_store(D, dst, out, i, cur, pb, D.surv_np[p_glob]);
    }
}

static size_t tile_smem_bytes(int tx, int ty, int tz, int hcap) {
    size_t hc = (size_t)(tx + 2) * (ty + 2) * (tz + 2);
    return 16 + (size_t)hcap * 80 + sizeof(WarpPool) * TILE_WARPS + (2 * hc + 1 + 3 * (size_t)ty * tz + 1) * sizeof(int) + 16;
}

int sdmd_k2_tile_configure(DevP &P, int device) {
    P.tile_tx = P.tile_ty = P.tile_tz = P.tile_hcap = P.tile_smem = P.tile_mode = 0;
    // Opt-in (SDMD_K2_TILE=1) while it is slower than the two-kernel path: bit-exact on every GEL parity case and 158 MB of
    // DRAM traffic per launch at 10^6 atoms (two-kernel path: 391 MB), but 430 us against 284 us — the kernel is bound by
    // instruction issue at two blocks per SM (profiles/r2_k2_tile.md), not by memory.
    const char *env = getenv("SDMD_K2_TILE");
    P.tile_mode = (env && (env[0] == '1' || env[0] == '2')) ? env[0] - '0' : 0;     // 1: fused tile kernel, 2: filter kernel + evaluation tile kernel
    if (!P.tile_mode) return 1;
    if (!(P.gel_mode && !P.hydrogen_bond && !P.single_body && P.no_cons)) return 1;   // bonds / HB / walls / constrained atoms: two-kernel path
    if (P.n >= (1 << 27) || !P.sloc) return 1;
    const int nx = P.cellnum[0], ny = P.cellnum[1], nz = P.cellnum[2];
    const double rho = (double)P.n / ((double)nx * ny * nz);        // atoms per cell
    // own cells per block: about 3/4 of the block's threads in atoms; x-rows as long as possible (fewer, longer bulk copies)
    double want = 0.75 * TILE_NT / (rho > 0.05 ? rho : 0.05);
    int tx = nx < 8 ? nx : 8, ty = 1, tz = 1;
    while ((double)tx * ty * tz < want) {
        if (ty <= tz && ty < ny && ty < 8) ty++;
        else if (tz < nz && tz < 8) tz++;
        else if (ty < ny && ty < 8) ty++;
        else if (tx < nx && tx < 16) tx++;
        else break;
    }
    int max_smem = 0;
    if (cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, device) != cudaSuccess) return -1;
    for (;;) {
        size_t hc = (size_t)(tx + 2) * (ty + 2) * (tz + 2);
        int hcap = (int)(hc * rho * 1.15) + 64;
        hcap = (hcap + 3) & ~3;
        size_t bytes = tile_smem_bytes(tx, ty, tz, hcap);
        if (hcap < 65536 && bytes <= (size_t)max_smem * 45 / 100) {   // two blocks per SM
            P.tile_tx = tx; P.tile_ty = ty; P.tile_tz = tz; P.tile_hcap = hcap; P.tile_smem = (int)bytes;
            break;
        }
        if (tz > 1 && tz >= ty) tz--;
        else if (ty > 1) ty--;
        else if (tx > 1) tx--;
        else return 1;
    }
    // superset sphere of the single-precision test: coordinates are relative to the halo corner (extent L), so the error of
    // r^2 is bounded by ~2 r sqrt(3) * 4 ulp(L) plus a few ulp(r^2); 64 * 2^-24 * L * r covers it several times over
    double L = 0;
    const int t3[3] = {P.tile_tx, P.tile_ty, P.tile_tz};
    for (int k = 0; k < 3; k++) { double e = (t3[k] + 2) * P.cellsize[k]; if (e > L) L = e; }
    double margin = 1e-4 * P.cut2 + 64.0 * 5.9604644775390625e-08 * L * sqrt(P.cut2);
    P.tile_cut2f = nextafterf((float)(P.cut2 + margin), INFINITY);
    if (cudaFuncSetAttribute(k_predict_all_tile, cudaFuncAttributeMaxDynamicSharedMemorySize, P.tile_smem) != cudaSuccess) return -1;
    if (cudaFuncSetAttribute(k_predict_all_evaltile, cudaFuncAttributeMaxDynamicSharedMemorySize, P.tile_smem) != cudaSuccess) return -1;
    return 0;
}

int sdmd_k2_tile_launch(const DevP &P, SlotBuf dst, Round0Out out, long long g_base, cudaStream_t stream) {
    const int ntx = (P.cellnum[0] + P.tile_tx - 1) / P.tile_tx, nty = (P.cellnum[1] + P.tile_ty - 1) / P.tile_ty,
              ntz = (P.cellnum[2] + P.tile_tz - 1) / P.tile_tz;
    if (P.tile_mode == 2) k_predict_all_evaltile<<<ntx * nty * ntz, TILE_NT, P.tile_smem, stream>>>(P, dst, out, g_base);
    else k_predict_all_tile<<<ntx * nty * ntz, TILE_NT, P.tile_smem, stream>>>(P, dst, out, g_base);
    return cudaGetLastError() == cudaSuccess ? 0 : -1;
}
