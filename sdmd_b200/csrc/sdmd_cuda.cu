// sdmd_cuda.cu — kernels and C-ABI of libsdmd_cuda.so (sm_100a).  See include/sdmd_cuda.h.
//
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -fmad=false -lineinfo -O3 -shared -Xcompiler -fPIC
// There is no CPU fallback anywhere in this file: every entry point needs a CUDA device.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <vector>
#include <dlfcn.h>
#include <cooperative_groups.h>
#include "sdmd_dev.cuh"
#include "sdmd_host.h"

// ===================================================================================================
// host context
// ===================================================================================================
static thread_local char g_err[512] = "";
static int set_err(int code, const char *what, cudaError_t e = cudaSuccess) {
    if (e != cudaSuccess) snprintf(g_err, sizeof(g_err), "%s: %s", what, cudaGetErrorString(e));
    else snprintf(g_err, sizeof(g_err), "%s", what);
    return code;
}
#define CK(call)                                                                  \
    do {                                                                          \
        cudaError_t e_ = (call);                                                  \
        if (e_ != cudaSuccess) return set_err(SDMD_ERR_CUDA, #call, e_);          \
    } while (0)

struct KStat { double ms = 0; long n = 0; };
struct PendingEv { std::string name; cudaEvent_t a, b; };

struct sdmd_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    DevP P;
    Ctl *d_ctl = nullptr;
    Ctl *d_ctl_scratch = nullptr;   // counters of the re-evaluation rounds of Predict(all) (see k_predict_all_fixpoint)
    Ctl *h_ctl = nullptr;  // pinned
    std::vector<void *> bufs;
    int n = 0;
    // scratch for predict_all / cell_build
    double *b_time = nullptr, *b_pot = nullptr, *c_time = nullptr, *c_pot = nullptr;
    int *b_partner = nullptr, *c_partner = nullptr;
    i64 *b_pcounter = nullptr, *c_pcounter = nullptr;
    uint8_t *b_type = nullptr, *b_sub = nullptr, *c_type = nullptr, *c_sub = nullptr;
    int *npair = nullptr, *pcount = nullptr, *pwho = nullptr, *queue_a = nullptr, *queue_b = nullptr, *nqueue = nullptr;
    double *second = nullptr, *bp_time = nullptr;
    int *bp_j = nullptr;
    int *cell_of = nullptr, *cell_count = nullptr, *cell_rank = nullptr, *cell_tmp = nullptr, *scan_tmp = nullptr;
    int n_draws = 0;
    void *scratch[4] = {nullptr, nullptr, nullptr, nullptr};   // grow-on-demand device buffers of the occasional calls (no cudaMalloc per call)
    size_t scratch_bytes[4] = {0, 0, 0, 0};
    unsigned long long *scan_state = nullptr;      // single-pass scan of the cell histogram: one word per block of 2048 cells
    unsigned *scan_ticket = nullptr, scan_ticket_base = 0, scan_epoch = 0;
    int *d_cell_start = nullptr, *d_sorted_atom = nullptr;
    SRec *d_srec = nullptr;
    float4 *d_sloc = nullptr;
    int *d_scell = nullptr;
    int *d_leaf_atom = nullptr;  // static leaf permutation of the tournament tree
    double *d_gauss = nullptr;
    i64 g_filled = 0;
    sdmd_trace_rec *d_trace = nullptr;
    i64 trace_cap = 0;
    unsigned long long *d_sum = nullptr;
    bool cells_fresh = false;
    int loop_mode = SDMD_LOOP_AUTO;
    int fix_grid = 0;
    i64 pa_base[3] = {0, 0, 0};
    int cluster_size = 16;     // blocks per cluster of the mid-size loop (16 = B200's non-portable maximum; 8 measured slower above 8000 atoms)
    double *x2 = nullptr, *y2 = nullptr, *z2 = nullptr, *t2 = nullptr;   // second coordinate/time buffer of the overlapped sweep
    bool counters_known = false;         // h_ctl's g_pos / pair_predictions match the device (nothing consumed since the last pull)
    int timing = 1;                      // 0 off, 1 per stage, 2 also the sub-stages of predict_all's fix-up
    std::map<std::string, KStat> kstat;
    std::vector<PendingEv> pending;      // launched, not yet resolved (kept asynchronous)
    std::vector<cudaEvent_t> ev_pool;
    long launches = 0;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr, evt0 = nullptr, evt1 = nullptr;
    cudaStream_t copy_stream = nullptr;  // second stream of sdmd_cuda_prepare (velocities + gaussians overlap the cell build)
    cudaEvent_t ev_copy = nullptr, ev_main = nullptr;
    // NCCL
    void *nccl_lib = nullptr;
    void *nccl_comm = nullptr;
    int nranks = 1, rank = 0;
    void *d_xchg = nullptr, *h_xchg = nullptr;   // exchange records: device (all ranks + own send slot), pinned host mirror
};

template <typename T>
static int dalloc(sdmd_ctx *c, T **p, size_t count) {
    void *q = nullptr;
    cudaError_t e = cudaMalloc(&q, (count ? count : 1) * sizeof(T));
    if (e != cudaSuccess) return set_err(SDMD_ERR_CUDA, "cudaMalloc", e);
    e = cudaMemsetAsync(q, 0, (count ? count : 1) * sizeof(T), c->stream);
    if (e != cudaSuccess) return set_err(SDMD_ERR_CUDA, "cudaMemset", e);
    c->bufs.push_back(q);
    *p = (T *)q;
    return 0;
}
// slot: 0 renew list / block sums, 1 energy addends, 2 replica descriptors, 3 per-replica results
static int scratch_get(sdmd_ctx *c, int slot, size_t bytes, void **out) {
    if (c->scratch_bytes[slot] < bytes) {
        if (c->scratch[slot]) { cudaStreamSynchronize(c->stream); cudaFree(c->scratch[slot]); c->scratch[slot] = nullptr; c->scratch_bytes[slot] = 0; }
        size_t want = bytes + bytes / 2 + 256;
        cudaError_t e = cudaMalloc(&c->scratch[slot], want);
        if (e != cudaSuccess) return set_err(SDMD_ERR_CUDA, "cudaMalloc (scratch)", e);
        c->scratch_bytes[slot] = want;
    }
    *out = c->scratch[slot];
    return 0;
}
template <typename T>
static int dupload(sdmd_ctx *c, T **p, const T *src, size_t count) {
    int rc = dalloc(c, p, count);
    if (rc) return rc;
    if (src && count) CK(cudaMemcpyAsync(*p, src, count * sizeof(T), cudaMemcpyHostToDevice, c->stream));
    return 0;
}
#define TRY(x)              \
    do {                    \
        int rc_ = (x);      \
        if (rc_) return rc_; \
    } while (0)

static cudaEvent_t pool_get(sdmd_ctx *c) {
    if (!c->ev_pool.empty()) { cudaEvent_t e = c->ev_pool.back(); c->ev_pool.pop_back(); return e; }
    cudaEvent_t e;
    cudaEventCreate(&e);
    return e;
}
// resolve every pending (begin, end) event pair into the per-kernel accumulators
static void resolve_pending(sdmd_ctx *c) {
    for (PendingEv &p : c->pending) {
        cudaEventSynchronize(p.b);
        float ms = 0;
        cudaEventElapsedTime(&ms, p.a, p.b);
        KStat &k = c->kstat[p.name];
        k.ms += ms;
        k.n++;
        c->ev_pool.push_back(p.a);
        c->ev_pool.push_back(p.b);
    }
    c->pending.clear();
}
// CUDA-event bracket on the launching stream; does not synchronise (resolved lazily)
struct Timer {
    sdmd_ctx *c;
    PendingEv p;
    bool on;
    Timer(sdmd_ctx *c_, const char *n, int level = 1) : c(c_), on(c_->timing >= level) {
        if (on) {
            p.name = n;
            p.a = pool_get(c);
            p.b = pool_get(c);
            cudaEventRecord(p.a, c->stream);
        }
    }
    ~Timer() {
        if (on) {
            cudaEventRecord(p.b, c->stream);
            c->pending.push_back(p);
            if (c->pending.size() > 4096) resolve_pending(c);
        }
    }
};

// ===================================================================================================
// kernels: cell list (K1)
// ===================================================================================================
// PBC("full"), Events.c:218-228
__global__ void k_pbc_full(DevP D) {
    int i = blockIdx.x * blockDim.x + threadIdx.x + 1;
    if (i > D.n) return;
    double c[3] = {D.x[i], D.y[i], D.z[i]};
#pragma unroll
    for (int k = 0; k < 3; k++) {
        if (!(D.pbc_mask >> k & 1)) continue;   // closed by a wall, Events.c:210-216
        if (c[k] >= d_box(D, k)) c[k] -= d_box(D, k);
        else if (c[k] < 0) c[k] += d_box(D, k);
    }
    // only what moved is written back (nothing at all for a box that is already wrapped: 24 B per atom instead of 48)
    if (c[0] != D.x[i]) D.x[i] = c[0];
    if (c[1] != D.y[i]) D.y[i] = c[1];
    if (c[2] != D.z[i]) D.z[i] = c[2];
}

// LinkList("full") step 1, Events.c:129-150: cell index per atom + histogram
// (the histogram's atomicAdd also hands every atom its arrival rank inside the cell, so the scatter needs no second atomic)
// PBC: PBC("full") (Events.c:218-228) applied on the way — what k_pbc_full does, without reading the coordinates twice
template <bool PBC>
__global__ void k_cell_index(DevP D, int *cell_of, int *cell_count, int *arrival) {
    SDMD_PDL_ENTRY();
    int i = blockIdx.x * blockDim.x + threadIdx.x + 1;
    if (i > D.n) return;
    int idx[3];
    double x = D.x[i], y = D.y[i], z = D.z[i];
    if (PBC) {
        double c[3] = {x, y, z};
#pragma unroll
        for (int k = 0; k < 3; k++) {
            if (!(D.pbc_mask >> k & 1)) continue;   // closed by a wall, Events.c:210-216
            if (c[k] >= d_box(D, k)) c[k] -= d_box(D, k);
            else if (c[k] < 0) c[k] += d_box(D, k);
        }
        if (c[0] != x) D.x[i] = x = c[0];
        if (c[1] != y) D.y[i] = y = c[1];
        if (c[2] != z) D.z[i] = z = c[2];
    }
    int cell = d_cell_of(D, x, y, z, idx);
    D.cell_idx[3 * i] = idx[0]; D.cell_idx[3 * i + 1] = idx[1]; D.cell_idx[3 * i + 2] = idx[2];
    cell_of[i] = cell;
    arrival[i] = atomicAdd(&cell_count[cell], 1);
}

// exclusive scan of cell_count[0..m) -> cell_start[0..m], 2048 items per block:
//   k_scan_local   per-block exclusive scan (4 items per thread, shuffle scan per warp, one pass over the warp sums)
//   k_scan_finish  every block sums the totals of the blocks before it (a few hundred values), adds the offset and also
//                  fills the scatter cursors; for very large grids (> SCAN_DIRECT_BLOCKS blocks) the block totals are
//                  scanned by k_scan_blocks and added by k_scan_add instead
#define SCAN_ITEMS 2048
#define SCAN_DIRECT_BLOCKS 2048
__device__ __forceinline__ int d_block_scan_excl(int v, int *total) {   // 512 threads
    __shared__ int wsum[16];
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    int inc = v;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        int t = __shfl_up_sync(0xffffffffu, inc, off);
        if (lane >= off) inc += t;
    }
    if (lane == 31) wsum[w] = inc;
    __syncthreads();
    if (w == 0) {
        int s = lane < 16 ? wsum[lane] : 0;
#pragma unroll
        for (int off = 1; off < 16; off <<= 1) {
            int t = __shfl_up_sync(0xffffffffu, s, off);
            if (lane >= off) s += t;
        }
        if (lane < 16) wsum[lane] = s;   // inclusive over warps
    }
    __syncthreads();
    int before = w ? wsum[w - 1] : 0;
    *total = wsum[15];
    return before + inc - v;
}
__global__ void __launch_bounds__(512) k_scan_local(const int *in, int *out, int *block_sum, int m) {
    int t0 = blockIdx.x * SCAN_ITEMS + threadIdx.x * 4;
    int v[4];
#pragma unroll
    for (int q = 0; q < 4; q++) v[q] = (t0 + q < m) ? in[t0 + q] : 0;
    int total;
    int ex = d_block_scan_excl(v[0] + v[1] + v[2] + v[3], &total);
#pragma unroll
    for (int q = 0; q < 4; q++) {
        if (t0 + q < m) out[t0 + q] = ex;
        ex += v[q];
    }
    if (threadIdx.x == 0) block_sum[blockIdx.x] = total;
}
// k_scan_finish / k_scan_add also publish the occupancy of every cell (DevP::cell_cnt), the list head of the EMPTY cells
// (k_cell_place writes the heads of the others) and clear the histogram for the next build.
__global__ void __launch_bounds__(512) k_scan_finish(DevP D, int *out, int *count, const int *block_sum, int m, int nb) {
    __shared__ int red[16];
    __shared__ int s_off;
    int part = 0;
    for (int q = threadIdx.x; q < (int)blockIdx.x; q += blockDim.x) part += block_sum[q];
#pragma unroll
    for (int off = 16; off; off >>= 1) part += __shfl_xor_sync(0xffffffffu, part, off);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = part;
    __syncthreads();
    if (threadIdx.x == 0) {
        int s = 0;
        for (int q = 0; q < 16; q++) s += red[q];
        s_off = s;
        if ((int)blockIdx.x == nb - 1) out[m] = s + block_sum[nb - 1];
    }
    __syncthreads();
    int off = s_off;
#pragma unroll
    for (int q = 0; q < 4; q++) {      // consecutive threads, consecutive cells
        int t = blockIdx.x * SCAN_ITEMS + q * 512 + threadIdx.x;
        if (t < m) {
            out[t] += off;
            int cnt = count[t];
            count[t] = 0;
            D.cell_cnt[t] = cnt;
            if (cnt == 0) D.celllist[D.n + t + 1] = 0;
        }
    }
}
__global__ void k_scan_blocks(int *block_sum, int nb) {
    // single thread block; nb is small (<= a few thousand)
    __shared__ int carry;
    if (threadIdx.x == 0) {
        int acc = 0;
        for (int b = 0; b < nb; b++) { int v = block_sum[b]; block_sum[b] = acc; acc += v; }
        carry = acc;
    }
    __syncthreads();
    if (threadIdx.x == 0) block_sum[nb] = carry;
}
__global__ void k_scan_add(int *out, const int *block_sum, int m, int nb) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < m) out[t] += block_sum[t / SCAN_ITEMS];
    if (t == 0) out[m] = block_sum[nb];
}
__global__ void k_scan_add_cells(DevP D, int *out, int *count, const int *block_sum, int m, int nb) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < m) {
        out[t] += block_sum[t / SCAN_ITEMS];
        int cnt = count[t];
        count[t] = 0;
        D.cell_cnt[t] = cnt;
        if (cnt == 0) D.celllist[D.n + t + 1] = 0;
    }
    if (t == 0) out[m] = block_sum[nb];
}

// The same scan in ONE launch: a block scans its 2048 counts, publishes its total and sums the totals of the blocks in front of
// it.  `state[b]` = epoch << 32 | total; the epoch changes with every build, so the array never has to be cleared.  (A
// decoupled look-back over running prefixes was measured first: 16 us, the same as the two launches — ten dependent hops.)  Blocks take their number from a ticket counter (a block only ever waits for
// blocks that started before it).  Also does what k_scan_finish does: cell occupancy, heads of the empty cells, histogram
// cleared for the next build.  Two dependent launches (16 us at 667 k cells) become one (~6 us).
__global__ void __launch_bounds__(512) k_scan_cells(DevP D, int *count, int *out, unsigned long long *state, unsigned *ticket, unsigned ticket_base,
                                                    unsigned epoch, int m, int nb) {
    SDMD_PDL_ENTRY();
    __shared__ int s_b, s_prefix;
    if (threadIdx.x == 0) s_b = (int)(atomicAdd(ticket, 1u) - ticket_base);
    __syncthreads();
    const int b = s_b;
    const int t0 = b * SCAN_ITEMS + threadIdx.x * 4;
    int v[4];
    if (t0 + 3 < m) {
        const int4 q = *reinterpret_cast<const int4 *>(count + t0);
        v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w;
    } else {
#pragma unroll
        for (int q = 0; q < 4; q++) v[q] = (t0 + q < m) ? count[t0 + q] : 0;
    }
    int total;
    int ex = d_block_scan_excl(v[0] + v[1] + v[2] + v[3], &total);
    // publish my total, then add up the totals of all blocks in front of me (one word per thread and 512 blocks): no block waits
    // for another block's PREFIX, so there is no chain of hops — every block is one round trip behind the slowest local scan
    volatile unsigned long long *st = state;
    if (threadIdx.x == 0) st[b] = ((unsigned long long)epoch << 32) | (unsigned)total;
    int part = 0;
    for (int q = threadIdx.x; q < b; q += 512) {
        unsigned long long w;
        do { w = st[q]; } while ((unsigned)(w >> 32) != epoch);
        part += (int)(unsigned)(w & 0xffffffffu);
    }
    __shared__ int red[16];
#pragma unroll
    for (int off = 16; off; off >>= 1) part += __shfl_xor_sync(0xffffffffu, part, off);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = part;
    __syncthreads();
    if (threadIdx.x == 0) {
        int sum = 0;
        for (int q = 0; q < 16; q++) sum += red[q];
        s_prefix = sum;
        if (b == nb - 1) out[m] = sum + total;
    }
    __syncthreads();
    ex += s_prefix;
    if (t0 + 3 < m) {       // four consecutive cells per thread: one 16-byte store per array
        const int e1 = ex + v[0], e2 = e1 + v[1], e3 = e2 + v[2];
        *reinterpret_cast<int4 *>(out + t0) = make_int4(ex, e1, e2, e3);
        *reinterpret_cast<int4 *>(count + t0) = make_int4(0, 0, 0, 0);
        *reinterpret_cast<int4 *>(D.cell_cnt + t0) = make_int4(v[0], v[1], v[2], v[3]);
#pragma unroll
        for (int q = 0; q < 4; q++)
            if (v[q] == 0) D.celllist[D.n + t0 + q + 1] = 0;
        return;
    }
#pragma unroll
    for (int q = 0; q < 4; q++) {
        const int t = t0 + q;
        if (t < m) {
            out[t] = ex;
            ex += v[q];
            count[t] = 0;
            D.cell_cnt[t] = v[q];
            if (v[q] == 0) D.celllist[D.n + t + 1] = 0;
        }
    }
}

// atoms of a cell side by side, in arrival order
__global__ void k_cell_scatter(DevP D, const int *cell_of, const int *arrival, const int *cell_start, int *unordered) {
    SDMD_PDL_ENTRY();
    int i = blockIdx.x * blockDim.x + threadIdx.x + 1;
    if (i > D.n) return;
    unordered[cell_start[cell_of[i]] + arrival[i]] = i;
}

// Every atom finds its final place itself: a fresh LinkList("full") inserts atoms 1..N at the head of their cell's list
// (Events.c:152-153), so inside a cell the list runs by descending atom id — the rank of an atom is the number of cell mates
// with a larger id, its successor the largest smaller id.  One pass then writes the reference-format link list and the
// cell-sorted working set (what k_cell_finish + k_cell_permute did in two passes with a serial sort per cell).
// parts: as k_cell_permute
__global__ void k_cell_place(DevP D, const int *cell_of, const int *cell_start, const int *unordered, int *sorted_atom, SRec *srec, float4 *sloc,
                             int *scell, int parts) {
    SDMD_PDL_ENTRY();
    int i = blockIdx.x * blockDim.x + threadIdx.x + 1;
    if (i > D.n) return;
    if ((parts & 4) && i == 1) D.ctl->cc_floor = ++D.ctl->cc_epoch;   // new link list: every cached candidate list is void
    const int c = cell_of[i];
    const int s0 = cell_start[c], s1 = cell_start[c + 1];
    int rank = 0, next = 0;
    for (int s = s0; s < s1; s++) {
        int j = unordered[s];
        rank += j > i;
        if (j < i && j > next) next = j;
    }
    const int p = s0 + rank;
    sorted_atom[p] = i;
    D.spos[i] = p;
    D.celllist[i] = next;
    if (rank == 0) D.celllist[D.n + c + 1] = i;
    double x = D.x[i], y = D.y[i], z = D.z[i];
    int t = D.type[i];
    // cell index triple from the cell id (three loads fewer per atom; the stage is bound by its load / store instruction count)
    const int nxy = D.cellnum[0] * D.cellnum[1];
    const int cz = c / nxy, crem = c - cz * nxy, cy = crem / D.cellnum[0], cx = crem - cy * D.cellnum[0];
    double2 *rp = reinterpret_cast<double2 *>(&srec[p]);
    // cell-local single-precision copy for the neighbour filter (see d_r0_filter)
    sloc[p] = make_float4((float)(x - cx * D.cellsize[0]), (float)(y - cy * D.cellsize[1]), (float)(z - cz * D.cellsize[2]), __int_as_float(i));
    scell[p] = c;
    rp[0] = make_double2(x, y);
    *reinterpret_cast<int4 *>(rp + 3) = make_int4(t, i, (int)((unsigned)cx | ((unsigned)cy << 16)), cz);
    if (parts & 2) {
        rp[1] = make_double2(z, D.vx[i]);
        rp[2] = make_double2(D.vy[i], D.vz[i]);
    } else srec[p].z = z;
}

// occupancy of every cell from an uploaded link list
__global__ void k_cell_count(DevP D) {
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= D.ncell) return;
    int cnt = 0;
    for (int j = D.celllist[D.n + c + 1]; j; j = D.celllist[j]) cnt++;
    D.cell_cnt[c] = cnt;
}

// cell-sorted SoA copies for the full-system prediction pass
// parts: bit 0 = positions, types and index maps, bit 1 = velocities (sdmd_cuda_prepare sorts them once their upload lands),
// bit 2 = also bump the candidate-cache epoch
__global__ void k_cell_permute(DevP D, const int *sorted_atom, SRec *srec, float4 *sloc, int *scell,
                               int parts) {
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= D.n) return;
    int j = sorted_atom[p];
    double2 *rp = reinterpret_cast<double2 *>(&srec[p]);
    if ((parts & 4) && p == 0) D.ctl->cc_floor = ++D.ctl->cc_epoch;   // new link list: every cached candidate list is void
    if (parts & 1) {
        D.spos[j] = p;
        double x = D.x[j], y = D.y[j], z = D.z[j];
        int t = D.type[j];
        int cx = D.cell_idx[3 * j], cy = D.cell_idx[3 * j + 1], cz = D.cell_idx[3 * j + 2];
        // cell-local single-precision copy for the neighbour filter (see d_r0_filter)
        sloc[p] = make_float4((float)(x - cx * D.cellsize[0]), (float)(y - cy * D.cellsize[1]), (float)(z - cz * D.cellsize[2]),
                              __int_as_float(j));
        scell[p] = (cz * D.cellnum[1] + cy) * D.cellnum[0] + cx;
        rp[0] = make_double2(x, y);
        *reinterpret_cast<int4 *>(rp + 3) = make_int4(t, j, (int)((unsigned)cx | ((unsigned)cy << 16)), cz);
        if (parts & 2) {
            rp[1] = make_double2(z, D.vx[j]);
            rp[2] = make_double2(D.vy[j], D.vz[j]);
        } else srec[p].z = z;
    } else if (parts & 2) {
        srec[p].vx = D.vx[j];
        rp[2] = make_double2(D.vy[j], D.vz[j]);
    }
}

// ===================================================================================================
// kernels: full-system prediction (K2)
// ===================================================================================================
__device__ __forceinline__ SlotSrc as_src(const SlotBuf &b) {
    SlotSrc S;
    S.time = b.time; S.pot = b.pot; S.partner = b.partner; S.pcounter = b.pcounter; S.type = b.type; S.sub = b.sub;
    return S;
}

// Full-system Predict (SGThread.c:47-52) = the recursion slot(i) = F(i, slot(j < i)): atom i copies the slot of a
// lower-index partner that points back at it (CheckDuplication, TimePrediction.c:1101-1117).  It is solved exactly:
//   round 0   every atom folds its candidates with no duplicate source and also records the value of its
//             second-best candidate (m2);
//   resolve   an atom pointed at by exactly ONE lower-index collision slot j is settled without re-evaluation:
//               j is not my own best and t_j > my best            -> my round-0 slot stands
//               j is my own best (same type) and t_j < m2          -> my slot is a copy of j's
//             (proof sketch in DESIGN.md); anything else — several pointers, other event types, near ties — is
//             queued for re-evaluation;
//   rounds    queued atoms are re-evaluated against the current slots of lower-index atoms until nothing changes;
//             a changed slot queues the higher-index atoms it pointed / points at.
// The stale fields of a slot (potential / subEventType / partnerCounter survive ResetTarget) always start from the
// live arrays, which this pass never writes until the final commit.
#ifndef K2F_MINBLOCKS
#define K2F_MINBLOCKS 5
#endif
__global__ void __launch_bounds__(256, K2F_MINBLOCKS) k_predict_all_filter(DevP D) {
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p < D.n) d_r0_filter(D, p);
}

#ifndef QSPREAD
#define QSPREAD 8
#endif
#ifndef K2_MINBLOCKS
#define K2_MINBLOCKS 6
#endif
#ifndef K2_MINBLOCKS_LEAN
#define K2_MINBLOCKS_LEAN 5   // 96 registers, no spills (6 blocks: 80 registers with spills, measured 3 % slower)
#endif
// round 0: every atom, no duplicate source (one thread per cell-sorted atom)
template <bool LEAN>
__global__ void __launch_bounds__(128, LEAN ? K2_MINBLOCKS_LEAN : K2_MINBLOCKS) k_predict_all_round(DevP D, SlotBuf dst, int *npair, double *second, double *bp_time, int *bp_j,
                                                           int *pcount, int *pwho, i64 g_base) {
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= D.n) return;
    int i = D.sorted_atom[p];
    Slot cur = d_reset_slot(D.ev_pot[i], D.ev_sub[i], D.ev_pcounter[i]);
    double g = 0;
    if (D.thermostat_on && (D.no_cons || !D.is_cons[i])) {
        i64 gi = g_base + (D.no_cons ? i - 1 : D.ther_rank[i]);   // g_base < g_cap (host reduces it), rank < n <= g_cap
        g = D.gauss[gi >= D.g_cap ? gi - D.g_cap : gi];
    }
    int np = 0;
    double m2 = SD_INF;
    PairBest pb;
    pb.bp_time = SD_INF; pb.sp_time = SD_INF; pb.bp_j = 0;
    SlotSrc none = {};
    cur = d_predict_atom<true, true, LEAN>(D, none, i, cur, g, 0, &np, &m2, &pb, D.surv ? p : -1);
    dst.time[i] = cur.time; dst.pot[i] = cur.pot; dst.partner[i] = cur.partner; dst.pcounter[i] = cur.pcounter;
    dst.type[i] = (uint8_t)cur.type; dst.sub[i] = (uint8_t)cur.sub;
    npair[i] = np;
    // a partner-less winner is settled against its best PAIR candidate instead (see k_predict_all_resolve)
    bool pairwin = cur.partner > 0;
    second[i] = pairwin ? m2 : pb.sp_time;
    bp_time[i] = pb.bp_time;
    bp_j[i] = pb.bp_j;
    // who points at whom: count and remember one lower-index pointer per atom
    if (cur.partner > i) { atomicAdd(&pcount[cur.partner], 1); pwho[cur.partner] = i; }
}

// One re-evaluation round over the queued atoms, `src` -> `dst`.  The paths diverge and are chains of dependent DRAM reads, so
// each queued atom gets a group of QSPREAD lanes: lane 0 evaluates, the others first pull everything lane 0 is about to read one
// dependent load at a time into this SM's L1 (survivor words -> 64-byte records -> the partners' slots; the atom's own state).
__device__ __forceinline__ void d_requeue_round(const DevP &D, const SlotBuf &src, const SlotBuf &dst, const int *queue, int total, int *npair,
                                                unsigned long long *pair_sum, i64 g_base, int gtid, int gsz) {
    const int sub = gtid % QSPREAD;
    gtid /= QSPREAD; gsz /= QSPREAD;
    for (int p = gtid; p < total; p += gsz) {
        int i = queue[p];
        double acc = 0;
        if (D.surv) {
            int ps = D.spos[i];
            int ns = D.surv_n[ps];
            for (int s = sub; s < ns; s += QSPREAD) {
                unsigned w = D.surv[(size_t)ps * D.surv_cap + s];
                const SRec *r = &D.srec[w & 0x7ffffffu];
                int j = r->atom;
                acc += r->x + r->vy + (double)src.partner[j] + (double)src.type[j] + (double)src.pcounter[j] + src.time[j] + src.pot[j] +
                       (double)src.sub[j] + (double)D.ev_counter[j];
            }
            if (sub == 1) acc += (double)D.surv_np[ps] + D.srec[ps].x + D.srec[ps].vy;
        }
        if (sub == 2) acc += D.ev_pot[i] + (double)D.ev_sub[i] + (double)D.ev_pcounter[i] + (double)D.ev_counter[i];
        if (sub == 3) acc += (double)D.cm_off[i] + (double)npair[i] + (double)D.is_cons[i];
        if (sub == 4 && D.thermostat_on) { i64 gi = g_base + (D.no_cons ? i - 1 : D.ther_rank[i]); acc += D.gauss[gi >= D.g_cap ? gi - D.g_cap : gi]; }
        if (sub != 0) {
            if (acc == 1.2345e-300) D.ctl->error_prints++;   // keeps the loads alive; never true
            continue;
        }
        Slot cur = d_reset_slot(D.ev_pot[i], D.ev_sub[i], D.ev_pcounter[i]);
        double g = 0;
        if (D.thermostat_on && (D.no_cons || !D.is_cons[i])) {
            i64 gi = g_base + (D.no_cons ? i - 1 : D.ther_rank[i]);
            g = D.gauss[gi >= D.g_cap ? gi - D.g_cap : gi];
        }
        int np = 0;
        cur = d_predict_atom<true, false, false>(D, as_src(src), i, cur, g, i, &np, nullptr, nullptr, D.surv ? D.spos[i] : -1);
        dst.time[i] = cur.time; dst.pot[i] = cur.pot; dst.partner[i] = cur.partner; dst.pcounter[i] = cur.pcounter;
        dst.type[i] = (uint8_t)cur.type; dst.sub[i] = (uint8_t)cur.sub;
        int old = npair[i];
        npair[i] = np;
        if (np != old) atomicAdd(pair_sum, (unsigned long long)(long long)(np - old));
    }
}

// The same round for bond-free systems whose survivor rows come from sdmd_k2_rows.cu (DevP::rows_mode): ONE WARP per queued atom.
// Lane s takes survivor s: record gather, the partner's slot (is it a CheckDuplication source?), the ordinary evaluation — all
// side by side; the results are parked in shared memory in visiting order (rank of the survivor word) and lane 0 replays the
// scan over them exactly like the sequential code does (copy where the partner's slot points back and undercuts the running best,
// TimePrediction.c:1101-1117, otherwise JobAssign).  A queued atom then costs ~6 dependent memory round trips instead of ~50.
struct RqRes {
    double t, pot, tj, dpot;
    long long cntj;
    int j, sub, dsub, flags;     // flags: 1 = the ordinary evaluation produced an event, 2 = the partner's slot points back at me
};
// G lanes per queued atom, CAP = 2 G entries per survivor row (8 lanes for 16-entry rows, 16 for 32-entry rows): a queued atom is
// one chain of dependent round trips, so what counts is how many chains are in flight.
template <int G>
__device__ __forceinline__ void d_requeue_round_rows(const DevP &D, const SlotBuf &src, const SlotBuf &dst, const int *queue, int total, int *npair,
                                                     unsigned long long *pair_sum, i64 g_base, int gtid, int gsz, RqRes *res_warp, unsigned *w_warp) {
    constexpr int CAP = 2 * G;
    const int sub = gtid & (G - 1), gid = gtid / G, ng = gsz / G;
    const int goff = (gtid & 31) & ~(G - 1);                   // first lane of my group inside the warp
    const unsigned gmask = (G == 32 ? 0xffffffffu : ((1u << G) - 1u)) << goff;
    RqRes *res = res_warp + (goff / G) * CAP;
    unsigned *wl = w_warp + (goff / G) * CAP;
    for (int p = gid; p < total; p += ng) {
        const int i = queue[p];
        // everything addressed by the atom id alone travels together ...
        const int ps = D.spos[i];
        const double st_pot = D.ev_pot[i];
        const int st_sub = D.ev_sub[i];
        const i64 st_pc = D.ev_pcounter[i];
        const int np_old = npair[i];
        double g = 0;
        if (D.thermostat_on) { i64 gi = g_base + i - 1; g = D.gauss[gi >= D.g_cap ? gi - D.g_cap : gi]; }
        // ... then everything addressed by the sorted position
        const int ns = D.surv_n[ps];
        const int np0 = D.surv_np[ps];
        const unsigned *row = D.surv + (size_t)ps * D.surv_cap;
        const unsigned wraw0 = row[sub], wraw1 = row[sub + G];     // (CAP == surv_cap: inside the row)
        Self me = d_load_self_sorted(D, ps);
        if (ns < 0) {   // overflowed row: the generic serial path
            if (sub == 0) {
                Slot cur = d_reset_slot(st_pot, st_sub, st_pc);
                int np = 0;
                cur = d_predict_atom<true, false, false>(D, as_src(src), i, cur, g, i, &np, nullptr, nullptr, ps);
                dst.time[i] = cur.time; dst.pot[i] = cur.pot; dst.partner[i] = cur.partner; dst.pcounter[i] = cur.pcounter;
                dst.type[i] = (uint8_t)cur.type; dst.sub[i] = (uint8_t)cur.sub;
                npair[i] = np;
                if (np != np_old) atomicAdd(pair_sum, (unsigned long long)(long long)(np - np_old));
            }
            __syncwarp(gmask);
            continue;
        }
        wl[sub] = sub < ns ? wraw0 : 0xffffffffu;
        wl[sub + G] = sub + G < ns ? wraw1 : 0xffffffffu;
        __syncwarp(gmask);
        const bool seam = me.cx == 0 || me.cy == 0 || me.cz == 0 || me.cx == D.cellnum[0] - 1 || me.cy == D.cellnum[1] - 1 ||
                          me.cz == D.cellnum[2] - 1;
        for (int s = sub; s < ns; s += G) {
            const unsigned w = s == sub ? wraw0 : wraw1;
            int rank = 0;           // visiting order = ascending survivor word (all words of a row differ)
            for (int k = 0; k < ns; k++) rank += wl[k] < w;
            const double2 *rp = reinterpret_cast<const double2 *>(&D.srec[w & 0x7ffffffu]);
            const double2 r0 = __ldg(rp), r1 = __ldg(rp + 1), r2v = __ldg(rp + 2);
            const int2 ta = __ldg(reinterpret_cast<const int2 *>(rp + 3));
            const int j = ta.y;
            RqRes r;
            r.flags = 0; r.j = j; r.t = 0; r.pot = 0; r.sub = 0;
            // the partner's slot: all fields requested at once, used only when it points back at me
            const int s_partner = src.partner[j];
            const int s_type = src.type[j];
            const i64 s_pc = src.pcounter[j];
            r.tj = src.time[j]; r.dpot = src.pot[j]; r.dsub = src.sub[j];
            r.cntj = D.ev_counter[j];
            if (j < i && s_partner == i && s_type == SDMD_EV_COLI && s_pc == me.counter) r.flags = 2;
            double shx = 0, shy = 0, shz = 0;
            if (seam) {
                int ncx, ncy, ncz;
                d_neighbor_cell(D, me.cx, me.cy, me.cz, (int)(w >> 27), &ncx, &ncy, &ncz);
                shx = d_shift1(me.cx, ncx, D.cellnum[0], d_box(D, 0));
                shy = d_shift1(me.cy, ncy, D.cellnum[1], d_box(D, 1));
                shz = d_shift1(me.cz, ncz, D.cellnum[2], d_box(D, 2));
            }
            double t, pot;
            int sb;
            if (d_eval_pair(D, me, j, ta.x, r0.x, r0.y, r1.x, r1.y, r2v.x, r2v.y, shx, shy, shz, 0, me.type, &t, &pot, &sb)) {
                r.flags |= 1; r.t = t; r.pot = pot; r.sub = sb;
            }
            res[rank] = r;
        }
        __syncwarp(gmask);
        if (sub == 0) {
            Slot cur = d_reset_slot(st_pot, st_sub, st_pc);
            int np = np0;
            for (int s = 0; s < ns; s++) {
                const RqRes q = res[s];
                if ((q.flags & 2) && cur.time > q.tj) {     // CheckDuplication fires: the slot is copied, the pair is not evaluated
                    cur.partner = q.j; cur.pot = q.dpot; cur.sub = q.dsub; cur.time = q.tj; cur.type = SDMD_EV_COLI; cur.pcounter = q.cntj;
                    np--;
                    continue;
                }
                if ((q.flags & 1) && q.t > 0 && cur.time > q.t) {   // JobAssign, TimePrediction.c:1140-1158
                    cur.pot = q.pot; cur.sub = q.sub; cur.time = q.t; cur.type = SDMD_EV_COLI; cur.partner = q.j; cur.pcounter = q.cntj;
                }
            }
            d_pcc_time(D, me, cur);
            if (D.thermostat_on) d_thermostat_time(D, cur, g);
            dst.time[i] = cur.time; dst.pot[i] = cur.pot; dst.partner[i] = cur.partner; dst.pcounter[i] = cur.pcounter;
            dst.type[i] = (uint8_t)cur.type; dst.sub[i] = (uint8_t)cur.sub;
            npair[i] = np;
            if (np != np_old) atomicAdd(pair_sum, (unsigned long long)(long long)(np - np_old));
        }
        __syncwarp(gmask);
    }
}

// scan slot (0..26) of neighbour cell offset (dx,dy,dz), inverse of the DMD.c:118-120 tables
static __constant__ int c_kinv[27] = {14, 11, 23, 12, 10, 17, 22, 16, 24, 13, 9, 19, 8, 0, 1, 20, 2, 6, 21, 15, 25, 18, 3, 5, 26, 4, 7};
__device__ __forceinline__ int d_scan_slot(const DevP &D, int i, int a) {
    int k = 0;
#pragma unroll
    for (int d = 0; d < 3; d++) {
        int diff = D.cell_idx[3 * a + d] - D.cell_idx[3 * i + d];
        int cn = D.cellnum[d];
        if (diff == cn - 1) diff = -1;          // wrapped below
        else if (diff == 1 - cn) diff = 1;      // wrapped above
        k += (diff + 1) * (d == 0 ? 1 : d == 1 ? 3 : 9);
    }
    return c_kinv[k];
}

// settle atoms with a single lower-index pointer, queue the ambiguous ones
__global__ void k_predict_all_resolve(DevP D, SlotBuf b, const int *pcount, const int *pwho, const double *second, const double *bp_time,
                                      const int *bp_j, int *queue, int *nqueue, int *npair, unsigned long long *pair_sum) {
    SDMD_PDL_ENTRY();
    int i = blockIdx.x * blockDim.x + threadIdx.x + 1;
    int copy = 0;
    int cnt = (i <= D.n) ? pcount[i] : 0;
    bool ambiguous = cnt > 0;
    if (cnt > 0) {
    if (cnt == 1) {
        int j = pwho[i];
        double tj = b.time[j];
        // only plain collision pairs with a live counter match take the shortcut
        if (b.type[j] == SDMD_EV_COLI && b.pcounter[j] == D.ev_counter[i]) {
            double m1 = b.time[i];
            bool j_is_best = (b.partner[i] == j && b.type[i] == SDMD_EV_COLI);
            if (b.partner[i] <= 0) {
                // partner-less winner W (PCC / thermostat, visited after the whole cell scan): it still wins iff t_j > m1.
                // Its partnerCounter is the counter of the last pair candidate that was a running best = the best pair
                // candidate q of the sections in front of PCC.
                int q = bp_j[i];
                if (tj > m1 && q > 0) {
                    if (q != j) {
                        // q visited before j and t_j > t_q: the copy cannot fire, nothing changes
                        if (tj > bp_time[i]) {
                            int kq = d_scan_slot(D, i, q), kj = d_scan_slot(D, i, j);
                            if (kq < kj || (kq == kj && q > j)) ambiguous = false;
                        }
                    } else if (tj < second[i]) {
                        // j is that best pair candidate: the copy fires at j (T_before >= second-best pair > t_j), the
                        // record still carries counter(j), W still wins; only the evaluation of j is saved
                        copy = 2;
                        ambiguous = false;
                    }
                }
            } else if (!j_is_best) {
                // my own winner q must be a collision candidate visited BEFORE j: then the running best at j is
                // already <= m1 < t_j, CheckDuplication cannot fire and both slot and evaluation count stand
                int q = b.partner[i];
                if (tj > m1 && q > 0 && b.type[i] == SDMD_EV_COLI) {
                    int kq = d_scan_slot(D, i, q), kj = d_scan_slot(D, i, j);
                    if (kq < kj || (kq == kj && q > j)) ambiguous = false;   // inside a cell: descending atom id
                }
            }
            else if (tj < second[i]) { copy = 1; ambiguous = false; }
        } else if (b.pcounter[j] != D.ev_counter[i]) ambiguous = false;  // CheckDuplication can never fire
    }
    }
    unsigned long long v = 0;
    if (i <= D.n) {
        int np = npair[i];
        if (copy) {
            // the duplicate is copied, not evaluated (TimePrediction.c:87-90).  The source slot j is never itself a
            // copy target (its own winner is i > j), so applying the copy here is race free.
            np -= 1;
            npair[i] = np;
            if (copy == 1) {
                int j = pwho[i];
                b.partner[i] = j; b.pot[i] = b.pot[j]; b.sub[i] = b.sub[j]; b.time[i] = b.time[j]; b.type[i] = SDMD_EV_COLI;
                b.pcounter[i] = D.ev_counter[j];
            }
        }
        if (ambiguous) queue[atomicAdd(nqueue, 1)] = i;
        v = (unsigned long long)np;
    }
    for (int o = 16; o; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    __shared__ unsigned long long wsum[8];
    if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long t = 0;
        for (int w = 0; w < 8; w++) t += wsum[w];
        if (t) atomicAdd(pair_sum, t);
    }
}

// commit round r: move re-evaluated slots into the current buffer, queue dependants of changed slots
__device__ __forceinline__ void d_requeue_commit(const SlotBuf &cur, const SlotBuf &nxt, const int *queue, int nqueue, int *queue_next,
                                                 int *nqueue_next, int gtid, int gsz) {
    for (int p = gtid; p < nqueue; p += gsz) {
        int i = queue[p];
        // all twelve loads up front (no short-circuit: each would be its own DRAM round trip)
        long long ct = __double_as_longlong(cur.time[i]), nt = __double_as_longlong(nxt.time[i]);
        long long cp = __double_as_longlong(cur.pot[i]), np = __double_as_longlong(nxt.pot[i]);
        int oldp = cur.partner[i], newp = nxt.partner[i];
        i64 cc = cur.pcounter[i], nc = nxt.pcounter[i];
        uint8_t cty = cur.type[i], nty = nxt.type[i], csb = cur.sub[i], nsb = nxt.sub[i];
        bool same = (ct == nt) & (cp == np) & (oldp == newp) & (cc == nc) & (cty == nty) & (csb == nsb);
        if (same) continue;
        cur.time[i] = __longlong_as_double(nt); cur.pot[i] = __longlong_as_double(np); cur.partner[i] = newp; cur.pcounter[i] = nc;
        cur.type[i] = nty; cur.sub[i] = nsb;
        if (oldp > i) queue_next[atomicAdd(nqueue_next, 1)] = oldp;
        if (newp > i && newp != oldp) queue_next[atomicAdd(nqueue_next, 1)] = newp;
    }
}

// The re-evaluation rounds run to their fixed point on the device: one persistent cooperative grid, a grid-wide barrier
// between evaluating a round and committing it, no look from the host in between.  Afterwards thread 0 books the pass into the
// control block: evaluation count (summed by k_predict_all_resolve and the rounds), thermostat draws, queue statistics.
#define FIX_NQ 16   // queue-length slots, one per round; the slots are recycled when a dependency chain is longer than that
// `Dq` is `D` with the control block replaced by a scratch one: a re-evaluated atom repeats FindPair calls round 0 already
// made for it (the pairs a copy skips never warn: the partner predicted an event for them), so the warning / error-print
// counters of the rounds go to the scratch block and only its error flag is carried over.
__global__ void __launch_bounds__(128, K2_MINBLOCKS) k_predict_all_fixpoint(DevP D, DevP Dq, SlotBuf B, SlotBuf C, int *qa, int *qb, int *nqueue, int *npair,
                                                                             unsigned long long *pair_sum, i64 g_base, i64 draws, int qcap) {
    cooperative_groups::grid_group grid = cooperative_groups::this_grid();
    const int gtid = blockIdx.x * blockDim.x + threadIdx.x, gsz = gridDim.x * blockDim.x;
    __shared__ RqRes rq_res[4][64];
    __shared__ unsigned rq_w[4][64];
    int slot = 0, rounds = 0, first = -1;
    bool overflow = false;
    for (;;) {
        const int nq = nqueue[slot];
        if (first < 0) first = nq;
        if (nq == 0) break;
        if (nq > qcap) { overflow = true; break; }
        const long long c0 = clock64();
        if (D.rows_mode && D.surv_cap == 16) d_requeue_round_rows<8>(Dq, B, C, qa, nq, npair, pair_sum, g_base, gtid, gsz, rq_res[threadIdx.x >> 5], rq_w[threadIdx.x >> 5]);
        else if (D.rows_mode && D.surv_cap == 32) d_requeue_round_rows<16>(Dq, B, C, qa, nq, npair, pair_sum, g_base, gtid, gsz, rq_res[threadIdx.x >> 5], rq_w[threadIdx.x >> 5]);
        else d_requeue_round(Dq, B, C, qa, nq, npair, pair_sum, g_base, gtid, gsz);
        const long long c1 = clock64();
        grid.sync();
        const long long c2 = clock64();
        d_requeue_commit(B, C, qa, nq, qb, nqueue + slot + 1, gtid, gsz);
        const long long c3 = clock64();
        grid.sync();
        if (gtid == 0) {    // block 0's view of a round: evaluate, barrier, commit, barrier (Replica.phase_cycles: x10/x11 and the loop's slots 8/9 are reused)
            Ctl *c = D.ctl;
            c->phase_cycles[8] += c1 - c0; c->phase_cycles[9] += c2 - c1; c->phase_cycles[10] += c3 - c2; c->phase_cycles[11] += clock64() - c3;
        }
        int *t = qa; qa = qb; qb = t;
        slot++; rounds++;
        if (slot + 1 >= FIX_NQ) {   // very long dependency chain: recycle the slots, keep the last queue
            const int last = nqueue[slot];
            grid.sync();
            if (gtid == 0) { nqueue[0] = last; for (int k = 1; k < FIX_NQ; k++) nqueue[k] = 0; }
            grid.sync();
            slot = 0;
        }
    }
    if (gtid == 0) {
        Ctl *c = D.ctl;
        c->pair_predictions += (i64)*pair_sum;
        c->g_pos += draws;
        c->pa_passes++; c->pa_rounds += rounds; c->pa_requeued += first;
        if (overflow) c->error = SDMD_ERR_STATE;
        if (Dq.ctl->error && !c->error) c->error = Dq.ctl->error;
        Dq.ctl->error = 0;
    }
}

// ===================================================================================================
// kernels: tournament tree (K5), advance (K4)
// ===================================================================================================
__global__ void k_cbt_leaves(DevP D, const int *leaf_atom) {
    SDMD_PDL_ENTRY();
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= D.n) return;
    int a = leaf_atom[k];
    int pos = (D.n == 1) ? 1 : D.n + k;
    D.cbt_node[pos] = a;
    D.cbt_leaf[a] = pos;
}
// internal node: winner = t[left] < t[right] ? left : right (CBT.c:62-67)
// bottom-up tournament inside one subtree per block: root f0 = 2^L + blockIdx.x, internal nodes of relative depth
// dmax..0, one __syncthreads per level (a block's writes are visible to its own threads after the barrier)
__global__ void __launch_bounds__(256) k_cbt_subtrees(DevP D, int L, int dmax) {
    SDMD_PDL_ENTRY();
    int f0 = (1 << L) + blockIdx.x;
    if (f0 >= D.n) return;
    for (int d = dmax; d >= 0; d--) {
        int base = f0 << d;
        int cnt = 1 << d;
        for (int t = threadIdx.x; t < cnt; t += blockDim.x) {
            int f = base + t;
            if (f < D.n) {
                int l = D.cbt_node[2 * f], r = D.cbt_node[2 * f + 1];
                D.cbt_node[f] = (D.ev_time[l] < D.ev_time[r]) ? l : r;
            }
        }
        __syncthreads();
    }
}
// the levels above the subtree roots (nodes 1 .. 2^L - 1), one block
__global__ void __launch_bounds__(256) k_cbt_top(DevP D, int L) {
    SDMD_PDL_ENTRY();
    for (int lv = L - 1; lv >= 0; lv--) {
        int base = 1 << lv;
        for (int t = threadIdx.x; t < base; t += blockDim.x) {
            int f = base + t;
            if (f < D.n) {
                int l = D.cbt_node[2 * f], r = D.cbt_node[2 * f + 1];
                D.cbt_node[f] = (D.ev_time[l] < D.ev_time[r]) ? l : r;
            }
        }
        __syncthreads();
    }
}

// UpdateData + TimeForward fused, ToolFunctions.c:23-36: 56 B read + 32 B written per atom
__global__ void __launch_bounds__(256) k_advance(DevP D, double dt_arg, int use_ctl) {
    if (use_ctl && D.ctl->stop) return;
    double dt = use_ctl ? D.ctl->dt : dt_arg;
    int stride = gridDim.x * blockDim.x;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x + 1; i <= D.n; i += stride) {
        double x = D.x[i], y = D.y[i], z = D.z[i];
        double vx = D.vx[i], vy = D.vy[i], vz = D.vz[i];
        double t = D.ev_time[i];
        D.x[i] = x + vx * dt;
        D.y[i] = y + vy * dt;
        D.z[i] = z + vz * dt;
        D.ev_time[i] = t - dt;
    }
}

// ===================================================================================================
// kernels: event loop
// ===================================================================================================
// loop head: stop conditions + AssignJob (SGThread.c:75,237-259); single thread
// Small systems: the whole loop in ONE persistent CTA (one replica per CTA); no host round trip per event.
// UpdateData + TimeForward (ToolFunctions.c:23-36) by one thread block.  All loads of up to four atoms per thread are issued
// before the first store: written as `x[i] = x[i] + vx[i] * dt` per component the compiler must keep every load behind the
// previous store (the arrays may alias), which made each atom four dependent L2 round trips — 4.6 us per 1000 atoms.
__device__ __forceinline__ void d_advance_cta(const DevP &D, double dt) {
    const int n = D.n, nt = blockDim.x;
    double *x = D.x, *y = D.y, *z = D.z, *t = D.ev_time;
    const double *vx = D.vx, *vy = D.vy, *vz = D.vz;
    int i = threadIdx.x + 1;
    for (; i + 3 * nt <= n; i += 4 * nt) {
        double xa[4], ya[4], za[4], ua[4], va[4], wa[4], ta[4];
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const int k = i + u * nt;
            xa[u] = x[k]; ya[u] = y[k]; za[u] = z[k]; ua[u] = vx[k]; va[u] = vy[k]; wa[u] = vz[k]; ta[u] = t[k];
        }
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const int k = i + u * nt;
            x[k] = xa[u] + ua[u] * dt; y[k] = ya[u] + va[u] * dt; z[k] = za[u] + wa[u] * dt; t[k] = ta[u] - dt;
        }
    }
    for (; i <= n; i += nt) {
        const double x0 = x[i], y0 = y[i], z0 = z[i], u0 = vx[i], v0 = vy[i], w0 = vz[i], t0 = t[i];
        x[i] = x0 + u0 * dt; y[i] = y0 + v0 * dt; z[i] = z0 + w0 * dt; t[i] = t0 - dt;
    }
}

__global__ void __launch_bounds__(256, 1) k_run_cta(DevP Dg, double t_until, i64 iter_limit, sdmd_trace_rec *trace, i64 trace_cap) {
    __shared__ PredShared sh;
    __shared__ Ctl s_ctl;      // the loop state lives in shared memory while the kernel runs (written back on exit)
    __shared__ double s_dt;
    __shared__ int s_stop;
    // the replica descriptor goes to shared memory once: a per-thread `DevP D = Dg` is a 1 KB local-memory copy per thread
    // (256 KB per block, more than L1 holds), and every D.field of the loop below would be a local load
    __shared__ DevP s_D;
    for (int k = threadIdx.x; k < (int)(sizeof(Ctl) / sizeof(int)); k += blockDim.x) ((int *)&s_ctl)[k] = ((const int *)Dg.ctl)[k];
    for (int k = threadIdx.x; k < (int)(sizeof(DevP) / sizeof(int)); k += blockDim.x) ((int *)&s_D)[k] = ((const int *)&Dg)[k];
    __syncthreads();
    if (threadIdx.x == 0) s_D.ctl = &s_ctl;
    __syncthreads();
    const DevP &D = s_D;
    for (;;) {
        long long h0 = clock64();
        if (threadIdx.x == 0) {
            d_loop_head(D, t_until, iter_limit);
            s_dt = D.ctl->dt;
            s_stop = D.ctl->stop;
            if (!s_stop) d_publish_job(D, sh);
        }
        __syncthreads();
        if (s_stop) break;
        long long h1 = d_clock_after(&s_stop);
        double dt = s_dt;
        d_advance_cta(D, dt);
        __syncthreads();
        if (threadIdx.x == 0) { D.ctl->phase_cycles[0] += h1 - h0; D.ctl->phase_cycles[1] += d_clock_after(&s_stop) - h1; }
        d_loop_body_cta(D, sh, t_until, trace, trace_cap);
    }
    __syncthreads();
    for (int k = threadIdx.x; k < (int)(sizeof(Ctl) / sizeof(int)); k += blockDim.x) ((int *)Dg.ctl)[k] = ((const int *)&s_ctl)[k];
}

// Many replicas per GPU (REMD ensembles, parameter sweeps): one persistent CTA per replica, all in one launch.
// Each block copies its replica's descriptor into shared memory and runs the same loop as k_run_cta.
struct MultiArg {
    DevP P;
    double t_until;
    i64 iter_limit;
};
// The same loop for mid-size systems (about 10^3 .. 10^4 atoms) on one thread-block CLUSTER: block 0 is the event block of
// k_run_cta, the other blocks of the cluster only help with UpdateData/TimeForward, whose cost in one block is one SM's L2
// bandwidth (~1.1 us per 1000 atoms).  Two cluster barriers per iteration (~380 cycles each, release/acquire at cluster scope,
// which also drops the L1 lines the other SMs have overwritten): after AssignJob has published dt in block 0's shared memory
// (read by the helpers through distributed shared memory), and after the advance.
__global__ void __launch_bounds__(256, 1) k_run_cluster(DevP Dg, double t_until, i64 iter_limit, sdmd_trace_rec *trace, i64 trace_cap) {
    namespace cg = cooperative_groups;
    cg::cluster_group cluster = cg::this_cluster();
    const unsigned rank = cluster.block_rank(), nblk = cluster.num_blocks();
    __shared__ PredShared sh;
    __shared__ Ctl s_ctl;
    __shared__ double s_dt;
    __shared__ int s_stop;
    __shared__ DevP s_D;       // see k_run_cta
    if (rank == 0)
        for (int k = threadIdx.x; k < (int)(sizeof(Ctl) / sizeof(int)); k += blockDim.x) ((int *)&s_ctl)[k] = ((const int *)Dg.ctl)[k];
    for (int k = threadIdx.x; k < (int)(sizeof(DevP) / sizeof(int)); k += blockDim.x) ((int *)&s_D)[k] = ((const int *)&Dg)[k];
    __syncthreads();
    if (threadIdx.x == 0) s_D.ctl = &s_ctl;
    __syncthreads();
    const DevP &D = s_D;
    const double *r_dt = cluster.map_shared_rank(&s_dt, 0);
    const int *r_stop = cluster.map_shared_rank(&s_stop, 0);
    for (;;) {
        long long h0 = clock64();
        if (rank == 0 && threadIdx.x == 0) {
            d_loop_head(D, t_until, iter_limit);
            s_dt = D.ctl->dt;
            s_stop = D.ctl->stop;
            if (!s_stop) d_publish_job(D, sh);
        }
        cluster.sync();
        long long h1 = clock64();
        const int stop = *r_stop;
        const double dt = *r_dt;
        if (stop) break;
        {
            // this block's share of the atoms, loads of four atoms per thread before the first store (see d_advance_cta)
            const int n = D.n, nt = (int)(blockDim.x * nblk);
            double *x = D.x, *y = D.y, *z = D.z, *t = D.ev_time;
            const double *vx = D.vx, *vy = D.vy, *vz = D.vz;
            int i = (int)(rank * blockDim.x + threadIdx.x) + 1;
            for (; i + 3 * nt <= n; i += 4 * nt) {
                double xa[4], ya[4], za[4], ua[4], va[4], wa[4], ta[4];
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    const int k = i + u * nt;
                    xa[u] = x[k]; ya[u] = y[k]; za[u] = z[k]; ua[u] = vx[k]; va[u] = vy[k]; wa[u] = vz[k]; ta[u] = t[k];
                }
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    const int k = i + u * nt;
                    x[k] = xa[u] + ua[u] * dt; y[k] = ya[u] + va[u] * dt; z[k] = za[u] + wa[u] * dt; t[k] = ta[u] - dt;
                }
            }
            for (; i <= n; i += nt) {
                const double x0 = x[i], y0 = y[i], z0 = z[i], u0 = vx[i], v0 = vy[i], w0 = vz[i], t0 = t[i];
                x[i] = x0 + u0 * dt; y[i] = y0 + v0 * dt; z[i] = z0 + w0 * dt; t[i] = t0 - dt;
            }
        }
        cluster.sync();
        if (rank == 0) {
            if (threadIdx.x == 0) { D.ctl->phase_cycles[0] += h1 - h0; D.ctl->phase_cycles[1] += clock64() - h1; }
            d_loop_body_cta(D, sh, t_until, trace, trace_cap);
        }
    }
    cluster.sync();   // block 0's shared memory stays alive until every helper has read the stop flag
    if (rank == 0) {
        __syncthreads();
        for (int k = threadIdx.x; k < (int)(sizeof(Ctl) / sizeof(int)); k += blockDim.x) ((int *)Dg.ctl)[k] = ((const int *)&s_ctl)[k];
    }
}

template <int NT, int MINB>
__global__ void __launch_bounds__(NT, MINB) k_run_cta_multi(const MultiArg *args) {
    __shared__ PredShared sh;
    __shared__ MultiArg sA;
    __shared__ double s_dt;
    __shared__ int s_stop;
    {
        const int *src = (const int *)&args[blockIdx.x];
        int *dst = (int *)&sA;
        for (int k = threadIdx.x; k < (int)(sizeof(MultiArg) / sizeof(int)); k += blockDim.x) dst[k] = src[k];
    }
    __syncthreads();
    __shared__ Ctl s_ctl;
    Ctl *g_ctl = sA.P.ctl;
    for (int k = threadIdx.x; k < (int)(sizeof(Ctl) / sizeof(int)); k += blockDim.x) ((int *)&s_ctl)[k] = ((const int *)g_ctl)[k];
    __syncthreads();
    if (threadIdx.x == 0) sA.P.ctl = &s_ctl;
    __syncthreads();
    const DevP &D = sA.P;
    double t_until = sA.t_until;
    i64 iter_limit = sA.iter_limit;
    for (;;) {
        if (threadIdx.x == 0) {
            d_loop_head(D, t_until, iter_limit);
            s_dt = D.ctl->dt;
            s_stop = D.ctl->stop;
            if (!s_stop) d_publish_job(D, sh);
        }
        __syncthreads();
        if (s_stop) break;
        double dt = s_dt;
        d_advance_cta(D, dt);
        __syncthreads();
        d_loop_body_cta(D, sh, t_until, nullptr, 0);
    }
    __syncthreads();
    for (int k = threadIdx.x; k < (int)(sizeof(Ctl) / sizeof(int)); k += blockDim.x) ((int *)g_ctl)[k] = ((const int *)&s_ctl)[k];
}
__global__ void k_potential_energy_multi(const MultiArg *args, double *out);

// Large systems: per event one grid-wide advance (k_advance with use_ctl) and this single-CTA kernel,
// which finishes event k and performs AssignJob for event k+1.
__global__ void __launch_bounds__(256, 1) k_event_exec(DevP Dg, double t_until, i64 iter_limit, sdmd_trace_rec *trace, i64 trace_cap) {
    __shared__ PredShared sh;
    __shared__ Ctl s_ctl;      // loop state in shared memory while the block works, written back at the end
    if (Dg.ctl->stop) return;
    __shared__ DevP s_D;       // see k_run_cta
    for (int k = threadIdx.x; k < (int)(sizeof(Ctl) / sizeof(int)); k += blockDim.x) ((int *)&s_ctl)[k] = ((const int *)Dg.ctl)[k];
    for (int k = threadIdx.x; k < (int)(sizeof(DevP) / sizeof(int)); k += blockDim.x) ((int *)&s_D)[k] = ((const int *)&Dg)[k];
    __syncthreads();
    if (threadIdx.x == 0) s_D.ctl = &s_ctl;
    __syncthreads();
    const DevP &D = s_D;
    if (threadIdx.x == 0) d_publish_job(D, sh);
    __syncthreads();
    d_loop_body_cta(D, sh, t_until, trace, trace_cap);
    if (threadIdx.x == 0) d_loop_head(D, t_until, iter_limit);
    __syncthreads();
    for (int k = threadIdx.x; k < (int)(sizeof(Ctl) / sizeof(int)); k += blockDim.x) ((int *)Dg.ctl)[k] = ((const int *)&s_ctl)[k];
}
__global__ void k_loop_head(DevP D, double t_until, i64 iter_limit) { d_loop_head(D, t_until, iter_limit); }

// ---- overlapped sweep -----------------------------------------------------------------------------------------
// One launch per loop iteration.  Block 0 is the event block (HazardCheck, DoEvent, Predict, UpdateCBT and the next
// AssignJob); all other blocks are the UpdateData/TimeForward sweep of the SAME iteration, reading buffer lz_* and
// writing buffer x/y/z/ev_time.  The event block does not wait for the sweep: it writes the two atoms of the event into
// the new buffer itself and reads everybody else through d_rx..d_rt (see DevP).  The sweep skips those two atoms.
__device__ __forceinline__ void d_publish_next_job(const DevP &D, int slot) {
    Ctl *c = D.ctl;
    Ctl::Job j;
    j.dt = c->dt; j.seq = c->iterations; j.target = c->target; j.partner = c->partner; j.go = c->stop ? 0 : 1; j.pad_ = 0;
    c->job[slot] = j;
}
__global__ void k_loop_head_job(DevP D, double t_until, i64 iter_limit, int slot) {
    d_loop_head(D, t_until, iter_limit);
    d_publish_next_job(D, slot);
}
__global__ void __launch_bounds__(256, 2) k_sweep_event(DevP Dg, int slot, i64 seq, double t_until, i64 iter_limit, sdmd_trace_rec *trace, i64 trace_cap) {
    __shared__ lz1::PredShared sh;
    // programmatic dependent launch: the blocks of the NEXT iteration's launch may already be resident while this one
    // finishes; nothing the previous launch produced is read before this point
    asm volatile("griddepcontrol.wait;" ::: "memory");
    if (blockIdx.x == 0) asm volatile("griddepcontrol.launch_dependents;");
    Ctl::Job job = Dg.ctl->job[slot];
    if (job.seq != seq) job.go = 0;   // a slot left over from before a stop: this launch has nothing to do
    if (blockIdx.x != 0) {
        if (!job.go) return;
        const double dt = job.dt;
        const int a = job.target, b = job.partner, n = Dg.n;
        const int stride = (gridDim.x - 1) * blockDim.x;
        // two atoms per trip: fourteen independent loads in flight per thread
        for (int i = (blockIdx.x - 1) * blockDim.x + threadIdx.x + 1; i <= n; i += 2 * stride) {
            int i2 = i + stride;
            bool ok2 = i2 <= n;
            // streaming (evict-first) accesses: the sweep must not push the event block's working set out of L2
            double x = __ldcs(&Dg.lz_x[i]), y = __ldcs(&Dg.lz_y[i]), z = __ldcs(&Dg.lz_z[i]), t = __ldcs(&Dg.lz_t[i]);
            double vx = __ldcs(&Dg.vx[i]), vy = __ldcs(&Dg.vy[i]), vz = __ldcs(&Dg.vz[i]);
            double x2 = 0, y2 = 0, z2 = 0, vx2 = 0, vy2 = 0, vz2 = 0, t2 = 0;
            if (ok2) {
                x2 = __ldcs(&Dg.lz_x[i2]); y2 = __ldcs(&Dg.lz_y[i2]); z2 = __ldcs(&Dg.lz_z[i2]); t2 = __ldcs(&Dg.lz_t[i2]);
                vx2 = __ldcs(&Dg.vx[i2]); vy2 = __ldcs(&Dg.vy[i2]); vz2 = __ldcs(&Dg.vz[i2]);
            }
            if (i != a && i != b) {
                __stcs(&Dg.x[i], x + vx * dt); __stcs(&Dg.y[i], y + vy * dt); __stcs(&Dg.z[i], z + vz * dt); __stcs(&Dg.ev_time[i], t - dt);
            }
            if (ok2 && i2 != a && i2 != b) {
                __stcs(&Dg.x[i2], x2 + vx2 * dt); __stcs(&Dg.y[i2], y2 + vy2 * dt); __stcs(&Dg.z[i2], z2 + vz2 * dt); __stcs(&Dg.ev_time[i2], t2 - dt);
            }
        }
        return;
    }
    if (!job.go) return;
    long long k0 = clock64();
    // the loop state lives in shared memory while the block works (thread 0 touches it dozens of times per iteration);
    // meanwhile threads 32.. fetch what thread 0 needs next, so the pick-up is one round trip to memory, not four
    __shared__ Ctl s_ctl;
    __shared__ int s_type0;
    for (int k = threadIdx.x; k < (int)(sizeof(Ctl) / sizeof(int)); k += blockDim.x) ((int *)&s_ctl)[k] = ((const int *)Dg.ctl)[k];
    if (threadIdx.x >= 128 && threadIdx.x < 136) {
        int q = (threadIdx.x - 128) >> 2, f = (threadIdx.x - 128) & 3;
        int i = q ? job.partner : job.target;
        if (i > 0) {
            const double *src = f == 0 ? Dg.lz_x : f == 1 ? Dg.lz_y : f == 2 ? Dg.lz_z : Dg.lz_t;
            double v = src[i];
            // the event's own atoms move to the new buffer now (same arithmetic as the sweep)
            if (f == 0) v = v + Dg.vx[i] * job.dt; else if (f == 1) v = v + Dg.vy[i] * job.dt; else if (f == 2) v = v + Dg.vz[i] * job.dt; else v = v - job.dt;
            double *dst = f == 0 ? Dg.x : f == 1 ? Dg.y : f == 2 ? Dg.z : Dg.ev_time;
            dst[i] = v;
        }
    }
    if (threadIdx.x == 160) s_type0 = Dg.ev_type[job.target];
    __shared__ volatile double s_sink[64];
    __syncthreads();
    DevP D = Dg;
    D.ctl = &s_ctl;
    D.lz_dt = job.dt; D.lz_a = job.target; D.lz_b = job.partner;
    if (threadIdx.x == 0) {
        s_ctl.live = slot ^ 1;
        sh.atom[0] = job.target;
        sh.atom[1] = job.partner;
        sh.gathered = (s_type0 != SDMD_EV_PCC && s_type0 != SDMD_EV_HB);   // as d_publish_job
    }
    __syncthreads();
    long long k1 = lz1::d_clock_after(&sh.atom[0]);
    // idle warps warm this SM's L1 with what thread 0 is about to read one dependent load at a time ...
    if (threadIdx.x >= 192 && threadIdx.x < 224) {
        int l = threadIdx.x - 192, i = (l & 16) ? job.partner : job.target;
        double v = 0;
        if (i > 0) {
            switch (l & 15) {
                case 0: v = Dg.mass[i]; break;
                case 1: v = (double)Dg.ev_counter[i]; break;
                case 2: v = (double)Dg.ev_pcounter[i]; break;
                case 3: v = (double)Dg.is_cons[i]; break;
                case 4: v = (double)Dg.ev_partner[i]; break;
                case 5: v = (double)Dg.ev_sub[i]; break;
                case 6: v = Dg.vx[i]; break;
                case 7: v = Dg.vy[i]; break;
                case 8: v = Dg.vz[i]; break;
                case 9: v = Dg.ev_pot[i]; break;
                case 10: v = (double)Dg.cell_idx[3 * i]; break;
                case 11: v = (double)Dg.type[i]; break;
                case 12: v = Dg.gauss[Dg.ctl->g_pos % Dg.g_cap]; break;
                case 13: v = Dg.gauss[(Dg.ctl->g_pos + 4) % Dg.g_cap]; break;
                case 14: v = (double)Dg.cbt_leaf[i]; break;
                default: break;
            }
        }
        s_sink[32 + l] = v;
    }
    if (threadIdx.x >= 224) {
        // ... and walk the tournament-tree paths of the event's atoms now (leaf -> siblings -> their times),
        // so that UpdateCBT at the end of the iteration finds them in this SM's L1
        int lane = threadIdx.x - 224;
        double acc = 0;
        for (int q = 0; q < 2; q++) {
            int i = q ? job.partner : job.target;
            if (i <= 0) continue;
            int child = Dg.cbt_leaf[i] >> lane;
            if (child > 1) {
                int sib = Dg.cbt_node[child ^ 1];
                acc += (double)Dg.cbt_node[child >> 1] + Dg.lz_t[sib] + (double)Dg.ev_partner[sib];
            }
        }
        s_sink[lane] = acc;
    }
    lz1::d_loop_body_cta(D, sh, t_until, trace, trace_cap);
    if (threadIdx.x == 0) {
        long long k2 = clock64();
        lz1::d_loop_head(D, t_until, iter_limit);
        d_publish_next_job(D, slot ^ 1);
        s_ctl.phase_cycles[0] += k1 - k0;               // job pick-up: state to shared memory, own atoms to the new buffer
        s_ctl.phase_cycles[1] += clock64() - k2;        // next AssignJob
    }
    __syncthreads();
    for (int k = threadIdx.x; k < (int)(sizeof(Ctl) / sizeof(int)); k += blockDim.x) ((int *)Dg.ctl)[k] = ((const int *)&s_ctl)[k];
}

// The same overlapped iteration as k_sweep_event, but the whole loop in ONE cooperative launch (SGThread.c:75-152 never leaves
// the device): one block per SM, block 0 is the event block, the others sweep.  The two coordinate / event-time buffers swap
// roles every iteration (Dg describes iteration parity 0: read lz_*, write x/y/z/ev_time).  No grid-wide barrier is needed:
//   * a sweep thread only ever touches its own atoms (the same ones every iteration), so sweep blocks do not depend on each
//     other; what they wait for is the NEXT job — (dt, target, partner), published by the event block in the control block's
//     per-parity slot with the iteration number written last (release) — and that also orders the event block's own writes of
//     the two event atoms in front of the sweep that reads them;
//   * the event block of iteration n + 1 reads other atoms from the buffer sweep n wrote, so every sweep block counts itself
//     done (one release-add per block and iteration) and the event block checks the count before it goes on — by then the sweep
//     (14 us) has long finished, the check is one load.
// The event block keeps the loop state in shared memory for the whole run.  One block per SM lets the event role keep its
// working set in registers (no spills); the sweep role makes up for the smaller grid with several atoms in flight per thread.
#ifndef SWEEP_UNROLL
#define SWEEP_UNROLL 2   // atoms in flight per sweep thread: 2 keeps the sweep (~20 us) just under the event block without flooding the memory system (scripts/loop_variants.py)
#endif
__device__ __forceinline__ long long ld_acquire_i64(const long long *p) {
    long long v;
    asm volatile("ld.acquire.gpu.global.s64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_i64(long long *p, long long v) {
    asm volatile("st.release.gpu.global.s64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__global__ void __launch_bounds__(256, 1) k_run_sweep(DevP Dg, i64 seq0, unsigned long long *sweep_done, double t_until, i64 iter_limit,
                                                       sdmd_trace_rec *trace, i64 trace_cap) {
    __shared__ lz1::PredShared sh;
    __shared__ Ctl s_ctl;
    __shared__ int s_type0;
    __shared__ volatile double s_sink[64];
    __shared__ Ctl::Job s_job;
    const bool ev = blockIdx.x == 0;
    const unsigned long long nsweep = gridDim.x - 1;
    // The event role sees the replica through a descriptor in SHARED memory, one per buffer parity, built once: a per-thread copy
    // of the ~1 KB parameter block every iteration (what `DevP D = Dg; D.x = ...` compiles to) is 256 KB of local-memory traffic
    // per event.  Per iteration only dt / target / partner change (thread 0, three stores).
    __shared__ DevP s_D[2];
    if (ev) {
        for (int k = threadIdx.x; k < (int)(sizeof(Ctl) / sizeof(int)); k += blockDim.x) ((int *)&s_ctl)[k] = ((const int *)Dg.ctl)[k];
        for (int k = threadIdx.x; k < (int)(sizeof(DevP) / sizeof(int)); k += blockDim.x) {
            int w = ((const int *)&Dg)[k];
            ((int *)&s_D[0])[k] = w; ((int *)&s_D[1])[k] = w;
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            s_D[0].ctl = &s_ctl; s_D[1].ctl = &s_ctl;
            // parity 1 reads what parity 0 writes and the other way round
            s_D[1].x = const_cast<double *>(Dg.lz_x); s_D[1].y = const_cast<double *>(Dg.lz_y); s_D[1].z = const_cast<double *>(Dg.lz_z);
            s_D[1].ev_time = const_cast<double *>(Dg.lz_t);
            s_D[1].lz_x = Dg.x; s_D[1].lz_y = Dg.y; s_D[1].lz_z = Dg.z; s_D[1].lz_t = Dg.ev_time;
        }
        __syncthreads();
    }
    i64 seq = seq0;                                          // iteration number of the job this trip executes
    for (int r = 0;; r ^= 1) {
        // this iteration reads buffer r and writes buffer r ^ 1
        Ctl::Job job;
        if (ev) {
            job = s_ctl.job[r];
            // everything sweep (seq - 1) wrote must be there before this iteration reads it
            if (threadIdx.x == 0 && seq > seq0) {
                const unsigned long long want = nsweep * (unsigned long long)(seq - seq0);
                long long w0 = clock64();
                while ((unsigned long long)ld_acquire_i64(reinterpret_cast<const long long *>(sweep_done)) < want) __nanosleep(20);
                s_ctl.phase_cycles[9] += clock64() - w0;     // waiting for the sweep: ~one load unless the sweep is the slower side
            }
            __syncthreads();
        } else {
            static_assert(sizeof(Ctl::Job) == 32 && offsetof(Ctl::Job, seq) == 8, "job slot: dt, seq, target, partner, go, pad");
            if (threadIdx.x == 0) {
                const long long *slot = reinterpret_cast<const long long *>(&Dg.ctl->job[r]);
                while (ld_acquire_i64(slot + 1) != seq) __nanosleep(20);              // the iteration number is written last
                long long w[4] = {__ldcg(slot), seq, __ldcg(slot + 2), __ldcg(slot + 3)};
                memcpy(&s_job, w, 32);
            }
            __syncthreads();
            job = s_job;
        }
        if (!job.go) break;
        if (!ev) {
            const double dt = job.dt;
            const int a = job.target, b = job.partner, n = Dg.n;
            struct { const double *lz_x, *lz_y, *lz_z, *lz_t, *vx, *vy, *vz; double *x, *y, *z, *ev_time; } D;
            D.lz_x = r ? Dg.x : Dg.lz_x; D.lz_y = r ? Dg.y : Dg.lz_y; D.lz_z = r ? Dg.z : Dg.lz_z; D.lz_t = r ? Dg.ev_time : Dg.lz_t;
            D.x = r ? const_cast<double *>(Dg.lz_x) : Dg.x; D.y = r ? const_cast<double *>(Dg.lz_y) : Dg.y;
            D.z = r ? const_cast<double *>(Dg.lz_z) : Dg.z; D.ev_time = r ? const_cast<double *>(Dg.lz_t) : Dg.ev_time;
            D.vx = Dg.vx; D.vy = Dg.vy; D.vz = Dg.vz;
            const int stride = (gridDim.x - 1) * blockDim.x;
            for (int i0 = (blockIdx.x - 1) * blockDim.x + threadIdx.x + 1; i0 <= n; i0 += SWEEP_UNROLL * stride) {
                double x[SWEEP_UNROLL], y[SWEEP_UNROLL], z[SWEEP_UNROLL], t[SWEEP_UNROLL], vx[SWEEP_UNROLL], vy[SWEEP_UNROLL], vz[SWEEP_UNROLL];
#pragma unroll
                for (int q = 0; q < SWEEP_UNROLL; q++) {
                    int i = i0 + q * stride;
                    if (i <= n) {
                        // streaming (evict-first) accesses: the sweep must not push the event block's working set out of L2
                        x[q] = __ldcs(&D.lz_x[i]); y[q] = __ldcs(&D.lz_y[i]); z[q] = __ldcs(&D.lz_z[i]); t[q] = __ldcs(&D.lz_t[i]);
                        vx[q] = __ldcs(&D.vx[i]); vy[q] = __ldcs(&D.vy[i]); vz[q] = __ldcs(&D.vz[i]);
                    }
                }
#pragma unroll
                for (int q = 0; q < SWEEP_UNROLL; q++) {
                    int i = i0 + q * stride;
                    if (i <= n && i != a && i != b) {
                        __stcs(&D.x[i], x[q] + vx[q] * dt); __stcs(&D.y[i], y[q] + vy[q] * dt); __stcs(&D.z[i], z[q] + vz[q] * dt);
                        __stcs(&D.ev_time[i], t[q] - dt);
                    }
                }
            }
            __syncthreads();                                  // this block's share of sweep `seq` is written ...
            if (threadIdx.x == 0) { __threadfence(); atomicAdd(sweep_done, 1ULL); }   // ... and counted (release)
        } else {
            long long k0 = clock64();
            const DevP &D = s_D[r];
            if (threadIdx.x >= 128 && threadIdx.x < 136) {
                int q = (threadIdx.x - 128) >> 2, f = (threadIdx.x - 128) & 3;
                int i = q ? job.partner : job.target;
                if (i > 0) {
                    const double *src = f == 0 ? D.lz_x : f == 1 ? D.lz_y : f == 2 ? D.lz_z : D.lz_t;
                    double v = src[i];
                    // the event's own atoms move to the new buffer now (same arithmetic as the sweep)
                    if (f == 0) v = v + D.vx[i] * job.dt; else if (f == 1) v = v + D.vy[i] * job.dt; else if (f == 2) v = v + D.vz[i] * job.dt; else v = v - job.dt;
                    double *dst = f == 0 ? D.x : f == 1 ? D.y : f == 2 ? D.z : D.ev_time;
                    dst[i] = v;
                }
            }
            if (threadIdx.x == 160) s_type0 = D.ev_type[job.target];
            if (threadIdx.x == 0) { s_D[r].lz_dt = job.dt; s_D[r].lz_a = job.target; s_D[r].lz_b = job.partner; }
            __syncthreads();
            if (threadIdx.x == 0) {
                s_ctl.live = r ^ 1;
                sh.atom[0] = job.target;
                sh.atom[1] = job.partner;
                sh.gathered = (s_type0 != SDMD_EV_PCC && s_type0 != SDMD_EV_HB);   // as d_publish_job
            }
            __syncthreads();
            long long k1 = lz1::d_clock_after(&sh.atom[0]);
            // idle warps warm this SM's L1 with what thread 0 is about to read one dependent load at a time ...
            if (threadIdx.x >= 192 && threadIdx.x < 224) {
                int l = threadIdx.x - 192, i = (l & 16) ? job.partner : job.target;
                double v = 0;
                if (i > 0) {
                    switch (l & 15) {
                        case 0: v = D.mass[i]; break;
                        case 1: v = (double)D.ev_counter[i]; break;
                        case 2: v = (double)D.ev_pcounter[i]; break;
                        case 3: v = (double)D.is_cons[i]; break;
                        case 4: v = (double)D.ev_partner[i]; break;
                        case 5: v = (double)D.ev_sub[i]; break;
                        case 6: v = D.vx[i]; break;
                        case 7: v = D.vy[i]; break;
                        case 8: v = D.vz[i]; break;
                        case 9: v = D.ev_pot[i]; break;
                        case 10: v = (double)D.cell_idx[3 * i]; break;
                        case 11: v = (double)D.type[i]; break;
                        case 12: v = D.gauss[s_ctl.g_pos % D.g_cap]; break;
                        case 13: v = D.gauss[(s_ctl.g_pos + 4) % D.g_cap]; break;
                        case 14: v = (double)D.cbt_leaf[i]; break;
                        default: break;
                    }
                }
                s_sink[32 + l] = v;
            }
            if (threadIdx.x >= 224) {
                // ... and walk the tournament-tree paths of the event's atoms now (leaf -> siblings -> their times)
                int lane = threadIdx.x - 224;
                double acc = 0;
                for (int q = 0; q < 2; q++) {
                    int i = q ? job.partner : job.target;
                    if (i <= 0) continue;
                    int child = D.cbt_leaf[i] >> lane;
                    if (child > 1) {
                        int sib = D.cbt_node[child ^ 1];
                        acc += (double)D.cbt_node[child >> 1] + D.lz_t[sib] + (double)D.ev_partner[sib];
                    }
                }
                s_sink[lane] = acc;
            }
            lz1::d_loop_body_cta(D, sh, t_until, trace, trace_cap);
            if (threadIdx.x == 0) {
                long long k2 = clock64();
                lz1::d_loop_head(D, t_until, iter_limit);
                d_publish_next_job(D, r ^ 1);
                // for the sweep blocks: payload first, the iteration number last (release) — it also publishes the event block's
                // writes of this iteration's two atoms
                const Ctl::Job &nj = s_ctl.job[r ^ 1];
                Ctl::Job *slot = &Dg.ctl->job[r ^ 1];
                slot->dt = nj.dt; slot->target = nj.target; slot->partner = nj.partner; slot->go = nj.go; slot->pad_ = 0;
                st_release_i64(reinterpret_cast<long long *>(&slot->seq), nj.seq);
                s_ctl.phase_cycles[0] += k1 - k0;
                s_ctl.phase_cycles[1] += clock64() - k2;
            }
            __syncthreads();
        }
        seq++;
    }
    if (ev) {
        __syncthreads();
        for (int k = threadIdx.x; k < (int)(sizeof(Ctl) / sizeof(int)); k += blockDim.x) ((int *)Dg.ctl)[k] = ((const int *)&s_ctl)[k];
    }
}

// ---- single-step kernels behind the per-function C-ABI ----
__global__ void k_predict_list(DevP D, int a0, int a1) { d_predict_list_serial(D, a0, a1); }
// Predict(renewList) for lists of any length (the reference builds them for CG residues, SGThread.c:106-117):
// ResetTarget over the whole list first, then the atoms in list order, one thread
__global__ void k_predict_list_n(DevP D, const int *list, int cnt) {
    SlotSrc S = d_live_slots(D);
    for (int k = 0; k < cnt; k++) {
        int i = list[k];
        D.ev_type[i] = SDMD_EV_INVD; D.ev_time[i] = SD_INF; D.ev_partner[i] = SDMD_INVALID;
    }
    for (int k = 0; k < cnt; k++) {
        int i = list[k];
        Slot cur = d_reset_slot(D.ev_pot[i], D.ev_sub[i], D.ev_pcounter[i]);
        double g = 0;
        if (D.thermostat_on && !D.is_cons[i]) g = d_next_gauss(D);
        int np = 0;
        cur = d_predict_atom<false>(D, S, i, cur, g, 0x7fffffff, &np);
        d_store_slot(D, i, cur);
        D.ctl->pair_predictions += np;
    }
}
// Predict(all) in the reference's own order by one thread: for systems with a moving wall, where a prediction (WallTime's retry)
// changes the box every later prediction sees — small confined systems, SGThread.c:47-52,158-211
__global__ void k_predict_all_serial(DevP D) {
    SlotSrc S = d_live_slots(D);
    for (int i = 1; i <= D.n; i++) { D.ev_type[i] = SDMD_EV_INVD; D.ev_time[i] = SD_INF; D.ev_partner[i] = SDMD_INVALID; }
    for (int i = 1; i <= D.n; i++) {
        Slot cur = d_reset_slot(D.ev_pot[i], D.ev_sub[i], D.ev_pcounter[i]);
        double g = 0;
        if (D.thermostat_on && !D.is_cons[i]) g = d_next_gauss(D);
        int np = 0;
        cur = d_predict_atom<false>(D, S, i, cur, g, 0x7fffffff, &np);
        d_store_slot(D, i, cur);
        D.ctl->pair_predictions += np;
    }
    D.ctl->pa_passes++;
}
// DoWallDyn, Events.c:528-532: a wall event predicted against the old wall position is stale
__global__ void k_wall_counter_bump(DevP D) {
    int i = blockIdx.x * blockDim.x + threadIdx.x + 1;
    if (i <= D.n && D.ev_type[i] == SDMD_EV_WALL) D.ev_counter[i]++;
}
__global__ void k_cbt_update(DevP D, int a0, int a1) {
    if (a0 > 0) d_cbt_update_run(D, a0);
    if (a1 > 0) d_cbt_update_run(D, a1);
}
__global__ void k_assign_job(DevP D) {
    Ctl *c = D.ctl;
    int target = D.cbt_node[1];
    c->target = target;
    c->dt = D.ev_time[target];
    int p = D.ev_partner[target];
    c->partner = p > 0 ? p : 0;
}
__global__ void k_hazard(DevP D, int target, int partner, int *out) {
    *out = (partner > 0 && D.ev_pcounter[target] < D.ev_counter[partner]) ? -1 : 0;
}
__global__ void k_do_event(DevP D) {
    Ctl *c = D.ctl;
    c->frame++;
    d_do_event(D, c->target, c->partner);
    D.ev_counter[c->target]++;
    if (c->partner && !D.is_cons[c->partner]) D.ev_counter[c->partner]++;
}
__global__ void k_cell_move(DevP D, int atom) { d_pcc_event(D, atom); }

// CalPotenE, ToolFunctions.c:659-752: the accumulation order is part of the result, so one thread adds the
// pair terms in the reference's order
// The addends atom i contributes to CalPotenE (ToolFunctions.c:678-749), in the reference's order.  The energy is a
// running sum over all of them in atom order, so the ORDER is part of the result: WRITE = false counts the non-zero
// addends, WRITE = true stores them (adding +-0.0 never changes a sum that started at +0.0, so zeros are dropped).
template <bool WRITE>
__device__ inline int d_potential_terms(const DevP &D, int i, double *out, int cap = 0x7fffffff) {
    int cnt = 0;
    double accum = 0;
    int cx = D.cell_idx[3 * i], cy = D.cell_idx[3 * i + 1], cz = D.cell_idx[3 * i + 2];
    double xi = D.x[i], yi = D.y[i], zi = D.z[i];
    for (int n = 0; n < 27; n++) {
        int nx = cx + c_offx[n], ny = cy + c_offy[n], nz = cz + c_offz[n];
        double shx = 0, shy = 0, shz = 0;
        if (nx < 0) { nx = D.cellnum[0] - 1; shx = -1 * d_box(D, 0); } else if (nx >= D.cellnum[0]) { nx = 0; shx = d_box(D, 0); }
        if (ny < 0) { ny = D.cellnum[1] - 1; shy = -1 * d_box(D, 1); } else if (ny >= D.cellnum[1]) { ny = 0; shy = d_box(D, 1); }
        if (nz < 0) { nz = D.cellnum[2] - 1; shz = -1 * d_box(D, 2); } else if (nz >= D.cellnum[2]) { nz = 0; shz = d_box(D, 2); }
        int cell = nz * D.cellnum[0] * D.cellnum[1] + ny * D.cellnum[0] + nx;
        for (int pair = D.celllist[D.n + cell + 1]; pair != 0; pair = D.celllist[pair]) {
            if (pair <= i) continue;
            double tx = D.x[pair] + shx, ty = D.y[pair] + shy, tz = D.z[pair] + shz;
            double rx = xi - tx, ry = yi - ty, rz = zi - tz;
            double d2 = rx * rx + ry * ry + rz * rz;
            int cm = d_cm_flags(D, i, pair);
            if (d2 <= D.cutfull2 && !(cm & SDMD_BOND_CONNECT)) {
                int ai = i, aj = pair, hbp = 0, mode;
                if (D.hb_bond[ai] == aj) {
                    mode = FP_HB;
                    double bar = d_hb_barrier(D, ai, aj);
                    if (bar != 0) { if (WRITE && cnt < cap) out[cnt] = -bar; cnt++; }     // TPE -= HBBarrier(...)
                } else if (cm & SDMD_CONSTRAINT_CONNECT) mode = FP_CONSTRAINT;
                else if (d_neighbor_connected(D, ai, aj)) {
                    if (D.hb_neighbor[D.hb_bond[ai]] == aj) { int s = ai; ai = aj; aj = s; }
                    hbp = D.hb_bond[aj];
                    mode = FP_NEIGHBOR;
                } else mode = FP_COLLISION;
                double wx = D.vx[ai] - D.vx[aj], wy = D.vy[ai] - D.vy[aj], wz = D.vz[ai] - D.vz[aj];
                double direction = rx * wx + ry * wy + rz * wz;
                accum = d_find_pair(D, ai, aj, mode, direction, d2, hbp, 0).accum;
                if (accum < SD_INF / 2 && accum != 0) { if (WRITE && cnt < cap) out[cnt] = accum; cnt++; }
            }
        }
    }
    return cnt;
}

// small systems (natom <= 2048): the threads of ONE block walk the atoms' neighbourhoods side by side, each atom's addends go to
// its row of DevP::pe_terms in CalPotenE's visiting order, then thread 0 adds them up atom by atom, addend by addend — the
// reference's summation order, hence its bits.  ONE walk per atom: FindPair's warnings count once, like CalPotenE's.
// Returns the energy in thread 0.
#define PE_ROW 64
__device__ inline double d_potential_energy(const DevP &D) {
    for (int i = threadIdx.x + 1; i <= D.n; i += blockDim.x)
        D.pe_cnt[i] = d_potential_terms<true>(D, i, D.pe_terms + (size_t)i * PE_ROW, PE_ROW);
    __syncthreads();
    double TPE = 0;
    if (threadIdx.x == 0) {
        for (int i = 1; i <= D.n; i++) {
            int cnt = D.pe_cnt[i];
            if (cnt > PE_ROW) { D.ctl->error = SDMD_ERR_STATE; continue; }   // more addends than any physical neighbourhood produces
            const double *row = D.pe_terms + (size_t)i * PE_ROW;
            for (int k = 0; k < cnt; k++) TPE += row[k];
        }
    }
    return TPE;
}
// large systems: addends produced by one thread per atom (count, scan, write), summed in order by one warp
__global__ void k_pe_count(DevP D, int *count) {
    int i = blockIdx.x * blockDim.x + threadIdx.x + 1;
    if (i <= D.n) count[i - 1] = d_potential_terms<false>(D, i, nullptr);
}
__global__ void k_pe_write(DevP D, const int *offset, double *terms) {
    int i = blockIdx.x * blockDim.x + threadIdx.x + 1;
    if (i <= D.n) d_potential_terms<true>(D, i, terms + offset[i - 1]);
}
__global__ void k_ordered_sum(const double *terms, const int *total_ptr, long long total_direct, double *out) {
    long long total = total_ptr ? (long long)*total_ptr : total_direct;
    int lane = threadIdx.x;
    double s = 0;
    for (long long base = 0; base < total; base += 32) {
        double v = (base + lane < total) ? terms[base + lane] : 0.0;
        int m = (total - base < 32) ? (int)(total - base) : 32;
        for (int l = 0; l < m; l++) s += __shfl_sync(0xffffffffu, v, l);   // every lane carries the same ordered sum
    }
    if (lane == 0) *out = s;
}
__global__ void k_potential_energy(DevP D, double *out) {
    double e = d_potential_energy(D);
    if (threadIdx.x == 0) *out = e;
}
__global__ void k_potential_energy_multi(const MultiArg *args, double *out) {
    double e = d_potential_energy(args[blockIdx.x].P);
    if (threadIdx.x == 0) out[blockIdx.x] = e;
}
// CalKinetE, ToolFunctions.c:626-656 (flow.mark < 2, constV = 0); ordered sum
__global__ void k_ke_terms(DevP D, double *terms) {
    int i = blockIdx.x * blockDim.x + threadIdx.x + 1;
    if (i > D.n) return;
    double a = D.vx[i] - D.flow_v[0], b = D.vy[i] - D.flow_v[1], c = D.vz[i] - D.flow_v[2];   // CalKinetE, ToolFunctions.c:631-635
    double s = a * a + b * b + c * c;
    terms[i - 1] = 0.5 * D.mass[i] * s;
}
__global__ void k_kinetic_energy(DevP D, double *out) {
    double TKE = 0;
    if (D.flow_mark < 2) {                     // relative to the drift velocity, ToolFunctions.c:630-636
        for (int i = 1; i <= D.n; i++) {
            double a = D.vx[i] - D.flow_v[0], b = D.vy[i] - D.flow_v[1], c = D.vz[i] - D.flow_v[2];
            double s = a * a + b * b + c * c;
            TKE += 0.5 * D.mass[i] * s;
        }
    } else {                                   // relative to each chain's centre of mass, :637-649 + CalCOMV :613-623
        int i = 1;
        while (i <= D.n) {
            int j = i;
            while (j <= D.n && D.protein[j] == D.protein[i]) j++;
            double pm = D.protein_mass ? D.protein_mass[i] : 0, nx = 0, ny = 0, nz = 0;
            for (int n = i; n < j; n++) { nx += D.vx[n] * D.mass[n]; ny += D.vy[n] * D.mass[n]; nz += D.vz[n] * D.mass[n]; }
            nx /= pm; ny /= pm; nz /= pm;
            for (int n = i; n < j; n++) {
                double a = D.vx[n] - nx, b = D.vy[n] - ny, c = D.vz[n] - nz;
                double s = a * a + b * b + c * c;
                TKE += 0.5 * D.mass[n] * s;
            }
            i = j;
        }
    }
    *out = TKE;
}

// ===================================================================================================
// C-ABI
// ===================================================================================================
extern "C" {

const char *sdmd_cuda_last_error(void) { return g_err; }

int sdmd_cuda_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
    return n;
}

static inline int grid_for(int n, int block) { return (n + block - 1) / block; }
/* Loop drivers by system size (profiles/r1_loop_crossover.txt, GEL boxes; us per loop iteration on a B200):
 *   one block        9.5 @ 512, 12.7 @ 1728, 18 @ 4096: its block-wide advance costs ~1.1 us per 1000 atoms (one SM's L2 bandwidth)
 *   one cluster (16) 13.9 @ 4096, 14.3 @ 8000, 16.6 @ 32768, 19.0 @ 64000, 24.9 @ 125000
 *   sweep            a flat 21-22 (GEL) to ~26 (proteins)
 * SDMD_SMALL_N is the largest system the cluster driver takes in AUTO mode . */
#define SDMD_SMALL_N 98304
#define SDMD_ONE_CTA_N 2048   /* up to here: one block; up to SDMD_SMALL_N: one cluster whose other blocks help with the advance */
#define SDMD_MANY_N 8192      /* run_many: one block per replica */

// leaf permutation of CreateCBT's sequential InsertCBTRun (CBT.c:86-105): time independent
static void cbt_leaf_permutation(int n, std::vector<int> &leaf_atom) {
    std::vector<int> node(2 * (size_t)n + 2, 0);
    int count = 0;
    for (int a = 1; a <= n; a++) {
        if (count == 0) { node[1] = a; count++; continue; }
        int i = node[count];
        node[count * 2] = i;
        node[count * 2 + 1] = a;
        count++;
    }
    leaf_atom.resize(n);
    if (n == 1) { leaf_atom[0] = 1; return; }
    for (int k = 0; k < n; k++) leaf_atom[k] = node[n + k];
}

static int create_impl(sdmd_ctx *c, const struct sdmd_host_view *v);
int sdmd_cuda_create(const struct sdmd_host_view *v, int device, sdmd_ctx **out) {
    if (!v || !out || v->abi_version != SDMD_ABI_VERSION || v->natom < 1) return set_err(SDMD_ERR_BAD_ARG, "bad host view");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return set_err(SDMD_ERR_NO_DEVICE, "no CUDA device (there is no CPU fallback)");
    if (device < 0 || device >= ndev) return set_err(SDMD_ERR_BAD_ARG, "bad device index");
    CK(cudaSetDevice(device));
    sdmd_ctx *c = new sdmd_ctx();
    c->device = device;
    c->n = v->natom;
    int rc = create_impl(c, v);
    if (rc) { sdmd_cuda_destroy(c); return rc; }   // releases whatever had been set up before the failure
    *out = c;
    return 0;
}
static int create_impl(sdmd_ctx *c, const struct sdmd_host_view *v) {
    CK(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    CK(cudaEventCreate(&c->ev0));
    CK(cudaEventCreate(&c->ev1));
    CK(cudaEventCreate(&c->evt0));
    CK(cudaEventCreate(&c->evt1));
    int n = v->natom;
    size_t n1 = (size_t)n + 1;
    DevP &P = c->P;
    memset(&P, 0, sizeof(P));
    SingleBody SB;
    memset(&SB, 0, sizeof(SB));
    P.n = n;
    // CreateCell, Initialization.c:1060-1073
    for (int k = 0; k < 3; k++) {
        P.box[k] = v->box[k];
        P.cellnum[k] = (int)floor(v->box[k] / v->cutoffr);
        if (P.cellnum[k] < 3) P.cellnum[k] = 3;
        if (P.cellnum[k] > 65535) return set_err(SDMD_ERR_UNSUPPORTED, "more than 65535 cells along one axis");
    }
    for (int k = 0; k < 3; k++) P.cellsize[k] = v->box[k] / P.cellnum[k];
    long long ncell = (long long)P.cellnum[0] * P.cellnum[1] * P.cellnum[2];
    if (ncell > (1LL << 30)) return set_err(SDMD_ERR_BAD_ARG, "cell grid too large");   // the caller releases c
    P.ncell = (int)ncell;
    P.cutoffr = v->cutoffr;
    P.cut2 = v->cutoffr * SDMD_EXTEND_RATIO;  // TimePrediction.c:53-54
    P.cut2 *= P.cut2;
    P.cutfull2 = v->cutoffr * v->cutoffr;      // Events.c:1031, ToolFunctions.c:709
    P.thermoF = v->thermoF;
    if (v->wall_kind < SDMD_WALL_NONE || v->wall_kind > SDMD_WALL_SPHERE) return set_err(SDMD_ERR_BAD_ARG, "bad wall_kind");
    if (v->wall_kind && (v->wall_type < 1 || v->wall_type > SDMD_NATOMTYPE)) return set_err(SDMD_ERR_BAD_ARG, "bad wall_type");
    P.wall_kind = v->wall_kind;
    P.wall_type = v->wall_type;
    for (int k = 0; k < 3; k++) P.wall_pos[k] = v->wall_pos[k];
    P.pbc_mask = v->wall_kind == SDMD_WALL_PARALLEL ? 5 : v->wall_kind == SDMD_WALL_CYLINDER ? 1 : v->wall_kind == SDMD_WALL_SPHERE ? 0 : 7;
    if (v->n_obst < 0 || v->n_obst > SDMD_MAX_OBST || v->n_sphobst < 0 || v->n_sphobst > SDMD_MAX_OBST)
        return set_err(SDMD_ERR_UNSUPPORTED, "more obstruction walls / spherical obstacles than SDMD_MAX_OBST");
    P.n_obst = v->n_obst;
    P.n_sphobst = v->n_sphobst;
    for (int n = 0; n < v->n_obst; n++) {
        if (v->obst_type[n] < 1 || v->obst_type[n] > SDMD_NATOMTYPE) return set_err(SDMD_ERR_BAD_ARG, "bad obst_type");
        SB.obst_type[n] = v->obst_type[n];
        for (int k = 0; k < 3; k++) SB.obst_pos[n][k] = v->obst_pos[n][k];
    }
    for (int n = 0; n < v->n_sphobst; n++) {
        SB.sph_r[n] = v->sphobst_radius[n];
        for (int k = 0; k < 3; k++) SB.sph_pos[n][k] = v->sphobst_pos[n][k];
    }
    if (v->n_tunnel < 0 || v->n_tunnel > SDMD_MAX_OBST) return set_err(SDMD_ERR_UNSUPPORTED, "more tunnels than SDMD_MAX_OBST");
    if (v->n_tunnel && (v->tunnel_type < 1 || v->tunnel_type > SDMD_NATOMTYPE)) return set_err(SDMD_ERR_BAD_ARG, "bad tunnel_type");
    P.n_tunnel = v->n_tunnel;
    P.tunnel_type = v->tunnel_type;
    for (int n = 0; n < v->n_tunnel; n++) {
        SB.tun_diam[n] = v->tunnel_diameter[n];
        for (int k = 0; k < 4; k++) SB.tun_pos[n][k] = v->tunnel_pos[n][k];
    }
    for (int n = 0; n < v->n_obst; n++) SB.obst_hole[n] = v->obst_hole[n];
    P.wall_dyn = (v->wall_dyn != 0 && v->wall_kind != 0) ? 1 : 0;
    P.wd_active = P.wall_dyn;
    P.wd_step = v->wall_dyn_step;
    P.wd_orig_y = v->wall_dyn_orig_box[1];
    P.single_body = P.wall_kind != 0 || P.n_obst > 0 || P.n_sphobst > 0 || P.n_tunnel > 0 || v->flow_mark == 3;
    P.thermostat_on = v->thermostat_on;
    P.method_ding = v->method_ding;
    P.hydrogen_bond = v->hydrogen_bond;
    if (P.hydrogen_bond) {
        // HBTime / HBNeighborTime (TimePrediction.c:292-578) do nothing for an atom with no donor/acceptor role, no bond,
        // no neighbour entries and interactionType already NoEvent; a system made only of such atoms (GEL boxes) can
        // never create any, so the HB stages are switched off for it instead of re-reading four arrays per prediction
        bool any = false;
        for (int i = 1; i <= v->natom && !any; i++) {
            int r = v->hb_role ? v->hb_role[i] : 0;
            any = r == 'A' || r == 'D' || r == 'S' || (v->hb_bond && v->hb_bond[i]) || (v->hbn_status && v->hbn_status[i]) ||
                  (v->hb_itype && v->hb_itype[i] != SDMD_HB_NOEVENT);
        }
        if (!any) P.hydrogen_bond = 0;
    }
    P.gel_mode = v->gel_mode;
    for (int k = 0; k < 3; k++) P.hb_barrier[k] = v->hb_barrier[k];

    TRY(dupload(c, &P.x, v->x, n1)); TRY(dupload(c, &P.y, v->y, n1)); TRY(dupload(c, &P.z, v->z, n1));
    TRY(dupload(c, &P.vx, v->vx, n1)); TRY(dupload(c, &P.vy, v->vy, n1)); TRY(dupload(c, &P.vz, v->vz, n1));
    TRY(dupload(c, (double **)&P.mass, v->mass, n1));
    if (v->flow_mark < 0 || v->flow_mark > 3) return set_err(SDMD_ERR_BAD_ARG, "flow type");
    if (v->n_charge < 0 || v->n_charge > SDMD_MAX_CHARGE) return set_err(SDMD_ERR_UNSUPPORTED, "more charge sources than SDMD_MAX_CHARGE");
    if (v->flow_mark == 3) {
        if (!v->charge) return set_err(SDMD_ERR_BAD_ARG, "flow_mark 3 needs per-atom charges");
        TRY(dupload(c, (double **)&P.charge, v->charge, n1));
        P.n_charge = v->n_charge;
        P.charge_com = v->charge_com;
        for (int q = 0; q < v->n_charge; q++) {
            for (int k = 0; k < 3; k++) { SB.chg_pos[q][k] = v->charge_pos[q][k]; SB.chg_vel[q][k] = v->charge_vel[q][k]; }
            SB.chg_grad[q] = v->charge_potgrad[q];
            SB.chg_gap[q] = v->charge_gap[q];
        }
    }
    P.flow_mark = v->flow_mark;
    for (int k = 0; k < 3; k++) P.flow_v[k] = v->flow_v[k];
    if (v->flow_mark) {
        if (v->flow_a) TRY(dupload(c, (double **)&P.flow_a, v->flow_a, 3 * n1));
        if (v->flow_time_rec) TRY(dupload(c, &P.flow_time_rec, v->flow_time_rec, n1));
        else TRY(dalloc(c, &P.flow_time_rec, n1));
        if (v->protein_mass) TRY(dupload(c, (double **)&P.protein_mass, v->protein_mass, n1));
        else if (v->flow_mark >= 2) return set_err(SDMD_ERR_BAD_ARG, "flow_mark >= 2 needs protein_mass");
    }
    if (P.single_body) TRY(dupload(c, (SingleBody **)&P.sb, &SB, 1));
    TRY(dupload(c, &P.type, v->type, n1));
    TRY(dupload(c, (int **)&P.aa_type, v->aa_type, n1));
    TRY(dupload(c, (int **)&P.protein, v->protein_num, n1));
    TRY(dupload(c, (int **)&P.aa_num, v->aa_num, n1));
    TRY(dupload(c, (uint8_t **)&P.is_cons, v->is_cons, n1));
    TRY(dupload(c, &P.hb_bond, v->hb_bond, n1));
    TRY(dupload(c, (int **)&P.hb_neighbor, v->hb_neighbor, n1));
    TRY(dupload(c, (uint8_t **)&P.hb_role, v->hb_role, n1));
    P.hb_name = nullptr;
    if (v->hb_name) TRY(dupload(c, (uint8_t **)&P.hb_name, v->hb_name, n1));
    TRY(dupload(c, &P.hb_itype, v->hb_itype, n1));
    TRY(dupload(c, &P.hbn_partner, v->hbn_partner, 5 * n1));
    TRY(dupload(c, &P.hbn_pnum, v->hbn_partner_num, n1));
    TRY(dupload(c, &P.hbn_status, v->hbn_status, n1));
    TRY(dupload(c, &P.ev_time, v->ev_time, n1));
    TRY(dupload(c, &P.ev_pot, v->ev_potential, n1));
    TRY(dupload(c, &P.ev_partner, v->ev_partner, n1));
    TRY(dupload(c, &P.ev_counter, (const i64 *)v->ev_counter, n1));
    TRY(dupload(c, &P.ev_pcounter, (const i64 *)v->ev_pcounter, n1));
    TRY(dupload(c, &P.ev_type, v->ev_type, n1));
    TRY(dupload(c, &P.ev_sub, v->ev_sub, n1));
    TRY(dupload(c, &P.cell_idx, v->cell_idx, 3 * n1));
    TRY(dalloc(c, &P.celllist, n1 + P.ncell));
    TRY(dalloc(c, &P.cbt_node, 2 * n1));
    TRY(dalloc(c, &P.cbt_leaf, n1));
    int nb = v->bond_off ? v->bond_off[n + 1] : 0, nc = v->con_off ? v->con_off[n + 1] : 0, ncm = v->cm_off ? v->cm_off[n + 1] : 0;
    std::vector<int> zero_off(n1 + 1, 0);
    TRY(dupload(c, (int **)&P.bond_off, v->bond_off ? v->bond_off : zero_off.data(), n1 + 1));
    TRY(dupload(c, (int **)&P.bond_j, v->bond_j, (size_t)nb));
    TRY(dupload(c, (double **)&P.bond_dmin2, v->bond_dmin2, (size_t)nb));
    TRY(dupload(c, (double **)&P.bond_dmax2, v->bond_dmax2, (size_t)nb));
    TRY(dupload(c, (int **)&P.con_off, v->con_off ? v->con_off : zero_off.data(), n1 + 1));
    TRY(dupload(c, (int **)&P.con_j, v->con_j, (size_t)nc));
    TRY(dupload(c, (int **)&P.con_table, v->con_table, (size_t)nc));
    TRY(dupload(c, (int **)&P.cm_off, v->cm_off ? v->cm_off : zero_off.data(), n1 + 1));
    TRY(dupload(c, (int **)&P.cm_j, v->cm_j, (size_t)ncm));
    TRY(dupload(c, (int **)&P.cm_flags, v->cm_flags, (size_t)ncm));
    {   // constraint table of every CONSTRAINT_CONNECT entry, resolved once from the constr list (ToolFunctions.c:70-77)
        std::vector<int> cmt((size_t)ncm + 1, 0);
        for (int i = 1; i <= n && ncm; i++)
            for (int e = v->cm_off[i]; e < v->cm_off[i + 1]; e++)
                if (v->cm_flags[e] & SDMD_CONSTRAINT_CONNECT)
                    for (int k = v->con_off[i]; k < v->con_off[i + 1]; k++)
                        if (v->con_j[k] == v->cm_j[e]) { cmt[e] = v->con_table[k]; break; }
        TRY(dupload(c, (int **)&P.cm_table, cmt.data(), (size_t)ncm));
    }
    TRY(dalloc(c, &P.cell_cnt, (size_t)P.ncell + 1));
    // per-atom candidate-list cache for the event loop (small and medium systems)
    {
        // row capacity from the mean neighbourhood size (27 cells) with head-room; one row per atom while that fits 8 GB
        double need = 27.0 * (double)n / (double)P.ncell * 1.6 + 24.0;
        P.cc_cap = need <= 64 ? 64 : (need <= 128 ? 128 : 256);
        P.cc_per_atom = ((double)n1 * P.cc_cap * 6.0 <= 8.0e9);
        if (!P.cc_per_atom) P.cc_cap = 256;
    }
    {
        size_t rows = P.cc_per_atom ? n1 : 2;
        TRY(dalloc(c, &P.cc_stamp, n1)); TRY(dalloc(c, &P.cc_n, n1));
        TRY(dalloc(c, &P.cell_epoch, (size_t)P.ncell));
        TRY(dalloc(c, &P.cc_j, (size_t)P.cc_cap * rows)); TRY(dalloc(c, &P.cc_k, (size_t)P.cc_cap * rows));
        TRY(dalloc(c, &P.cc_pass, (size_t)P.cc_cap * rows));
        TRY(dalloc(c, &P.cc_passidx, 32 * rows)); TRY(dalloc(c, &P.cc_npass, rows));
    }
    P.cm_dense = nullptr;
    if (n <= 2048) {   // dense static connection matrix: one load instead of a row search in the candidate evaluation
        std::vector<unsigned> dense(n1 * n1, 0u);
        for (int i = 1; i <= n && ncm; i++)
            for (int e = v->cm_off[i]; e < v->cm_off[i + 1]; e++) {
                unsigned tbl = 0;
                if (v->cm_flags[e] & SDMD_CONSTRAINT_CONNECT)
                    for (int k = v->con_off[i]; k < v->con_off[i + 1]; k++)
                        if (v->con_j[k] == v->cm_j[e]) { tbl = (unsigned)v->con_table[k]; break; }
                dense[(size_t)i * n1 + v->cm_j[e]] = (unsigned)(v->cm_flags[e] & 0xff) | (tbl << 8);
            }
        TRY(dupload(c, (unsigned **)&P.cm_dense, dense.data(), n1 * n1));
    }
    TRY(dupload(c, (sdmd_step_table **)&P.tables, v->tables, (size_t)v->ntable));
    TRY(dupload(c, (int **)&P.coll_index, v->coll_index, (size_t)SD_NT * SD_NT));
    TRY(dupload(c, (int **)&P.hb_index, v->hb_index, (size_t)SDMD_NHBCLASS * SD_NT * SD_NT));

    // gaussian ring: room for a full-system pass plus a generous event-loop window
    P.g_cap = (i64)n + (1 << 20);
    TRY(dalloc(c, &c->d_gauss, (size_t)P.g_cap));
    P.gauss = c->d_gauss;
    std::vector<int> rank(n1, 0);
    int r = 0;
    for (int i = 1; i <= n; i++) { rank[i] = r; if (!v->is_cons[i]) r++; }
    P.no_cons = (r == n);   // nobody is constrained: draw index of atom i is simply i - 1
    c->n_draws = r;         // gaussians a full-system pass consumes: constrained atoms skip ThermostatTime (SGThread.c:165-168)
    TRY(dupload(c, (int **)&P.ther_rank, rank.data(), n1));

    // scratch
    TRY(dalloc(c, &c->b_time, n1)); TRY(dalloc(c, &c->b_pot, n1)); TRY(dalloc(c, &c->b_partner, n1));
    TRY(dalloc(c, &c->b_pcounter, n1)); TRY(dalloc(c, &c->b_type, n1)); TRY(dalloc(c, &c->b_sub, n1));
    TRY(dalloc(c, &c->c_time, n1)); TRY(dalloc(c, &c->c_pot, n1)); TRY(dalloc(c, &c->c_partner, n1));
    TRY(dalloc(c, &c->c_pcounter, n1)); TRY(dalloc(c, &c->c_type, n1)); TRY(dalloc(c, &c->c_sub, n1));
    TRY(dalloc(c, &c->npair, n1)); TRY(dalloc(c, &c->pcount, n1)); TRY(dalloc(c, &c->pwho, n1));
    TRY(dalloc(c, &c->queue_a, 2 * n1 + 2)); TRY(dalloc(c, &c->queue_b, 2 * n1 + 2)); TRY(dalloc(c, &c->nqueue, 32)); TRY(dalloc(c, &c->second, n1)); TRY(dalloc(c, &c->bp_time, n1)); TRY(dalloc(c, &c->bp_j, n1));
    c->d_sum = reinterpret_cast<unsigned long long *>(c->nqueue + FIX_NQ);   // zeroed together with the queue lengths
    TRY(dalloc(c, &c->cell_of, n1)); TRY(dalloc(c, &c->cell_count, (size_t)P.ncell + 1));
    TRY(dalloc(c, &c->cell_rank, n1)); TRY(dalloc(c, &c->cell_tmp, n1));
    TRY(dalloc(c, &c->scan_tmp, (size_t)P.ncell / SCAN_ITEMS + 4));
    TRY(dalloc(c, &c->scan_state, (size_t)P.ncell / SCAN_ITEMS + 4)); TRY(dalloc(c, &c->scan_ticket, 4));
    TRY(dalloc(c, &c->d_cell_start, (size_t)P.ncell + 1)); TRY(dalloc(c, &c->d_sorted_atom, n1));
    TRY(dalloc(c, &c->d_srec, n1)); TRY(dalloc(c, &c->d_sloc, n1)); TRY(dalloc(c, &c->d_scell, n1));
    {
        int per_sm = 0, sms = 0;
        CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_predict_all_fixpoint, 128, 0));
        CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, c->device));
        c->fix_grid = per_sm * sms;        // the cooperative grid of the fix-up rounds: every block resident
        if (c->fix_grid <= 0) return set_err(SDMD_ERR_CUDA, "cooperative grid size");
    }
    P.surv = nullptr; P.surv_n = nullptr; P.surv_np = nullptr;
    TRY(dalloc(c, &P.spos, n1));
    P.cut2_loose = P.cut2 * (1.0 + 1e-9);
    P.cut2_f32 = nextafterf((float)(P.cut2 * (1.0 + 1e-4)), INFINITY);
    for (int k = 0; k < 3; k++) P.cellsize_f[k] = (float)P.cellsize[k];
    P.sloc = c->d_sloc; P.scell = c->d_scell;
    {
        int trc = sdmd_k2_tile_configure(P, c->device);   // bond-free systems: round 0 as one tile kernel, no survivor rows in HBM
        if (trc < 0) return set_err(SDMD_ERR_CUDA, "tile kernel configuration", cudaGetLastError());
    }
    if (n < (1 << 27) && P.tile_mode != 1) {
        // row size: a sparse bond-free box (the GEL lattices: ~4 atoms expected inside the scan sphere) gets 64-byte rows, which
        // halves the footprint the evaluation kernel streams through L2 (287 -> 275 us at 1M atoms); longer lists overflow
        // into the exact in-kernel re-scan either way
        {
            double vol = P.box[0] * P.box[1] * P.box[2];
            double expected = vol > 0 ? (double)n / vol * (4.0 / 3.0) * 3.14159265358979 * P.cut2 * sqrt(P.cut2) : 1e9;
            P.surv_cap = (P.gel_mode && !P.hydrogen_bond && expected <= 6.0) ? 16 : R0_SURV_CAP;
        }
        TRY(dalloc(c, &P.surv, (size_t)R0_SURV_CAP * n1)); TRY(dalloc(c, &P.surv_n, n1)); TRY(dalloc(c, &P.surv_np, n1));
    }
    P.cell_start = c->d_cell_start; P.sorted_atom = c->d_sorted_atom;
    P.srec = c->d_srec;
    P.cm_empty = (ncm == 0);
    P.uni_table = -1;
    if (!P.hydrogen_bond) {     // (atom types only change through hydrogen bonds)
        bool same = true;
        for (int i = 2; i <= n && same; i++) same = v->type[i] == v->type[1];
        int idx = same ? v->coll_index[v->type[1] * SD_NT + v->type[1]] : -1;
        if (idx >= 0 && v->tables[idx].nstep <= 2 && v->tables[idx].dmax2 == 0 && v->tables[idx].dmin2 != 0) P.uni_table = idx;
    }
    P.rows_mode = 0;
    if (P.surv) sdmd_k2_rows_configure(P);   // bond-free systems: row-batched filter + cp.async evaluation (sdmd_k2_rows.cu)
    std::vector<int> leaf_atom;
    cbt_leaf_permutation(n, leaf_atom);
    TRY(dupload(c, &c->d_leaf_atom, leaf_atom.data(), (size_t)n));

    TRY(dalloc(c, &c->d_ctl, 1));
    TRY(dalloc(c, &c->d_ctl_scratch, 1));
    P.pe_terms = nullptr; P.pe_cnt = nullptr;
    if (n <= 2048) { TRY(dalloc(c, &P.pe_terms, n1 * 64)); TRY(dalloc(c, &P.pe_cnt, n1)); }
    CK(cudaMallocHost((void **)&c->h_ctl, sizeof(Ctl)));
    memset(c->h_ctl, 0, sizeof(Ctl));
    c->h_ctl->currenttime = v->currenttime;
    c->h_ctl->timestep = v->timestep;
    c->h_ctl->temperature = v->target_temperature;
    c->h_ctl->frame = v->frame;
    c->h_ctl->hbnumformed = v->hb_num_formed;
    for (int k = 0; k < 3; k++) c->h_ctl->dbox[k] = v->box[k];
    c->h_ctl->wd_touch = v->wall_dyn_touch;
    c->h_ctl->alphaformed = v->alpha_hb_formed;
    c->h_ctl->cc_epoch = 1;
    c->h_ctl->cc_floor = 1;
    CK(cudaMemcpyAsync(c->d_ctl, c->h_ctl, sizeof(Ctl), cudaMemcpyHostToDevice, c->stream));
    P.ctl = c->d_ctl;
    CK(cudaStreamSynchronize(c->stream));
    return 0;
}

static void remd_forget(sdmd_ctx *c);
void sdmd_cuda_destroy(sdmd_ctx *c) {
    if (!c) return;
    cudaSetDevice(c->device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    for (void *p : c->bufs) cudaFree(p);
    for (int k = 0; k < 4; k++) if (c->scratch[k]) cudaFree(c->scratch[k]);
    if (c->d_trace) cudaFree(c->d_trace);
    remd_forget(c);
    if (c->d_xchg) cudaFree(c->d_xchg);
    if (c->h_xchg) cudaFreeHost(c->h_xchg);
    if (c->h_ctl) cudaFreeHost(c->h_ctl);
    resolve_pending(c);
    for (cudaEvent_t e : c->ev_pool) cudaEventDestroy(e);
    if (c->ev0) cudaEventDestroy(c->ev0);
    if (c->ev1) cudaEventDestroy(c->ev1);
    if (c->evt0) cudaEventDestroy(c->evt0);
    if (c->evt1) cudaEventDestroy(c->evt1);
    if (c->ev_copy) cudaEventDestroy(c->ev_copy);
    if (c->ev_main) cudaEventDestroy(c->ev_main);
    if (c->copy_stream) cudaStreamDestroy(c->copy_stream);
    if (c->stream) cudaStreamDestroy(c->stream);
    delete c;
}

int sdmd_cuda_cell_grid(const sdmd_ctx *c, int cellnum[3], double cellsize[3]) {
    if (!c) return set_err(SDMD_ERR_BAD_ARG, "null ctx");
    for (int k = 0; k < 3; k++) { cellnum[k] = c->P.cellnum[k]; cellsize[k] = c->P.cellsize[k]; }
    return 0;
}

void *sdmd_cuda_stream(const sdmd_ctx *c) { return c ? (void *)c->stream : nullptr; }
int64_t sdmd_cuda_launch_count(const sdmd_ctx *c) { return c ? c->launches : 0; }
double sdmd_cuda_kernel_ms(const sdmd_ctx *c, const char *name) {
    resolve_pending(const_cast<sdmd_ctx *>(c));
    auto it = c->kstat.find(name);
    if (it == c->kstat.end() || it->second.n == 0) return -1.0;
    return it->second.ms / it->second.n;
}
int sdmd_cuda_kernel_stats(const sdmd_ctx *c, const char *name, double *total_ms, int64_t *count) {
    resolve_pending(const_cast<sdmd_ctx *>(c));
    if (!strcmp(name, "predict_all_rounds") || !strcmp(name, "predict_all_requeued")) {
        // booked on the device by the fix-up kernel (no host look inside the pass): totals since the last reset
        sdmd_ctx *m = const_cast<sdmd_ctx *>(c);
        Ctl h;
        if (cudaMemcpyAsync(&h, m->d_ctl, sizeof(Ctl), cudaMemcpyDeviceToHost, m->stream) != cudaSuccess || cudaStreamSynchronize(m->stream) != cudaSuccess)
            return set_err(SDMD_ERR_CUDA, "kernel_stats", cudaGetLastError());
        *total_ms = (double)((name[12] == 'r' && name[13] == 'o') ? h.pa_rounds - m->pa_base[1] : h.pa_requeued - m->pa_base[2]);
        *count = h.pa_passes - m->pa_base[0];
        return 0;
    }
    auto it = c->kstat.find(name);
    if (it == c->kstat.end()) { *total_ms = 0; *count = 0; return 0; }
    *total_ms = it->second.ms;
    *count = it->second.n;
    return 0;
}
void sdmd_cuda_reset_kernel_stats(sdmd_ctx *c) {
    resolve_pending(c);
    c->kstat.clear();
    c->launches = 0;
    Ctl h;
    if (cudaMemcpyAsync(&h, c->d_ctl, sizeof(Ctl), cudaMemcpyDeviceToHost, c->stream) == cudaSuccess && cudaStreamSynchronize(c->stream) == cudaSuccess) {
        c->pa_base[0] = h.pa_passes; c->pa_base[1] = h.pa_rounds; c->pa_base[2] = h.pa_requeued;
    }
}
// device-time bracket on the context's stream for callers (bench.py): begin records, end synchronises
int sdmd_cuda_timer_begin(sdmd_ctx *c) {
    CK(cudaSetDevice(c->device));
    CK(cudaStreamSynchronize(c->stream));
    CK(cudaEventRecord(c->evt0, c->stream));
    return 0;
}
int sdmd_cuda_timer_end(sdmd_ctx *c, double *ms) {
    CK(cudaEventRecord(c->evt1, c->stream));
    CK(cudaEventSynchronize(c->evt1));
    float f = 0;
    CK(cudaEventElapsedTime(&f, c->evt0, c->evt1));
    *ms = f;
    return 0;
}
int sdmd_cuda_synchronize(sdmd_ctx *c) {
    CK(cudaSetDevice(c->device));
    CK(cudaStreamSynchronize(c->stream));
    return 0;
}
// page-lock a caller-owned host buffer so H2D/D2H copies run at full PCIe rate and asynchronously
int sdmd_cuda_host_register(void *ptr, size_t bytes) {
    cudaError_t e = cudaHostRegister(ptr, bytes, cudaHostRegisterDefault);
    if (e != cudaSuccess && e != cudaErrorHostMemoryAlreadyRegistered) return set_err(SDMD_ERR_CUDA, "cudaHostRegister", e);
    cudaGetLastError();
    return 0;
}
int sdmd_cuda_host_unregister(void *ptr) {
    cudaHostUnregister(ptr);
    cudaGetLastError();
    return 0;
}
int sdmd_cuda_set_loop_mode(sdmd_ctx *c, int mode) {
    if (mode != SDMD_LOOP_AUTO && mode != SDMD_LOOP_ONE_CTA && mode != SDMD_LOOP_SWEEP && mode != SDMD_LOOP_SWEEP_SERIAL &&
        mode != SDMD_LOOP_CLUSTER && mode != SDMD_LOOP_SWEEP_LAUNCH)
        return set_err(SDMD_ERR_BAD_ARG, "loop mode");
    c->loop_mode = mode;
    return 0;
}
void sdmd_cuda_set_timing(sdmd_ctx *c, int on) { c->timing = on < 0 ? 0 : on; }

static int bump_epoch(sdmd_ctx *c);
static int pull_ctl(sdmd_ctx *c) {
    CK(cudaMemcpyAsync(c->h_ctl, c->d_ctl, sizeof(Ctl), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    c->counters_known = true;
    return 0;
}
// the gaussian cursor and the evaluation counter only move on the device inside Predict(list) / DoEvent / the event loop;
// between those the host copy is current and the full-system pass needs no round trip to read them
static int need_counters(sdmd_ctx *c) { return c->counters_known ? 0 : pull_ctl(c); }
static int push_ctl(sdmd_ctx *c) {
    CK(cudaMemcpyAsync(c->d_ctl, c->h_ctl, sizeof(Ctl), cudaMemcpyHostToDevice, c->stream));
    return 0;
}
// Gaussian cursor / fill mark of the device control block, by VALUE in stream order.  (An asynchronous copy out of the pinned
// host mirror would read it when the copy executes, and the host moves its own cursor on as soon as a full-system pass is
// queued — the pass advances the device cursor itself.)
__global__ void k_set_rng_cursor(Ctl *ctl, i64 g_pos, i64 g_filled, int which) {
    if (which & 1) ctl->g_pos = g_pos;
    if (which & 2) ctl->g_filled = g_filled;
}
static int set_rng_cursor(sdmd_ctx *c, int which) {
    k_set_rng_cursor<<<1, 1, 0, c->stream>>>(c->d_ctl, c->h_ctl->g_pos, c->h_ctl->g_filled, which);
    CK(cudaGetLastError());
    return 0;
}
// invalidate every cached candidate list (link list rebuilt / replaced from the host)
__global__ void k_bump_epoch(Ctl *ctl) { ctl->cc_floor = ++ctl->cc_epoch; }
static int bump_epoch(sdmd_ctx *c) {   // on the device, in stream order: no host round trip
    k_bump_epoch<<<1, 1, 0, c->stream>>>(c->d_ctl);
    CK(cudaGetLastError());
    return 0;
}

// ---- state transfer -------------------------------------------------------------------------------
int sdmd_cuda_upload_coords(sdmd_ctx *c, const double *x, const double *y, const double *z, const double *vx, const double *vy,
                            const double *vz) {
    CK(cudaSetDevice(c->device));
    size_t b = ((size_t)c->n + 1) * sizeof(double);
    const double *src[6] = {x, y, z, vx, vy, vz};
    double *dst[6] = {c->P.x, c->P.y, c->P.z, c->P.vx, c->P.vy, c->P.vz};
    for (int k = 0; k < 6; k++)
        if (src[k]) CK(cudaMemcpyAsync(dst[k], src[k], b, cudaMemcpyHostToDevice, c->stream));
    c->cells_fresh = false;
    return 0;
}

int sdmd_cuda_upload_celllist(sdmd_ctx *c, const int32_t *celllist) {
    CK(cudaSetDevice(c->device));
    CK(cudaMemcpyAsync(c->P.celllist, celllist, ((size_t)c->n + c->P.ncell + 1) * sizeof(int), cudaMemcpyHostToDevice, c->stream));
    k_cell_count<<<(c->P.ncell + 255) / 256, 256, 0, c->stream>>>(c->P);
    c->launches++;
    TRY(bump_epoch(c));
    c->cells_fresh = false;
    return 0;
}

int sdmd_cuda_download(sdmd_ctx *c, struct sdmd_host_view *v) {
    CK(cudaSetDevice(c->device));
    size_t n1 = (size_t)c->n + 1;
#define DL(dst, src, cnt)                                                                                        \
    if (v->dst) CK(cudaMemcpyAsync(v->dst, c->P.src, (cnt) * sizeof(*v->dst), cudaMemcpyDeviceToHost, c->stream))
    DL(x, x, n1); DL(y, y, n1); DL(z, z, n1); DL(vx, vx, n1); DL(vy, vy, n1); DL(vz, vz, n1);
    DL(type, type, n1); DL(hb_bond, hb_bond, n1); DL(hb_itype, hb_itype, n1);
    DL(hbn_partner, hbn_partner, 5 * n1); DL(hbn_partner_num, hbn_pnum, n1); DL(hbn_status, hbn_status, n1);
    DL(ev_time, ev_time, n1); DL(ev_potential, ev_pot, n1); DL(ev_partner, ev_partner, n1);
    DL(ev_type, ev_type, n1); DL(ev_sub, ev_sub, n1); DL(cell_idx, cell_idx, 3 * n1);
    if (c->P.flow_time_rec) { DL(flow_time_rec, flow_time_rec, n1); }
#undef DL
    if (v->ev_counter) CK(cudaMemcpyAsync(v->ev_counter, c->P.ev_counter, n1 * 8, cudaMemcpyDeviceToHost, c->stream));
    if (v->ev_pcounter) CK(cudaMemcpyAsync(v->ev_pcounter, c->P.ev_pcounter, n1 * 8, cudaMemcpyDeviceToHost, c->stream));
    TRY(pull_ctl(c));
    v->currenttime = c->h_ctl->currenttime;
    v->frame = c->h_ctl->frame;
    v->target_temperature = c->h_ctl->temperature;
    return 0;
}

int sdmd_cuda_download_slots(sdmd_ctx *c, double *time, double *pot, int32_t *partner, uint8_t *type, uint8_t *sub) {
    CK(cudaSetDevice(c->device));
    size_t n1 = (size_t)c->n + 1;
    if (time) CK(cudaMemcpyAsync(time, c->P.ev_time, n1 * 8, cudaMemcpyDeviceToHost, c->stream));
    if (pot) CK(cudaMemcpyAsync(pot, c->P.ev_pot, n1 * 8, cudaMemcpyDeviceToHost, c->stream));
    if (partner) CK(cudaMemcpyAsync(partner, c->P.ev_partner, n1 * 4, cudaMemcpyDeviceToHost, c->stream));
    if (type) CK(cudaMemcpyAsync(type, c->P.ev_type, n1, cudaMemcpyDeviceToHost, c->stream));
    if (sub) CK(cudaMemcpyAsync(sub, c->P.ev_sub, n1, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    return 0;
}

int sdmd_cuda_download_celllist(sdmd_ctx *c, int32_t *celllist) {
    CK(cudaSetDevice(c->device));
    CK(cudaMemcpyAsync(celllist, c->P.celllist, ((size_t)c->n + c->P.ncell + 1) * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    return 0;
}

int sdmd_cuda_download_cbt(sdmd_ctx *c, int32_t *node, int32_t *leaf) {
    CK(cudaSetDevice(c->device));
    if (node) CK(cudaMemcpyAsync(node, c->P.cbt_node, (size_t)c->n * 2 * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    if (leaf) CK(cudaMemcpyAsync(leaf, c->P.cbt_leaf, ((size_t)c->n + 1) * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    return 0;
}

// ---- thermostat stream ----------------------------------------------------------------------------
// Maxwell_Boltzmann_Distribution(0,1), ToolFunctions.c:517-522,544-569, on the process-global rand()
double sdmd_host_gaussian_libc(void) {
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

int64_t sdmd_cuda_rng_available(const sdmd_ctx *c) {
    sdmd_ctx *m = const_cast<sdmd_ctx *>(c);
    if (need_counters(m)) return -1;   // no round trip when nothing consumed gaussians since the last one
    return c->g_filled - c->h_ctl->g_pos;
}

int sdmd_cuda_rng_feed(sdmd_ctx *c, const double *g, int64_t n) {
    CK(cudaSetDevice(c->device));
    if (n <= 0) return 0;
    TRY(need_counters(c));
    i64 cap = c->P.g_cap;
    if (c->g_filled + n - c->h_ctl->g_pos > cap) return set_err(SDMD_ERR_BAD_ARG, "gaussian ring overflow");
    i64 pos = c->g_filled % cap;
    i64 first = (n < cap - pos) ? n : cap - pos;
    CK(cudaMemcpyAsync(c->d_gauss + pos, g, first * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    if (n > first) CK(cudaMemcpyAsync(c->d_gauss, g + first, (n - first) * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    c->g_filled += n;
    c->h_ctl->g_filled = c->g_filled;
    TRY(set_rng_cursor(c, 2));
    CK(cudaStreamSynchronize(c->stream));
    return 0;
}

// move the consumption cursor (benchmarks replay the same synthetic stream; must stay within the ring)
int sdmd_cuda_rng_seek(sdmd_ctx *c, int64_t pos) {
    TRY(need_counters(c));
    if (pos < 0 || pos > c->g_filled || c->g_filled - pos > c->P.g_cap) return set_err(SDMD_ERR_BAD_ARG, "rng_seek out of range");
    c->h_ctl->g_pos = pos;
    TRY(set_rng_cursor(c, 1));
    return 0;
}

// forget everything fed so far (fresh stream)
int sdmd_cuda_rng_reset(sdmd_ctx *c) {
    TRY(need_counters(c));
    c->g_filled = 0;
    c->h_ctl->g_filled = 0;
    c->h_ctl->g_pos = 0;
    TRY(set_rng_cursor(c, 3));
    return 0;
}

int sdmd_cuda_rng_fill_libc(sdmd_ctx *c, int64_t n) {
    std::vector<double> g((size_t)(n > 0 ? n : 0));
    for (int64_t k = 0; k < n; k++) g[k] = sdmd_host_gaussian_libc();
    return sdmd_cuda_rng_feed(c, g.data(), n);
}

// ---- hot path -------------------------------------------------------------------------------------
int sdmd_cuda_pbc_full(sdmd_ctx *c) {
    CK(cudaSetDevice(c->device));
    k_pbc_full<<<grid_for(c->n, 256), 256, 0, c->stream>>>(c->P);
    c->launches++;
    c->cells_fresh = false;
    CK(cudaGetLastError());
    return 0;
}

static int cell_build_impl(sdmd_ctx *c, int parts, bool bump, bool pbc = false) {
    Timer t(c, "cell_build");
    int n = c->n, m = c->P.ncell;
    // (the histogram is all zero here: zero-filled at create, cleared again by the scan of the previous build)
    {
        Timer t2(c, "k1_index", 2);
        if (pbc) CK(sdmd_launch_pdl(k_cell_index<true>, dim3(grid_for(n, 256)), dim3(256), c->stream, c->P, c->cell_of, c->cell_count, c->cell_rank));
        else CK(sdmd_launch_pdl(k_cell_index<false>, dim3(grid_for(n, 256)), dim3(256), c->stream, c->P, c->cell_of, c->cell_count, c->cell_rank));
    }
    int nb = (m + SCAN_ITEMS - 1) / SCAN_ITEMS;
    {
        Timer t2(c, "k1_scan", 2);
        c->scan_epoch = (c->scan_epoch + 1) & 0x3fffffffu;
        if (c->scan_epoch == 0) c->scan_epoch = 1;      // (0 is what a never-written state word carries)
        CK(sdmd_launch_pdl(k_scan_cells, dim3(nb), dim3(512), c->stream, c->P, c->cell_count, c->d_cell_start, c->scan_state, c->scan_ticket, c->scan_ticket_base,
                           c->scan_epoch, m, nb));
        c->scan_ticket_base += (unsigned)nb;
    }
    {
        Timer t2(c, "k1_scatter", 2);
        CK(sdmd_launch_pdl(k_cell_scatter, dim3(grid_for(n, 256)), dim3(256), c->stream, c->P, (const int *)c->cell_of, (const int *)c->cell_rank,
                           (const int *)c->d_cell_start, c->cell_tmp));
    }
    {
        Timer t2(c, "k1_place", 2);
CK(sdmd_launch_pdl(k_cell_place, dim3(grid_for(n, 256)), dim3(256), c->stream, c->P, (const int *)c->cell_of, (const int *)c->d_cell_start,
                           (const int *)c->cell_tmp, c->d_sorted_atom, c->d_srec, c->d_sloc, c->d_scell, parts | (bump ? 4 : 0)));
    }
    c->launches += 4;      // index, scan, scatter, place
    c->cells_fresh = true;
    CK(cudaGetLastError());
    return 0;
}
int sdmd_cuda_cell_build(sdmd_ctx *c) {
    CK(cudaSetDevice(c->device));
    return cell_build_impl(c, 3, true);
}

// `late` (optional): an event the evaluation stage has to wait for — velocities and gaussians still in flight on the
// copy stream while the position-only neighbour filter runs (sdmd_cuda_prepare); h_ctl must be current
static int predict_all_impl(sdmd_ctx *c, cudaEvent_t late) {
    int n = c->n;
    i64 need = 0;
    if (c->P.thermostat_on) need = c->n_draws;  // one draw per predicted, non-constrained atom (SGThread.c:165-168, TimePrediction.c:591)
    if (c->g_filled - c->h_ctl->g_pos < need) return set_err(SDMD_ERR_BAD_ARG, "predict_all: not enough gaussians fed");
    i64 g_pos0 = c->h_ctl->g_pos;
    if (c->P.wall_dyn) {      // moving wall: a prediction can change the box the next one sees -> the reference's order, one thread
        Timer t(c, "predict_all");
        k_predict_all_serial<<<1, 1, 0, c->stream>>>(c->P);
        c->launches++;
        c->h_ctl->g_pos = g_pos0 + need;
        c->counters_known = false;
        CK(cudaGetLastError());
        return 0;
    }
    i64 g_base = g_pos0 % c->P.g_cap;       // ring offset of the first draw
    SlotBuf B = {c->b_time, c->b_pot, c->b_partner, c->b_pcounter, c->b_type, c->b_sub};
    SlotBuf C = {c->c_time, c->c_pot, c->c_partner, c->c_pcounter, c->c_type, c->c_sub};
    int grid = grid_for(n, 128), grid1 = grid_for(n, 256);
    if (!c->P.rows_mode) {    // (the row filter clears both itself: no memset between the kernels of the pass)
        CK(cudaMemsetAsync(c->pcount, 0, ((size_t)n + 1) * sizeof(int), c->stream));
        CK(cudaMemsetAsync(c->nqueue, 0, (FIX_NQ + 2) * sizeof(int), c->stream));   // queue lengths and, behind them, the evaluation-count sum
    }
    {
        Timer t(c, "predict_all");
        const bool tile = c->P.tile_tx > 0;
        if (c->P.rows_mode) {
            Timer tf(c, "k2_filter", 2);
            if (sdmd_k2_rows_filter_launch(c->P, c->pcount, c->nqueue, FIX_NQ + 2, c->stream)) return set_err(SDMD_ERR_CUDA, "k_r0_filter_rows", cudaGetLastError());
            c->launches++;
        } else if (c->P.surv && c->P.tile_mode != 1) {
            Timer tf(c, "k2_filter", 2);
            k_predict_all_filter<<<grid_for(n, 256), 256, 0, c->stream>>>(c->P);
            c->launches++;
        }
        if (late) {
            CK(cudaStreamWaitEvent(c->stream, late, 0));
            k_cell_permute<<<grid_for(n, 256), 256, 0, c->stream>>>(c->P, c->d_sorted_atom, c->d_srec, c->d_sloc, c->d_scell, 2);
            c->launches++;
        }
        // systems without bonds, hydrogen bonds and single-body sources (the GEL boxes) run the instance with those
        // predictors compiled out
        const bool lean = K2_LEAN && c->P.gel_mode && !c->P.hydrogen_bond && !c->P.single_body;
        Timer te(c, "k2_eval", 2);
        if (c->P.rows_mode) {
            Round0Out ro = {c->npair, c->second, c->bp_time, c->bp_j, c->pcount, c->pwho};
            if (sdmd_k2_rows_eval_launch(c->P, B, ro, g_base, c->stream)) return set_err(SDMD_ERR_CUDA, "k_r0_eval_rows", cudaGetLastError());
        } else if (tile) {
            Round0Out ro = {c->npair, c->second, c->bp_time, c->bp_j, c->pcount, c->pwho};
            if (sdmd_k2_tile_launch(c->P, B, ro, g_base, c->stream)) return set_err(SDMD_ERR_CUDA, "k_predict_all_tile", cudaGetLastError());
        } else if (lean)
            k_predict_all_round<true><<<grid, 128, 0, c->stream>>>(c->P, B, c->npair, c->second, c->bp_time, c->bp_j, c->pcount, c->pwho, g_base);
        else
            k_predict_all_round<false><<<grid, 128, 0, c->stream>>>(c->P, B, c->npair, c->second, c->bp_time, c->bp_j, c->pcount, c->pwho, g_base);
        c->launches++;
    }
    {
        Timer t2(c, "predict_all_fixup");
        {
            Timer t3(c, "fixup_resolve", 2);
CK(sdmd_launch_pdl(k_predict_all_resolve, dim3(grid1), dim3(256), c->stream, c->P, B, (const int *)c->pcount, (const int *)c->pwho,
                               (const double *)c->second, (const double *)c->bp_time, (const int *)c->bp_j, c->queue_a, c->nqueue, c->npair, c->d_sum));
        }
        {
            Timer t4(c, "fixup_rounds", 2);
            DevP P = c->P, Pq = c->P;
            Pq.ctl = c->d_ctl_scratch;
            int *qa = c->queue_a, *qb = c->queue_b, *nq = c->nqueue, *npair = c->npair;
            unsigned long long *sum = c->d_sum;
            i64 gb = g_base, draws = need;
            int qcap = 2 * n;
            void *args[] = {&P, &Pq, &B, &C, &qa, &qb, &nq, &npair, &sum, &gb, &draws, &qcap};
            CK(cudaLaunchCooperativeKernel((const void *)k_predict_all_fixpoint, dim3(c->fix_grid), dim3(128), args, 0, c->stream));
        }
        c->launches += 2;
    }
    // the relaxed buffer becomes the live slot arrays (pointer swap, no copy); the host keeps its own copy of the gaussian cursor,
    // the evaluation counter lives on the device
    std::swap(c->P.ev_time, c->b_time); std::swap(c->P.ev_pot, c->b_pot); std::swap(c->P.ev_partner, c->b_partner);
    std::swap(c->P.ev_pcounter, c->b_pcounter); std::swap(c->P.ev_type, c->b_type); std::swap(c->P.ev_sub, c->b_sub);
    c->h_ctl->g_pos = g_pos0 + need;
    CK(cudaGetLastError());
    return 0;
}
int sdmd_cuda_predict_all(sdmd_ctx *c) {
    CK(cudaSetDevice(c->device));
    if (!c->cells_fresh) return set_err(SDMD_ERR_STATE, "predict_all needs a fresh cell_build (cell-sorted tiles)");
    TRY(need_counters(c));
    return predict_all_impl(c, nullptr);
}

int sdmd_cuda_predict(sdmd_ctx *c, const int32_t *renew) {
    CK(cudaSetDevice(c->device));
    if (!renew) return set_err(SDMD_ERR_BAD_ARG, "null renew list");
    int cnt = 0;
    while (renew[1 + cnt]) cnt++;
    // the event loop only ever lists one or two atoms (SGThread.c:238-251); longer lists take the generic serial kernel
    if (cnt > 2) {
        int *d_list = nullptr;
        TRY(scratch_get(c, 0, sizeof(int) * cnt, (void **)&d_list));
        CK(cudaMemcpyAsync(d_list, renew + 1, sizeof(int) * cnt, cudaMemcpyHostToDevice, c->stream));
        c->counters_known = false;
        k_predict_list_n<<<1, 1, 0, c->stream>>>(c->P, d_list, cnt);
        c->launches++;
        CK(cudaStreamSynchronize(c->stream));
        CK(cudaGetLastError());
        return 0;
    }
    c->counters_known = false;
    k_predict_list<<<1, 1, 0, c->stream>>>(c->P, cnt > 0 ? renew[1] : 0, cnt > 1 ? renew[2] : 0);
    c->launches++;
    CK(cudaGetLastError());
    return 0;
}

int sdmd_cuda_cbt_build(sdmd_ctx *c) {
    CK(cudaSetDevice(c->device));
    Timer t(c, "cbt_build");
    int n = c->n;
    CK(sdmd_launch_pdl(k_cbt_leaves, dim3(grid_for(n, 256)), dim3(256), c->stream, c->P, (const int *)c->d_leaf_atom));
    c->launches++;
    if (n > 1) {
        // internal nodes 1..n-1: subtrees of height <= 11 bottom-up in one launch, then the few levels above them
        int depth = 0;
        while ((2 << depth) <= n - 1) depth++;          // level of the deepest internal node
        int L = depth > 10 ? depth - 10 : 0;
        int nsub = 1 << L;
        CK(sdmd_launch_pdl(k_cbt_subtrees, dim3(nsub), dim3(256), c->stream, c->P, L, depth - L));
        c->launches++;
        if (L > 0) {
            CK(sdmd_launch_pdl(k_cbt_top, dim3(1), dim3(256), c->stream, c->P, L));
            c->launches++;
        }
    }
    CK(cudaGetLastError());
    return 0;
}

// Prepare() of the reference (SGThread.c:44-62) from host inputs in one call.  Positions go up first on the replica's
// stream; velocities and the thermostat stream follow on a second stream while PBC("full"), LinkList("full") and the
// position-only neighbour filter of Predict(all) already run; the evaluation stage waits for them.
int sdmd_cuda_prepare(sdmd_ctx *c, const double *x, const double *y, const double *z, const double *vx, const double *vy,
                      const double *vz, const double *gauss, int64_t n_gauss) {
    CK(cudaSetDevice(c->device));
    if (!x || !y || !z || !vx || !vy || !vz) return set_err(SDMD_ERR_BAD_ARG, "prepare: null coordinate array");
    if (n_gauss < 0 || n_gauss > c->P.g_cap || (n_gauss && !gauss)) return set_err(SDMD_ERR_BAD_ARG, "prepare: gaussian count");
    if (!c->copy_stream) {
        CK(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
        CK(cudaEventCreateWithFlags(&c->ev_copy, cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&c->ev_main, cudaEventDisableTiming));
    }
    TRY(need_counters(c));
    size_t b = ((size_t)c->n + 1) * sizeof(double);
    CK(cudaMemcpyAsync(c->P.x, x, b, cudaMemcpyHostToDevice, c->stream));
    CK(cudaMemcpyAsync(c->P.y, y, b, cudaMemcpyHostToDevice, c->stream));
    CK(cudaMemcpyAsync(c->P.z, z, b, cudaMemcpyHostToDevice, c->stream));
    // the second stream starts once the positions have landed: it must not overwrite what earlier work still reads, and
    // positions should have the link to themselves so the cell build can begin as early as possible
    CK(cudaEventRecord(c->ev_main, c->stream));
    CK(cudaStreamWaitEvent(c->copy_stream, c->ev_main, 0));
    CK(cudaMemcpyAsync(c->P.vx, vx, b, cudaMemcpyHostToDevice, c->copy_stream));
    CK(cudaMemcpyAsync(c->P.vy, vy, b, cudaMemcpyHostToDevice, c->copy_stream));
    CK(cudaMemcpyAsync(c->P.vz, vz, b, cudaMemcpyHostToDevice, c->copy_stream));
    if (n_gauss) CK(cudaMemcpyAsync(c->d_gauss, gauss, n_gauss * sizeof(double), cudaMemcpyHostToDevice, c->copy_stream));
    CK(cudaEventRecord(c->ev_copy, c->copy_stream));
    c->g_filled = n_gauss;                              // fresh thermostat stream: drop what was fed before
    c->h_ctl->g_filled = n_gauss;
    c->h_ctl->g_pos = 0;
    TRY(set_rng_cursor(c, 3));
    TRY(cell_build_impl(c, 1, true, true));
    TRY(predict_all_impl(c, c->ev_copy));
    return sdmd_cuda_cbt_build(c);
}

// The same on the state that is already in HBM: PBC("full") + LinkList("full") + Predict(all) + CreateCBT = the reference's
// Prepare() (SGThread.c:44-62) as one call; the thermostat stream continues where it stands.
int sdmd_cuda_prepare_resident(sdmd_ctx *c) {
    CK(cudaSetDevice(c->device));
    TRY(need_counters(c));
    TRY(cell_build_impl(c, 3, true, true));
    TRY(predict_all_impl(c, nullptr));
    return sdmd_cuda_cbt_build(c);
}

// DoWallDyn, Events.c:502-540 (the arithmetic in the reference's order; `box` are boxDimension[1..3])
int sdmd_cuda_wall_dyn_step(sdmd_ctx *c, struct sdmd_wall_dyn *wd) {
    CK(cudaSetDevice(c->device));
    if (!wd) return set_err(SDMD_ERR_BAD_ARG, "wall_dyn_step: null argument");
    if (!c->P.wall_dyn) return set_err(SDMD_ERR_STATE, "wall_dyn_step: the context was created without a moving wall");
    TRY(pull_ctl(c));
    Ctl *h = c->h_ctl;
    double *box = h->dbox;
    wd->touch = h->wd_touch;
    const double currenttime = h->currenttime;
    if (wd->mode != 0 && currenttime >= wd->cur_time) {
        if (wd->mode == 3) {                                                                   // "pulse", :504-514
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
        if (c->P.wall_kind == SDMD_WALL_PARALLEL) box[1] += wd->sign * wd->step;               // :518-527
        else if (c->P.wall_kind == SDMD_WALL_CYLINDER) { box[1] += wd->sign * wd->step; box[2] += wd->sign * wd->step; }
        else { box[0] += wd->sign * wd->step; box[1] += wd->sign * wd->step; box[2] += wd->sign * wd->step; }
        k_wall_counter_bump<<<grid_for(c->n, 256), 256, 0, c->stream>>>(c->P);                 // :529-533
        c->launches++;
        if (((wd->orig_box[1] - box[1] >= wd->size) && wd->mode == 1) || ((wd->orig_box[1] - box[1] <= wd->step) && wd->mode == 2)) wd->mode = 0;   // :535-538
        h->wd_touch = wd->touch;
        CK(cudaMemcpyAsync(c->d_ctl->dbox, h->dbox, sizeof(h->dbox), cudaMemcpyHostToDevice, c->stream));
        CK(cudaMemcpyAsync(&c->d_ctl->wd_touch, &h->wd_touch, sizeof(int), cudaMemcpyHostToDevice, c->stream));
        c->P.wd_active = wd->mode != 0;       // once the motion has ended WallTime no longer retries (strcmp(wallDyn.mark, "no"), TimePrediction.c:678)
    }
    for (int k = 0; k < 3; k++) wd->box[k] = box[k];
    CK(cudaGetLastError());
    return 0;
}
int sdmd_cuda_wall_dyn_state(sdmd_ctx *c, double box[3], int32_t *touch) {
    CK(cudaSetDevice(c->device));
    TRY(pull_ctl(c));
    for (int k = 0; k < 3; k++) box[k] = c->P.wall_dyn ? c->h_ctl->dbox[k] : c->P.box[k];
    if (touch) *touch = c->h_ctl->wd_touch;
    return 0;
}

int sdmd_cuda_cbt_update(sdmd_ctx *c, const int32_t *renew) {
    CK(cudaSetDevice(c->device));
    int cnt = 0;
    while (renew[1 + cnt]) cnt++;
    for (int k = 0; k < cnt; k += 2) {
        k_cbt_update<<<1, 1, 0, c->stream>>>(c->P, renew[1 + k], k + 1 < cnt ? renew[2 + k] : 0);
        c->launches++;
    }
    CK(cudaGetLastError());
    return 0;
}

int sdmd_cuda_assign_job(sdmd_ctx *c, int32_t *target, int32_t *partner, double *dt, int32_t *renew) {
    CK(cudaSetDevice(c->device));
    k_assign_job<<<1, 1, 0, c->stream>>>(c->P);
    c->launches++;
    TRY(pull_ctl(c));
    if (target) *target = c->h_ctl->target;
    if (partner) *partner = c->h_ctl->partner;
    if (dt) *dt = c->h_ctl->dt;
    if (renew) {
        renew[0] = c->h_ctl->partner ? 2 : 1;
        renew[1] = c->h_ctl->target;
        renew[2] = c->h_ctl->partner;
        renew[3] = 0;
    }
    return 0;
}

// sweep blocks of the overlapped kernel: two blocks fit per SM next to the event block (register use of the event role)
static int sweep_grid(const sdmd_ctx *c) {
    int blocks = grid_for(c->n, 256);
    int cap = 148 * 8;   // measured flat between 1 and 16 blocks per SM: the iteration is bound by the event block
    return blocks < cap ? blocks : cap;
}
static int advance_grid(const sdmd_ctx *c) {
    int blocks = grid_for(c->n, 256);
    int cap = 148 * 8;
    return blocks < cap ? blocks : cap;
}

int sdmd_cuda_advance(sdmd_ctx *c, double dt) {
    CK(cudaSetDevice(c->device));
    Timer t(c, "advance");
    k_advance<<<advance_grid(c), 256, 0, c->stream>>>(c->P, dt, 0);
    c->launches++;
    CK(cudaGetLastError());
    return 0;
}

int sdmd_cuda_add_time(sdmd_ctx *c, double dt) {  // currenttime += timeInr, SGThread.c:80
    TRY(pull_ctl(c));
    c->h_ctl->currenttime += dt;
    CK(cudaMemcpyAsync(&c->d_ctl->currenttime, &c->h_ctl->currenttime, sizeof(double), cudaMemcpyHostToDevice, c->stream));
    return 0;
}

int sdmd_cuda_hazard_check(sdmd_ctx *c, int32_t target, int32_t partner, int32_t *hazard) {
    CK(cudaSetDevice(c->device));
    k_hazard<<<1, 1, 0, c->stream>>>(c->P, target, partner, (int *)c->d_sum + 2);
    c->launches++;
    CK(cudaMemcpyAsync(hazard, (int *)c->d_sum + 2, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    return 0;
}

int sdmd_cuda_cell_move(sdmd_ctx *c, int32_t atom) {
    CK(cudaSetDevice(c->device));
    k_cell_move<<<1, 1, 0, c->stream>>>(c->P, atom);
    c->launches++;
    c->cells_fresh = false;
    CK(cudaGetLastError());
    return 0;
}

int sdmd_cuda_do_event(sdmd_ctx *c) {
    CK(cudaSetDevice(c->device));
    c->counters_known = false;
    k_do_event<<<1, 1, 0, c->stream>>>(c->P);
    c->launches++;
    c->cells_fresh = false;
    CK(cudaGetLastError());
    return 0;
}

int sdmd_cuda_potential_energy(sdmd_ctx *c, double *e) {
    CK(cudaSetDevice(c->device));
    double *d = (double *)(c->d_sum + 1);
    int n = c->n;
    if (n <= 2048) {
        k_potential_energy<<<1, 128, 0, c->stream>>>(c->P, d);
        c->launches++;
    } else {
        // ordered parallel form: per-atom addend lists (count -> exclusive scan -> write), then one in-order sum
        int *cnt = c->cell_of;                       // scratch int[n+1]
        int *off = c->pcount;                        // scratch int[n+1]
        int nb = (n + SCAN_ITEMS - 1) / SCAN_ITEMS;
        int *blk = nullptr;
        TRY(scratch_get(c, 0, ((size_t)nb + 2) * sizeof(int), (void **)&blk));
        DevP Pq = c->P;
        Pq.ctl = c->d_ctl_scratch;      // the counting walk repeats the FindPair calls of the writing walk: its warnings go to the scratch block
        k_pe_count<<<grid_for(n, 128), 128, 0, c->stream>>>(Pq, cnt);
        k_scan_local<<<nb, 512, 0, c->stream>>>(cnt, off, blk, n);
        k_scan_blocks<<<1, 32, 0, c->stream>>>(blk, nb);
        k_scan_add<<<grid_for(n, 256), 256, 0, c->stream>>>(off, blk, n, nb);
        int total = 0;
        CK(cudaMemcpyAsync(&total, off + n, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
        CK(cudaStreamSynchronize(c->stream));
        double *terms = nullptr;
        TRY(scratch_get(c, 1, ((size_t)total + 32) * sizeof(double), (void **)&terms));
        k_pe_write<<<grid_for(n, 128), 128, 0, c->stream>>>(c->P, off, terms);
        k_ordered_sum<<<1, 32, 0, c->stream>>>(terms, nullptr, total, d);
        c->launches += 6;
        CK(cudaMemcpyAsync(e, d, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
        CK(cudaStreamSynchronize(c->stream));
        return 0;
    }
    CK(cudaMemcpyAsync(e, d, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    return 0;
}
int sdmd_cuda_kinetic_energy(sdmd_ctx *c, double *e) {
    CK(cudaSetDevice(c->device));
    double *d = (double *)(c->d_sum + 1);
    if (c->n <= 2048 || c->P.flow_mark >= 2) {   // per-chain centre-of-mass terms: serial kernel
        k_kinetic_energy<<<1, 1, 0, c->stream>>>(c->P, d);
        c->launches++;
    } else {
        double *terms = c->second;                   // scratch double[n+1]
        k_ke_terms<<<grid_for(c->n, 256), 256, 0, c->stream>>>(c->P, terms);
        k_ordered_sum<<<1, 32, 0, c->stream>>>(terms, nullptr, c->n, d);
        c->launches += 2;
    }
    CK(cudaMemcpyAsync(e, d, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    return 0;
}

static void fill_stats(const Ctl *h, struct sdmd_run_stats *s) {
    s->frame = h->frame; s->iterations = h->iterations; s->count_recal = h->count_recal;
    s->collision = h->n_coli; s->bond = h->n_bond; s->hb = h->n_hb; s->hb_neighbor = h->n_hbne;
    s->thermostat = h->n_ther; s->pcc = h->n_pcc; s->warnings = h->warnings;
    s->pair_predictions = h->pair_predictions; s->gaussians_used = h->g_pos; s->currenttime = h->currenttime;
    s->stop_reason = h->stop ? h->stop - 1 : 0; s->error = h->error;
    s->wall = h->n_wall; s->error_prints = h->error_prints;
    s->hb_num_formed = h->hbnumformed; s->alpha_hb_formed = h->alphaformed;
}

int sdmd_cuda_get_stats(sdmd_ctx *c, struct sdmd_run_stats *s) {
    TRY(pull_ctl(c));
    fill_stats(c->h_ctl, s);
    return 0;
}

// debug: accumulated clock64() cycles of the loop phases {head, advance, event, gather, eval, fold, cbt, count}
int sdmd_cuda_phase_cycles(sdmd_ctx *c, int64_t out[12]) {
    TRY(pull_ctl(c));
    for (int k = 0; k < 12; k++) out[k] = c->h_ctl->phase_cycles[k];
    return 0;
}

int sdmd_cuda_set_temperature(sdmd_ctx *c, double t) {
    c->h_ctl->temperature = t;
    CK(cudaMemcpyAsync(&c->d_ctl->temperature, &c->h_ctl->temperature, sizeof(double), cudaMemcpyHostToDevice, c->stream));
    return 0;
}

int sdmd_cuda_run(sdmd_ctx *c, double t_until, int64_t max_iter, struct sdmd_trace_rec *trace, int64_t trace_cap,
                  struct sdmd_run_stats *stats) {
    CK(cudaSetDevice(c->device));
    c->cells_fresh = false;
    if (trace_cap > c->trace_cap) {
        if (c->d_trace) cudaFree(c->d_trace);
        CK(cudaMalloc((void **)&c->d_trace, (size_t)trace_cap * sizeof(sdmd_trace_rec)));
        c->trace_cap = trace_cap;
    }
    TRY(pull_ctl(c));
    c->h_ctl->stop = 0;
    c->h_ctl->trace_pos = 0;
    c->h_ctl->live = 0;
    i64 iter_limit = c->h_ctl->iterations + max_iter;
    if (max_iter > (1LL << 61)) iter_limit = (1LL << 62);
    TRY(push_ctl(c));
    sdmd_trace_rec *dtr = (trace && trace_cap > 0) ? c->d_trace : nullptr;
    cudaEventRecord(c->ev0, c->stream);
    if (c->loop_mode == SDMD_LOOP_CLUSTER || (c->loop_mode == SDMD_LOOP_AUTO && c->n > SDMD_ONE_CTA_N && c->n <= SDMD_SMALL_N)) {
        c->counters_known = false;
        if (const char *e = getenv("SDMD_CLUSTER_SIZE")) c->cluster_size = atoi(e);   // measurement knob
        if (c->cluster_size < 1 || c->cluster_size > 16) return set_err(SDMD_ERR_BAD_ARG, "cluster size");
        if (c->cluster_size > 8) CK(cudaFuncSetAttribute(k_run_cluster, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(c->cluster_size); cfg.blockDim = dim3(256); cfg.dynamicSmemBytes = 0; cfg.stream = c->stream;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = c->cluster_size; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr; cfg.numAttrs = 1;
        CK(cudaLaunchKernelEx(&cfg, k_run_cluster, c->P, t_until, iter_limit, dtr, trace_cap));
        c->launches++;
        CK(cudaGetLastError());
        TRY(pull_ctl(c));
    } else if (c->loop_mode == SDMD_LOOP_ONE_CTA || (c->loop_mode == SDMD_LOOP_AUTO && c->n <= SDMD_ONE_CTA_N)) {
        c->counters_known = false;
        k_run_cta<<<1, 256, 0, c->stream>>>(c->P, t_until, iter_limit, dtr, trace_cap);
        c->launches++;
        CK(cudaGetLastError());
        TRY(pull_ctl(c));
    } else if (c->loop_mode == SDMD_LOOP_SWEEP_SERIAL) {
        c->counters_known = false;
        k_loop_head<<<1, 1, 0, c->stream>>>(c->P, t_until, iter_limit);
        c->launches++;
        const int batch = 64;
        for (;;) {
            for (int k = 0; k < batch; k++) {
                k_advance<<<advance_grid(c), 256, 0, c->stream>>>(c->P, 0.0, 1);
                k_event_exec<<<1, 256, 0, c->stream>>>(c->P, t_until, iter_limit, dtr, trace_cap);
            }
            c->launches += 2 * batch;
            CK(cudaGetLastError());
            TRY(pull_ctl(c));
            if (c->h_ctl->stop) break;
        }
    } else if (c->loop_mode != SDMD_LOOP_SWEEP_LAUNCH) {
        // overlapped sweep, the whole loop in one cooperative launch: one block per SM, block 0 the event block (k_run_sweep)
        size_t n1 = (size_t)c->n + 1;
        if (!c->x2) { TRY(dalloc(c, &c->x2, n1)); TRY(dalloc(c, &c->y2, n1)); TRY(dalloc(c, &c->z2, n1)); TRY(dalloc(c, &c->t2, n1)); }
        DevP Pw = c->P;      // parity 0: reads the live arrays, writes the spare ones
        Pw.x = c->x2; Pw.y = c->y2; Pw.z = c->z2; Pw.ev_time = c->t2;
        Pw.lz_x = c->P.x; Pw.lz_y = c->P.y; Pw.lz_z = c->P.z; Pw.lz_t = c->P.ev_time;
        Pw.lz_on = 1;
        c->counters_known = false;
        k_loop_head_job<<<1, 1, 0, c->stream>>>(c->P, t_until, iter_limit, 0);
        int sms = 0;
        CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, c->device));
        i64 seq0 = c->h_ctl->iterations;        // iteration number of the first job (k_loop_head_job publishes it in slot 0)
        unsigned long long *done = c->d_sum + 3;  // scratch counter: sweep blocks x iterations finished
        CK(cudaMemsetAsync(done, 0, sizeof(*done), c->stream));
        void *args[] = {&Pw, &seq0, &done, &t_until, &iter_limit, &dtr, &trace_cap};
        // cooperative launch: every block must be resident (they wait on each other through memory), nothing else is used of it
        CK(cudaLaunchCooperativeKernel((const void *)k_run_sweep, dim3(sms), dim3(256), args, 0, c->stream));
        c->launches += 2;
        CK(cudaGetLastError());
        TRY(pull_ctl(c));
        if (c->h_ctl->live == 1) {   // the state ended up in the spare buffer: it becomes the live one
            std::swap(c->P.x, c->x2); std::swap(c->P.y, c->y2); std::swap(c->P.z, c->z2); std::swap(c->P.ev_time, c->t2);
        }
    } else {
        // the same overlap with one launch per iteration (k_sweep_event), kept for comparison: its rate follows the host
        size_t n1 = (size_t)c->n + 1;
        if (!c->x2) { TRY(dalloc(c, &c->x2, n1)); TRY(dalloc(c, &c->y2, n1)); TRY(dalloc(c, &c->z2, n1)); TRY(dalloc(c, &c->t2, n1)); }
        DevP Pw[2];   // Pw[r]: reads buffer r, writes buffer r ^ 1; buffer 0 = the live arrays at entry
        double *bx[2] = {c->P.x, c->x2}, *by[2] = {c->P.y, c->y2}, *bz[2] = {c->P.z, c->z2}, *bt[2] = {c->P.ev_time, c->t2};
        for (int r = 0; r < 2; r++) {
            Pw[r] = c->P;
            Pw[r].x = bx[r ^ 1]; Pw[r].y = by[r ^ 1]; Pw[r].z = bz[r ^ 1]; Pw[r].ev_time = bt[r ^ 1];
            Pw[r].lz_x = bx[r]; Pw[r].lz_y = by[r]; Pw[r].lz_z = bz[r]; Pw[r].lz_t = bt[r];
            Pw[r].lz_on = 1;
        }
        c->counters_known = false;
        k_loop_head_job<<<1, 1, 0, c->stream>>>(c->P, t_until, iter_limit, 0);
        c->launches++;
        const int batch = 64;
        int r = 0;
        i64 seq = c->h_ctl->iterations;   // iteration index the next launch executes
        for (;;) {
            for (int k = 0; k < batch; k++, r ^= 1, seq++)
            {
                cudaLaunchConfig_t cfg = {};
                cfg.gridDim = dim3(1 + sweep_grid(c)); cfg.blockDim = dim3(256); cfg.dynamicSmemBytes = 0; cfg.stream = c->stream;
                cudaLaunchAttribute attr[1];
                attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
                attr[0].val.programmaticStreamSerializationAllowed = 1;
                cfg.attrs = attr; cfg.numAttrs = 1;
                CK(cudaLaunchKernelEx(&cfg, k_sweep_event, Pw[r], r, seq, t_until, iter_limit, dtr, trace_cap));
            }
            c->launches += batch;
            CK(cudaGetLastError());
            TRY(pull_ctl(c));
            if (c->h_ctl->stop) break;
        }
        if (c->h_ctl->live == 1) {   // the state ended up in the spare buffer: it becomes the live one
            std::swap(c->P.x, c->x2); std::swap(c->P.y, c->y2); std::swap(c->P.z, c->z2); std::swap(c->P.ev_time, c->t2);
        }
    }
    cudaEventRecord(c->ev1, c->stream);
    cudaEventSynchronize(c->ev1);
    float ms = 0;
    cudaEventElapsedTime(&ms, c->ev0, c->ev1);
    c->kstat["run"].ms += ms;
    c->kstat["run"].n++;
    if (dtr) {
        i64 cnt = c->h_ctl->trace_pos < trace_cap ? c->h_ctl->trace_pos : trace_cap;
        CK(cudaMemcpyAsync(trace, c->d_trace, (size_t)cnt * sizeof(sdmd_trace_rec), cudaMemcpyDeviceToHost, c->stream));
        CK(cudaStreamSynchronize(c->stream));
    }
    if (stats) fill_stats(c->h_ctl, stats);
    if (c->h_ctl->error) return set_err(c->h_ctl->error, "device-side error during run");
    return 0;
}

// Run n replicas (all on the same device) concurrently, one persistent CTA each, until every one of them stops
// (its own t_until / max_iter / RNG low / end of run).  Systems larger than SDMD_MANY_N are not eligible.
int sdmd_cuda_run_many(sdmd_ctx **ctxs, int n, const double *t_until, int64_t max_iter, struct sdmd_run_stats *stats) {
    if (n <= 0 || !ctxs) return set_err(SDMD_ERR_BAD_ARG, "run_many: no replicas");
    sdmd_ctx *c0 = ctxs[0];
    CK(cudaSetDevice(c0->device));
    std::vector<MultiArg> args((size_t)n);
    for (int k = 0; k < n; k++) {
        sdmd_ctx *c = ctxs[k];
        if (c->device != c0->device) return set_err(SDMD_ERR_BAD_ARG, "run_many: replicas on different devices");
        if (c->n > SDMD_MANY_N) return set_err(SDMD_ERR_UNSUPPORTED, "run_many: system too large for the one-CTA loop");
        c->cells_fresh = false;
        TRY(pull_ctl(c));
        c->h_ctl->stop = 0;
        c->h_ctl->trace_pos = 0;
        TRY(push_ctl(c));
        CK(cudaStreamSynchronize(c->stream));
        args[k].P = c->P;
        args[k].t_until = t_until[k];
        args[k].iter_limit = (max_iter > (1LL << 61)) ? (1LL << 62) : c->h_ctl->iterations + max_iter;
    }
    MultiArg *d_args = nullptr;
    TRY(scratch_get(c0, 2, sizeof(MultiArg) * n, (void **)&d_args));
    CK(cudaMemcpyAsync(d_args, args.data(), sizeof(MultiArg) * n, cudaMemcpyHostToDevice, c0->stream));
    cudaEventRecord(c0->ev0, c0->stream);
    for (int k = 0; k < n; k++) ctxs[k]->counters_known = false;
    {
        // up to two replicas per SM: 256-thread blocks (fastest per replica); beyond that four 128-thread blocks per SM
        // carry more replicas at once (measured: 16.9 M vs 15.5 M iterations/s at 592 replicas, slower below 296)
        int sms = 148;
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, c0->device);
        const char *force = getenv("SDMD_MULTI_THREADS");   // tests force the narrow variant with a handful of replicas
        if (force ? atoi(force) == 128 : n > 2 * sms) k_run_cta_multi<128, 4><<<n, 128, 0, c0->stream>>>(d_args);
        else k_run_cta_multi<256, 2><<<n, 256, 0, c0->stream>>>(d_args);
    }
    c0->launches++;
    cudaEventRecord(c0->ev1, c0->stream);
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(c0->stream));
    float ms = 0;
    cudaEventElapsedTime(&ms, c0->ev0, c0->ev1);
    c0->kstat["run_many"].ms += ms;
    c0->kstat["run_many"].n++;
    int rc = 0;
    for (int k = 0; k < n; k++) {
        TRY(pull_ctl(ctxs[k]));
        if (stats) fill_stats(ctxs[k]->h_ctl, &stats[k]);
        if (ctxs[k]->h_ctl->error) rc = set_err(ctxs[k]->h_ctl->error, "device-side error during run_many");
    }
    return rc;
}

// CalPotenE of n replicas in one launch
int sdmd_cuda_potential_energy_many(sdmd_ctx **ctxs, int n, double *energy) {
    if (n <= 0 || !ctxs) return set_err(SDMD_ERR_BAD_ARG, "potential_energy_many: no replicas");
    sdmd_ctx *c0 = ctxs[0];
    CK(cudaSetDevice(c0->device));
    std::vector<MultiArg> args((size_t)n);
    for (int k = 0; k < n; k++) { CK(cudaStreamSynchronize(ctxs[k]->stream)); args[k].P = ctxs[k]->P; args[k].t_until = 0; args[k].iter_limit = 0; }
    MultiArg *d_args = nullptr;
    double *d_out = nullptr;
    TRY(scratch_get(c0, 2, sizeof(MultiArg) * n, (void **)&d_args));
    TRY(scratch_get(c0, 3, sizeof(double) * n, (void **)&d_out));
    CK(cudaMemcpyAsync(d_args, args.data(), sizeof(MultiArg) * n, cudaMemcpyHostToDevice, c0->stream));
    k_potential_energy_multi<<<n, 128, 0, c0->stream>>>(d_args, d_out);
    c0->launches++;
    CK(cudaMemcpyAsync(energy, d_out, sizeof(double) * n, cudaMemcpyDeviceToHost, c0->stream));
    CK(cudaStreamSynchronize(c0->stream));
    return 0;
}

// ---- REMD -----------------------------------------------------------------------------------------
// rearrange, Server.c:145-192
int sdmd_remd_rearrange(double *epot, struct sdmd_remd_temp *T, int nrep, const double *u01, int n_u01, int *n_used,
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
                if (attempt) attempt[lo] += 1;
                bool ok = dE <= 0.0;
                if (!ok) {
                    if (used >= n_u01) return set_err(SDMD_ERR_BAD_ARG, "rearrange: out of uniform draws");
                    ok = u01[used++] < exp(-1 * dE);
                }
                if (ok) {
                    sdmd_remd_temp z = T[i]; T[i] = T[j]; T[j] = z;
                    lo = (T[i].num < T[j].num ? T[i].num : T[j].num) - 1;
                    if (accept) accept[lo] += 1;
                }
                break;
            }
        }
    }
    if (n_used) *n_used = used;
    return 0;
}

// NCCL is resolved at run time (torch's bundled libnccl.so.2 or the system one); only the five entry
// points needed for the exchange step are bound.
struct nccl_uid_val { char b[128]; };  // ncclUniqueId
struct NcclApi {
    void *lib = nullptr;
    int (*GetUniqueId)(void *) = nullptr;
    int (*CommInitRank)(void **, int, struct nccl_uid_val, int) = nullptr;
    int (*AllGather)(const void *, void *, size_t, int, void *, cudaStream_t) = nullptr;
    int (*CommDestroy)(void *) = nullptr;
    const char *(*GetErrorString)(int) = nullptr;
};
static NcclApi g_nccl;
static int load_nccl() {
    if (g_nccl.lib) return 0;
    const char *names[] = {getenv("SDMD_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
    for (const char *nm : names) {
        if (!nm) continue;
        g_nccl.lib = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
        if (g_nccl.lib) break;
    }
    if (!g_nccl.lib) return set_err(SDMD_ERR_NCCL, "cannot dlopen libnccl.so.2 (set SDMD_NCCL_LIB)");
    g_nccl.GetUniqueId = (int (*)(void *))dlsym(g_nccl.lib, "ncclGetUniqueId");
    g_nccl.CommInitRank = (int (*)(void **, int, nccl_uid_val, int))dlsym(g_nccl.lib, "ncclCommInitRank");
    g_nccl.AllGather = (int (*)(const void *, void *, size_t, int, void *, cudaStream_t))dlsym(g_nccl.lib, "ncclAllGather");
    g_nccl.CommDestroy = (int (*)(void *))dlsym(g_nccl.lib, "ncclCommDestroy");
    g_nccl.GetErrorString = (const char *(*)(int))dlsym(g_nccl.lib, "ncclGetErrorString");
    if (!g_nccl.GetUniqueId || !g_nccl.CommInitRank || !g_nccl.AllGather) return set_err(SDMD_ERR_NCCL, "libnccl lacks required symbols");
    return 0;
}

int sdmd_remd_nccl_unique_id(unsigned char id[128]) {
    TRY(load_nccl());
    int rc = g_nccl.GetUniqueId(id);
    if (rc) return set_err(SDMD_ERR_NCCL, g_nccl.GetErrorString ? g_nccl.GetErrorString(rc) : "ncclGetUniqueId");
    return 0;
}

int sdmd_remd_nccl_init(sdmd_ctx *c, int nranks, int rank, const unsigned char id[128]) {
    CK(cudaSetDevice(c->device));
    TRY(load_nccl());
    nccl_uid_val uid;
    memcpy(uid.b, id, 128);
    int rc = g_nccl.CommInitRank(&c->nccl_comm, nranks, uid, rank);
    if (rc) return set_err(SDMD_ERR_NCCL, g_nccl.GetErrorString ? g_nccl.GetErrorString(rc) : "ncclCommInitRank");
    c->nranks = nranks;
    c->rank = rank;
    if (c->d_xchg) { cudaFree(c->d_xchg); c->d_xchg = nullptr; }          // sized for the new communicator on first use
    if (c->h_xchg) { cudaFreeHost(c->h_xchg); c->h_xchg = nullptr; }
    return 0;
}

// glibc rand() restated for the replicated server stream (TYPE_3 additive feedback generator: r[i] =
// r[i-3] + r[i-31], output >> 1; seeding per srand()).  This is glibc's published algorithm, not sDMD code.
struct GlibcRand {
    int f = 3, b = 0;
    uint32_t state[31];
    void seed(uint32_t s) {
        if (s == 0) s = 1;
        int32_t word = (int32_t)s;
        state[0] = (uint32_t)word;
        for (int i = 1; i < 31; i++) {
            long hi = word / 127773, lo = word % 127773;
            word = (int32_t)(16807 * lo - 2836 * hi);
            if (word < 0) word += 2147483647;
            state[i] = (uint32_t)word;
        }
        f = 3; b = 0;
        for (int i = 0; i < 310; i++) next();
    }
    uint32_t next() {
        state[f] += state[b];
        uint32_t res = state[f] >> 1;
        f = (f + 1) % 31;
        b = (b + 1) % 31;
        return res;
    }
};
// one replicated server stream per replica context (seeded by sdmd_remd_seed_server, dropped by sdmd_cuda_destroy)
static std::map<sdmd_ctx *, GlibcRand> g_server_rng;
static std::mutex g_server_rng_mu;
static void remd_forget(sdmd_ctx *c) {      // called by sdmd_cuda_destroy
    {
        std::lock_guard<std::mutex> lk(g_server_rng_mu);
        g_server_rng.erase(c);
    }
    if (c->nccl_comm && g_nccl.CommDestroy) g_nccl.CommDestroy(c->nccl_comm);
    c->nccl_comm = nullptr;
}

// the restated generator by itself (tests compare it with libc rand() for 10^6 draws)
int sdmd_host_glibc_rand(uint32_t seed, int64_t n, uint32_t *out) {
    GlibcRand r;
    r.seed(seed);
    for (int64_t k = 0; k < n; k++) out[k] = r.next();
    return 0;
}

int sdmd_remd_seed_server(sdmd_ctx *c, uint32_t seed) {  // srand(seed) of sServer, Server.c:57-58
    std::lock_guard<std::mutex> lk(g_server_rng_mu);
    g_server_rng[c].seed(seed);
    return 0;
}

// {E, T, num} of this replica straight into the all-gather's send record: CalPotenE (ToolFunctions.c:684-752) by the same
// serial device function as sdmd_cuda_potential_energy, so the energy never visits the host before the exchange
struct XchgRec { double e; double T; int32_t num; int32_t pad; double pad2; };
static_assert(sizeof(XchgRec) == 32, "exchange record is 32 bytes");
__global__ void k_xchg_fill(DevP D, XchgRec *rec, double T, int num) {
    double e = d_potential_energy(D);
    if (threadIdx.x == 0) { rec->e = e; rec->T = T; rec->num = num; rec->pad = 0; rec->pad2 = 0; }
}

// ReplicaExchange (REMD.c:135-192) + collect/rearrange/distribute (Server.c:121-211) without the server:
// every rank all-gathers {E, T, num}, runs the identical sweep with an identically seeded stream and keeps
// its own entry.
// `epot` NULL: the energy is computed on the device and written straight into the send record (systems the serial energy
// kernel covers, natom <= 2048 — every REMD demo); otherwise *epot is what the caller computed.  On return *epot_out (if given)
// holds this replica's energy as it went into the exchange.
static int remd_exchange_impl(sdmd_ctx *c, const double *epot, struct sdmd_remd_temp *T, double *epot_out, int *active_out = nullptr) {
    CK(cudaSetDevice(c->device));
    typedef XchgRec Rec;
    int R = c->nranks;
    GlibcRand *rngp = nullptr;
    {
        std::lock_guard<std::mutex> lk(g_server_rng_mu);
        auto it = g_server_rng.find(c);
        if (it != g_server_rng.end()) rngp = &it->second;
    }
    if (!rngp) return set_err(SDMD_ERR_STATE, "remd_exchange: sdmd_remd_seed_server has not been called for this replica");
    if (!epot && c->n > 2048) return set_err(SDMD_ERR_BAD_ARG, "remd_exchange: device-side energy needs natom <= 2048; pass the energy");
    if (!c->d_xchg) CK(cudaMalloc(&c->d_xchg, (size_t)(R + 1) * sizeof(Rec)));
    if (!c->h_xchg) CK(cudaMallocHost(&c->h_xchg, (size_t)(R + 1) * sizeof(Rec)));
    Rec *all = (Rec *)c->h_xchg;
    Rec *d_send = (Rec *)c->d_xchg + R, *d_recv = (Rec *)c->d_xchg;
    if (epot) {
        all[R] = Rec{*epot, T->T, T->num, 0, 0.0};
        CK(cudaMemcpyAsync(d_send, &all[R], sizeof(Rec), cudaMemcpyHostToDevice, c->stream));
    } else {
        k_xchg_fill<<<1, 128, 0, c->stream>>>(c->P, d_send, T->T, T->num);
        c->launches++;
    }
    if (R == 1 || !c->nccl_comm) {
        CK(cudaMemcpyAsync(all, d_send, sizeof(Rec), cudaMemcpyDeviceToHost, c->stream));
    } else {
        int rc = g_nccl.AllGather(d_send, d_recv, sizeof(Rec), /*ncclChar*/ 0, c->nccl_comm, c->stream);
        if (rc) return set_err(SDMD_ERR_NCCL, g_nccl.GetErrorString ? g_nccl.GetErrorString(rc) : "ncclAllGather");
        CK(cudaMemcpyAsync(all, d_recv, sizeof(Rec) * R, cudaMemcpyDeviceToHost, c->stream));
    }
    CK(cudaStreamSynchronize(c->stream));
    if (epot_out) *epot_out = all[c->rank].e;
    GlibcRand &rng = *rngp;
    std::vector<double> e(R);
    std::vector<sdmd_remd_temp> t(R);
    int active = 0;
    for (int k = 0; k < R; k++) { e[k] = all[k].e; t[k].num = all[k].num; t[k].pad_ = 0; t[k].T = all[k].T; active += all[k].T != 0; }
    if (active_out) *active_out = active;
    // draw lazily: copy the generator, count how many the sweep used, then advance the real one
    GlibcRand probe = rng;
    std::vector<double> u(2 * R + 2);
    for (auto &x : u) x = 0 + (double)probe.next() / ((double)2147483647 / (1 - 0));  // RANDOMVALUE(0,1), Server.c:16-18
    int used = 0;
    TRY(sdmd_remd_rearrange(e.data(), t.data(), R, u.data(), (int)u.size(), &used, nullptr, nullptr));
    for (int k = 0; k < used; k++) rng.next();
    *T = t[c->rank];
    if (T->T != 0) TRY(sdmd_cuda_set_temperature(c, T->T));
    return 0;
}
// End of a replica's run (REMD.c:119-122 + Server.c:108-113): the reference's replica sends the server a (E = 0, T = 0) sentinel
// and leaves; the server keeps serving rounds to the replicas still running and skips the ones that left.  A collective needs
// every rank in every round, so a replica that is done keeps contributing its sentinel — one all-gather per round of the
// others — until every entry is a sentinel (the server's CheckEnd); the sweep skips sentinels exactly like Server.c:150-153,
// and every rank keeps drawing the same server stream.
int sdmd_remd_finish(sdmd_ctx *c) {
    if (c->nranks <= 1 || !c->nccl_comm) return 0;
    for (;;) {
        double zero = 0;
        struct sdmd_remd_temp t = {0, 0, 0.0};
        int active = 0;
        TRY(remd_exchange_impl(c, &zero, &t, nullptr, &active));
        if (active == 0) return 0;
    }
}
int sdmd_remd_exchange(sdmd_ctx *c, double epot, struct sdmd_remd_temp *T, uint32_t /*unused*/) { return remd_exchange_impl(c, &epot, T, nullptr); }
int sdmd_remd_exchange_energy(sdmd_ctx *c, struct sdmd_remd_temp *T, double *epot) { return remd_exchange_impl(c, nullptr, T, epot); }

}  // extern "C"
