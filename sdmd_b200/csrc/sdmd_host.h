// sdmd_host.h — launchers that live in their own translation units (one object per .cu, see sdmd_b200/capi.py::build)
#pragma once
#include <utility>
#include "sdmd_dev.cuh"

// Programmatic dependent launch for the kernels of the full-system pass: a kernel launched with sdmd_launch_pdl may have its
// blocks resident (parked in griddepcontrol.wait) while the kernel in front of it in the stream drains, which takes the
// launch latency and the tail of the previous grid out of the ~4 us gap between two dependent launches.  Every such kernel
// starts with SDMD_PDL_ENTRY() before it touches global memory; launched the ordinary way the two instructions do nothing.
#define SDMD_PDL_ENTRY() asm volatile("griddepcontrol.launch_dependents;\n\tgriddepcontrol.wait;" ::: "memory")
template <typename... KArgs, typename... Args>
static inline cudaError_t sdmd_launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, cudaStream_t stream, Args &&...args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = 0; cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}

// outputs of round 0 of the full-system Predict besides the slot buffer (see k_predict_all_round / k_predict_all_resolve)
struct Round0Out {
    int *npair;          // pair evaluations per atom (TimePrediction.c:105)
    double *second;      // value of the best candidate that did not win (or second-best pair candidate of a partner-less winner)
    double *bp_time;     // best pair candidate in front of PCC
    int *bp_j;
    int *pcount, *pwho;  // how many lower-index atoms point at me, and one of them
};

// sdmd_k2_tile.cu: round 0 as one tile kernel (TMA-staged cell tiles in shared memory) for bond-free systems.
// configure: choose the tile shape for this cell grid / density and fill DevP::tile_*; returns 0 when the tile kernel is
// usable for this system, 1 when it is not (the two-kernel path stays), < 0 on a CUDA error.
int sdmd_k2_tile_configure(DevP &P, int device);
int sdmd_k2_tile_launch(const DevP &P, SlotBuf dst, Round0Out out, long long g_base, cudaStream_t stream);

// sdmd_k2_rows.cu: round 0 for bond-free systems as a row-batched filter (predicated independent loads, survivors in row order)
// plus an evaluation kernel that gathers survivor records by cp.async and folds them order-independently.
// configure: 0 when usable for this system (sets DevP::rows_mode), 1 when not (the kernels of sdmd_cuda.cu stay).
int sdmd_k2_rows_configure(DevP &P);
// `zero` / `nzero`: words the filter clears for the stages behind it (who-points-at-whom counters, queue lengths), so the pass needs no memset
int sdmd_k2_rows_filter_launch(const DevP &P, int *pcount, int *zero, int nzero, cudaStream_t stream);
int sdmd_k2_rows_eval_launch(const DevP &P, SlotBuf dst, Round0Out out, long long g_base, cudaStream_t stream);
