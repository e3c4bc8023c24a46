// sdmd_host.h — launchers that live in their own translation units (one object per .cu, see sdmd_b200/capi.py::build)
#pragma once
#include "sdmd_dev.cuh"

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
int sdmd_k2_rows_filter_launch(const DevP &P, cudaStream_t stream);
int sdmd_k2_rows_eval_launch(const DevP &P, SlotBuf dst, Round0Out out, long long g_base, cudaStream_t stream);
