/*
 * sdmd_oracle.h — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Plain-C CPU restatement of sDMD's event-prediction / event-scheduling hot path, operating on the same
 * flattened `struct sdmd_host_view` the C-ABI takes.  It exists only to check the CUDA path (tests/,
 * __graft_entry__.smoke(), bench.py's cpu_baseline leg); nothing under sdmd_b200/ may call it.
 *
 * Parity status: PINNED.  tests/test_oracle_vs_reference.py checks every function here against dumps of
 * the unmodified reference (oracle/_ref/ref_harness, built by oracle/Makefile from the reference's own
 * sources) — event slots after the full Predict, link list, CBT, 3000-iteration event traces on ala10
 * and on the GEL box, and CalPotenE — bit for bit.
 */
#ifndef SDMD_ORACLE_H
#define SDMD_ORACLE_H
#include "sdmd_cuda.h"

typedef struct oracle_ctx oracle_ctx;

/* The context aliases the arrays of `v` and mutates them in place. */
oracle_ctx *oracle_create(struct sdmd_host_view *v);
void oracle_destroy(oracle_ctx *c);
void oracle_rng_feed(oracle_ctx *c, const double *gauss, int64_t n);
void oracle_cell_grid(const oracle_ctx *c, int cellnum[3], double cellsize[3]);
void oracle_pbc_full(oracle_ctx *c);
void oracle_cell_build(oracle_ctx *c);
void oracle_set_celllist(oracle_ctx *c, const int32_t *celllist);
const int32_t *oracle_celllist(const oracle_ctx *c);
void oracle_predict(oracle_ctx *c, const int32_t *renew_list);
void oracle_predict_all(oracle_ctx *c);
void oracle_cbt_build(oracle_ctx *c);
void oracle_cbt_update(oracle_ctx *c, const int32_t *renew_list);
const int32_t *oracle_cbt_node(const oracle_ctx *c);
const int32_t *oracle_cbt_leaf(const oracle_ctx *c);
void oracle_advance(oracle_ctx *c, double dt);
int  oracle_run(oracle_ctx *c, double t_until, int64_t max_iter, struct sdmd_trace_rec *trace, int64_t trace_cap,
                struct sdmd_run_stats *stats);
double oracle_potential_energy(oracle_ctx *c);
double oracle_kinetic_energy(oracle_ctx *c);
void oracle_get_stats(const oracle_ctx *c, struct sdmd_run_stats *stats);
void oracle_set_temperature(oracle_ctx *c, double t);
/* Server.c:145-192 restated (same signature as sdmd_remd_rearrange) */
int oracle_remd_rearrange(double *epot, struct sdmd_remd_temp *T, int nrep, const double *u01, int n_u01, int *n_used,
                          double *attempt, double *accept);
/* ToolFunctions.c:544-569 with libc rand(): one standard-normal draw */
double oracle_gaussian_libc(void);
#endif
