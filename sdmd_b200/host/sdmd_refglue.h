/*
 * sdmd_refglue.h — host-side C binding between the sDMD driver and libsdmd_cuda.so.
 *
 * This translation unit is compiled TOGETHER with the reference's own sources (it includes the
 * reference's DMD.h, found on the include path at build time — it is never copied into this repo).
 * It flattens the pointer-rich process globals of DMD.h:475-613 into the pointer-free
 * `struct sdmd_host_view` that the C-ABI takes, and scatters device results back so that the kept
 * reference code (DataSave.c, Events.c) keeps working on `atom[]`.
 */
#ifndef SDMD_REFGLUE_H
#define SDMD_REFGLUE_H

#include "sdmd_cuda.h"

#ifdef __cplusplus
extern "C" {
#endif

/* Build a host view from the reference globals (after InputData(), DMD.c:196).  Arrays are malloc'ed
 * and owned by the view; release with sdmd_glue_free.  Returns 0 or SDMD_ERR_*. */
int  sdmd_glue_flatten(struct sdmd_host_view *v);
void sdmd_glue_free(struct sdmd_host_view *v);
/* Refresh only the mutable per-atom arrays of an existing view from atom[] (cheap re-flatten). */
void sdmd_glue_gather_dynamic(struct sdmd_host_view *v);
/* Write the mutable per-atom arrays of the view back into atom[] (coordinates, velocities, cell
 * index, event slot, HB state, typeofAtom). */
void sdmd_glue_scatter_dynamic(const struct sdmd_host_view *v);

/* ---- snapshot files (tagged arrays; read by sdmd_b200/snapshot.py) ---- */
typedef struct sdmd_snap sdmd_snap;
sdmd_snap *sdmd_snap_open(const char *path);
void sdmd_snap_put(sdmd_snap *s, const char *name, const char *dtype, const void *data, uint64_t count);
void sdmd_snap_put_view(sdmd_snap *s, const struct sdmd_host_view *v);
void sdmd_snap_close(sdmd_snap *s);

#ifdef __cplusplus
}
#endif
#endif
