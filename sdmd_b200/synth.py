"""Synthetic DMD systems of the shapes BASELINE.json names.

``gel_box`` builds the reference's GEL particle system (atom type 32, m = 12, hard core 4.0 A and one
potential step at 6.5 A — the shipped row ``GEL GEL 4.0 6.5 0.2`` of InteractionPotentialTable.txt; pass
``step_e=-0.2`` for the attractive square-well variant) on a simple-cubic lattice exactly like
CreateGELCoordinate (ToolFunctions.c:775-798): position ((k + 0.5) * a), box n * a.  Velocities are
sqrt(T/m) * N(0,1) with the centre-of-mass velocity removed (GenerateVelocity, Initialization.c:1850-1881,
single "protein").  The gaussian stream for the Andersen thermostat comes from numpy here; parity runs feed
the same array to the oracle and to the CUDA path.
"""
from __future__ import annotations

import numpy as np

from .hostview import HostView, NT, STEP_TABLE_DTYPE

GEL_TYPE = 32


def gel_table(core=4.0, step_d=6.5, step_e=0.2):
    t = np.zeros(1, dtype=STEP_TABLE_DTYPE)
    t["dmin2"] = core * core
    t["dmax2"] = 0.0
    t["nstep"] = 1
    t["d2"][0, 0] = step_d * step_d
    t["e"][0, 0] = step_e
    t["acc"][0, 0] = step_e  # CalAccumulatedPoten, ToolFunctions.c:897-919
    return t


def gel_box(n_axis: int, a: float = 8.0, temperature: float = 0.6, seed: int = 1575263273, step_e: float = 0.2,
            timestep: float = 1e9, jitter: float = 0.0, wall_pad: float | None = None) -> tuple[HostView, np.random.Generator]:
    """n_axis^3 GEL particles.  Returns (host view, rng) — draw thermostat gaussians from ``rng``.

    ``wall_pad`` (A) confines the box between two smooth parallel plates normal to y (WallExist smooth / WallType
    parallel with a GEL-type wall atom, ReadWall Initialization.c:1907-1935): the box grows by ``wall_pad`` on either side
    in y so the outermost lattice planes sit a/2 + wall_pad from the plates."""
    n = n_axis ** 3
    hv = HostView(n)
    hv.box[:] = n_axis * a
    hv.cutoffr = 9.0
    hv.thermoF = 6.0
    hv.target_temperature = temperature
    hv.timestep = timestep
    hv.thermostat_on = 1
    hv.method_ding = 1
    hv.hydrogen_bond = 1
    hv.gel_mode = 1
    rng = np.random.default_rng(seed)
    k = np.arange(n)
    m_, n_, i_ = k % n_axis, (k // n_axis) % n_axis, k // (n_axis * n_axis)  # x fastest, like the reference writer
    hv.x[1:] = (m_ + 0.5) * a
    hv.y[1:] = (n_ + 0.5) * a
    hv.z[1:] = (i_ + 0.5) * a
    if jitter:
        for arr in (hv.x, hv.y, hv.z):
            arr[1:] += rng.uniform(-jitter, jitter, n)
    if wall_pad is not None:
        hv.box[1] += 2 * wall_pad
        hv.y[1:] += wall_pad
        hv.wall_kind = 1
        hv.wall_type = GEL_TYPE
        hv.wall_pos[:] = (0.0, hv.box[1] / 2, 0.0)
    hv.mass[1:] = 12.0
    hv.type[1:] = GEL_TYPE
    hv.aa_type[1:] = 21
    hv.protein_num[1:] = 1
    hv.aa_num[1:] = k + 1
    sd = np.sqrt(temperature / 12.0)
    for arr in (hv.vx, hv.vy, hv.vz):
        v = rng.standard_normal(n) * sd
        v -= v.mean()  # equal masses: COM removal
        v[v == 0] = sd
        arr[1:] = v
    hv.tables = gel_table(step_e=step_e)
    hv.coll_index[GEL_TYPE * NT + GEL_TYPE] = 0
    return hv, rng
