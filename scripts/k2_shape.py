"""Locality probe for the full-system pass: the same 10^6 GEL atoms in a cube (100^3) and in a slab (25 x 25 x 1600).  In the
x-major cell order a cube's y / z neighbour rows lie 8 KB / 728 KB away in the cell-sorted arrays, the slab's 2 KB / 45 KB."""
import sys
sys.path.insert(0, ".")
import numpy as np
from sdmd_b200 import capi, synth
from sdmd_b200.hostview import HostView, NT


def box(nx, ny, nz, a=8.0, jitter=0.3, seed=1575263273, temperature=0.6):
    n = nx * ny * nz
    hv = HostView(n)
    hv.box[:] = (nx * a, ny * a, nz * a)
    hv.timestep = 1e9
    hv.gel_mode = 1
    rng = np.random.default_rng(seed)
    k = np.arange(n)
    hv.x[1:] = (k % nx + 0.5) * a
    hv.y[1:] = ((k // nx) % ny + 0.5) * a
    hv.z[1:] = (k // (nx * ny) + 0.5) * a
    for arr in (hv.x, hv.y, hv.z):
        arr[1:] += rng.uniform(-jitter, jitter, n)
    hv.mass[1:] = 12.0
    hv.type[1:] = synth.GEL_TYPE
    hv.aa_type[1:] = 21
    hv.protein_num[1:] = 1
    hv.aa_num[1:] = k + 1
    sd = np.sqrt(temperature / 12.0)
    for arr in (hv.vx, hv.vy, hv.vz):
        v = rng.standard_normal(n) * sd
        v -= v.mean()
        v[v == 0] = sd
        arr[1:] = v
    hv.tables = synth.gel_table()
    hv.coll_index[synth.GEL_TYPE * NT + synth.GEL_TYPE] = 0
    return hv, rng


for shape in ((100, 100, 100), (25, 25, 1600), (10, 10, 10000)):
    hv, rng = box(*shape)
    g = rng.standard_normal(hv.natom)
    r = capi.Replica(hv, device=0)
    r.set_timing(2)
    r.rng_feed(g)
    for it in range(8):
        if it == 3:
            r.reset_kernel_stats()
        r.rng_seek(0)
        r.pbc_full(); r.cell_build(); r.predict_all(); r.cbt_build()
    r.synchronize()
    row = {k: r.kernel_stats(k) for k in ("cell_build", "predict_all", "predict_all_fixup")}
    print(shape, r.cellnum, {k: round(v[0] / max(v[1], 1) * 1e3, 1) for k, v in row.items()}, "us")
    r.close()
