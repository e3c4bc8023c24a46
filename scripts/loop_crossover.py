"""Where do the one-block loop (whole event loop in one persistent CTA, advance included) and the one-cluster loop (the other
blocks of the cluster help with the advance) lose to the sweep driver (grid-wide advance overlapped with one event block per
launch)?  GEL boxes of growing size, iterations/s per driver."""
import sys, time
sys.path.insert(0, ".")
import numpy as np
from sdmd_b200 import capi, synth

iters = 20000
for n_axis in (12, 16, 20, 25, 32, 40, 50):
    hv, rng = synth.gel_box(n_axis, a=8.0, jitter=0.3)
    g = rng.standard_normal(hv.natom + 40 * iters)
    row = []
    for mode in (1, 4, 2):
        r = capi.Replica(hv.copy(), device=0)
        r.set_loop_mode(mode)
        r.rng_feed(g)
        r.pbc_full(); r.cell_build(); r.predict_all(); r.cbt_build()
        r.run(max_iter=2000)
        r.synchronize()
        t0 = time.time()
        st, _ = r.run(max_iter=iters)
        r.synchronize()
        dt = time.time() - t0
        row.append(iters / dt)
        r.close()
    print(f"N={hv.natom:7d}  one block {row[0]:9.0f} it/s ({1e6/row[0]:6.1f} us)   cluster of 8 {row[1]:9.0f} it/s ({1e6/row[1]:6.1f} us)   sweep {row[2]:9.0f} it/s ({1e6/row[2]:6.1f} us)")
