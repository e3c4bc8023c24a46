"""Throughput of R independent ala10 replicas on one GPU (one persistent CTA each, single launch)."""
import sys, time
import numpy as np
sys.path.insert(0, '.')
from sdmd_b200 import capi
from sdmd_b200.hostview import HostView
n_iter = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
pre = HostView.from_snapshot('tests/golden/ala10_pre.npz')
for R in (int(a) for a in (sys.argv[2].split(",") if len(sys.argv) > 2 else "1,8,37,74,148,296,592".split(","))):
    reps = []
    for k in range(R):
        hv = pre.copy(); hv.timestep = 1e12
        r = capi.Replica(hv, device=0)
        r.rng_feed(np.random.default_rng(k).standard_normal(hv.natom + 4 * n_iter + 4000))
        r.pbc_full(); r.cell_build(); r.predict_all(); r.cbt_build()
        reps.append(r)
    capi.run_many(reps, 1e300, max_iter=500)
    t0 = time.perf_counter(); st = capi.run_many(reps, 1e300, max_iter=n_iter); dt = time.perf_counter() - t0
    its = sum(s['iterations'] for s in st) - 500 * R
    ev = sum(s['frame'] for s in st)
    print(f"R={R:4d}: {its/dt/1e3:9.1f} k it/s aggregate ({its/dt/R/1e3:6.1f} k per replica), wall {dt*1e3:.0f} ms")
    for r in reps: r.close()
