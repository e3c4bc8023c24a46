"""Run N loop iterations of the ala10 golden input on cuda:0 (profiling driver)."""
import sys, time
import numpy as np
sys.path.insert(0, '.')
from sdmd_b200 import capi
from sdmd_b200.hostview import HostView
n_iter = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
pre = HostView.from_snapshot('tests/golden/ala10_pre.npz')
hv = pre.copy(); hv.timestep = 1e12
r = capi.Replica(hv, device=0)
g = np.random.default_rng(1).standard_normal(hv.natom + 4 * n_iter + 1000)
r.rng_feed(g)
r.pbc_full(); r.cell_build(); r.predict_all(); r.cbt_build()
r.run(max_iter=200)
pc0 = r.phase_cycles()
t0 = time.perf_counter(); st, _ = r.run(max_iter=n_iter); dt = time.perf_counter() - t0
print(f"{n_iter} iterations in {dt*1e3:.1f} ms -> {n_iter/dt:.0f} it/s, {dt/n_iter*1e6:.2f} us/it; events {st['frame']}")
pc = r.phase_cycles()
it = pc["iterations"] - pc0["iterations"]
print({k: round((pc[k] - pc0[k]) / it) for k in pc if k != "iterations"}, "cycles/iteration")
