"""Cycles per loop iteration by phase for the large-system loop driver (k_advance + k_event_exec)."""
import sys
sys.path.insert(0, ".")
import numpy as np
from sdmd_b200 import capi, synth

n_axis = int(sys.argv[1]) if len(sys.argv) > 1 else 100
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 4000
hv, rng = synth.gel_box(n_axis, a=8.0, jitter=0.3, seed=1575263273)
r = capi.Replica(hv, device=0)
r.rng_feed(rng.standard_normal(hv.natom + 200000))
r.pbc_full(); r.cell_build(); r.predict_all(); r.cbt_build()
r.run(max_iter=500)
p0 = np.array(list(r.phase_cycles().values())); s0 = r.stats()
r.reset_kernel_stats()
st, _ = r.run(max_iter=iters)
p1 = np.array(list(r.phase_cycles().values()))
d = p1 - p0
n = max(d[7], 1)
names = ["pick-up", "next AssignJob", "event||gather", "gather(after)", "evaluate", "fold", "cbt", "(count)", "event only"]
for k, nm in enumerate(names):
    if k != 7: print(f"{nm:16s} {d[k] / n:9.0f} cycles/iteration")
ms, cnt = r.kernel_stats("run")
print(f"iterations {st['iterations'] - s0['iterations']}, events {st['frame'] - s0['frame']}, {ms * 1e3 / n:.2f} us/iteration wall on device")
print({k: st[k] for k in ("collision", "bond", "hb", "thermostat", "pcc", "count_recal")})
