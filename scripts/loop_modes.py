"""us per iteration and cycles per phase of the 1M-atom loop for the large-system drivers (2 persistent, 5 per-launch, 3 serial)."""
import sys
sys.path.insert(0, ".")
import numpy as np
from sdmd_b200 import capi, synth

n_axis = int(sys.argv[1]) if len(sys.argv) > 1 else 100
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 4000
names = ["pick-up", "next AssignJob", "event||gather", "gather(after)", "evaluate", "fold", "cbt", "(count)", "event only", "wait for sweep"]
for mode in (2, 5, 3):
    hv, rng = synth.gel_box(n_axis, a=8.0, jitter=0.3, seed=1575263273)
    r = capi.Replica(hv, device=0)
    r.set_loop_mode(mode)
    r.rng_feed(rng.standard_normal(hv.natom + 200000))
    r.pbc_full(); r.cell_build(); r.predict_all(); r.cbt_build()
    r.run(max_iter=500)
    p0 = np.array(list(r.phase_cycles().values())); s0 = r.stats()
    r.reset_kernel_stats()
    st, _ = r.run(max_iter=iters)
    p1 = np.array(list(r.phase_cycles().values()))
    d = p1 - p0
    n = max(d[7], 1)
    ms, cnt = r.kernel_stats("run")
    print(f"mode {mode}: {ms * 1e3 / n:.2f} us/iteration | " + ", ".join(f"{nm} {d[k] / n:.0f}" for k, nm in enumerate(names) if k != 7))
    r.close()
