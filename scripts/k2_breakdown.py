"""Fine-grained CUDA-event breakdown of the full-system pass on the 1M-atom box (timing level 2)."""
import sys, time
sys.path.insert(0, ".")
import numpy as np
from sdmd_b200 import capi, synth

n_axis = int(sys.argv[1]) if len(sys.argv) > 1 else 100
a = float(sys.argv[2]) if len(sys.argv) > 2 else 8.0
hv, rng = synth.gel_box(n_axis, a=a, jitter=0.3, seed=1575263273)
g = rng.standard_normal(hv.natom)
r = capi.Replica(hv, device=0)
r.set_timing(2)
r.rng_feed(g)
for it in range(8):
    if it == 3: r.reset_kernel_stats()
    r.rng_seek(0)
    r.pbc_full(); r.cell_build(); r.predict_all(); r.cbt_build()
r.synchronize()
for k in ("cell_build", "k1_index", "k1_scan", "k1_scatter", "k1_place", "predict_all", "k2_filter", "k2_eval", "predict_all_fixup", "fixup_resolve", "fixup_rounds", "cbt_build"):
    v = r.kernel_stats(k)
    print(f"{k:28s} total {v[0]:9.4f} ms  n={v[1]:4d}  avg {v[0]/max(v[1],1)*1e3:9.2f} us")
pc = r.phase_cycles()
print("fix-up rounds, block 0, cycles per pass: evaluate %.0f  barrier %.0f  commit %.0f  barrier %.0f" % tuple(pc[k] / 8 for k in ("event_thread0", "wait_sweep", "x10", "x11")))
