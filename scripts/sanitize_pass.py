"""Small full-system passes for compute-sanitizer (memcheck / racecheck): sparse and dense GEL boxes, so the row filter, the cp.async
evaluation, the overflow fallback, the resolve stage and the warp-per-atom re-evaluation rounds all run.
    compute-sanitizer --tool memcheck python scripts/sanitize_pass.py"""
import sys
sys.path.insert(0, ".")
from sdmd_b200 import capi, synth
for n_axis, a in ((16, 8.0), (14, 5.0), (12, 4.4)):
    hv, rng = synth.gel_box(n_axis, a=a, jitter=0.3, seed=11)
    r = capi.Replica(hv, device=0)
    r.rng_feed(rng.standard_normal(hv.natom + 50000))
    for _ in range(2):
        r.rng_seek(0); r.prepare_resident()
    st, _ = r.run(max_iter=200)
    print(n_axis, a, r.stats()["pair_predictions"], st["frame"], r.kernel_stats("predict_all_requeued"))
    r.close()
