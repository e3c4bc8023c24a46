"""Ad-hoc GPU probe: large GEL box parity vs the oracle and first timings (not a benchmark)."""
import sys, time
import numpy as np
sys.path.insert(0, '.')
from oracle.oracle_py import Oracle
from sdmd_b200 import capi
from sdmd_b200.synth import gel_box

def probe_parity(n_axis=32, iters=2000):
    hv0, rng = gel_box(n_axis, jitter=0.3)
    gauss = rng.standard_normal(hv0.natom + 8000)
    ho = hv0.copy(); o = Oracle(ho); o.rng_feed(gauss)
    hg = hv0.copy(); r = capi.Replica(hg, device=0); r.rng_feed(gauss)
    o.pbc_full(); o.cell_build(); o.predict_all(); o.cbt_build()
    r.pbc_full(); r.cell_build(); r.predict_all(); r.cbt_build()
    so, to = o.run(max_iter=iters, trace_cap=iters)
    sg, tg = r.run(max_iter=iters, trace_cap=iters)
    print('oracle', so); print('gpu   ', sg)
    m = min(len(to), len(tg))
    print('trace prefix equal', to[:m].tobytes() == tg[:m].tobytes(), len(to), len(tg))
    print('rounds', r.kernel_stats('predict_all_rounds'), 'run ms', r.kernel_stats('run'))

def probe_time(n_axis, a=8.0):
    hv, rng = gel_box(n_axis, a=a, jitter=0.3)
    n = hv.natom
    r = capi.Replica(hv, device=0)
    r.rng_feed(rng.standard_normal(n + 100000))
    for rep in range(3):
        r.reset_kernel_stats()
        r.pbc_full(); r.cell_build(); r.predict_all(); r.cbt_build()
        for _ in range(20): r.advance(1e-6)
        out = {k: r.kernel_stats(k) for k in ('cell_build','predict_all','predict_all_fixup','predict_all_rounds','cbt_build','advance')}
        if rep == 2:
            st = r.stats()
            print(f'N={n} a={a}', {k: (round(v[0]/max(v[1],1),4), v[1]) for k,v in out.items()}, 'pairs/atom', st['pair_predictions']/(3*n))
            adv = out['advance'][0]/out['advance'][1]
            print(f'   advance {88*n/adv/1e6:.1f} GB/s ; predict_all {88*n/(out["predict_all"][0])/1e6:.1f} GB/s-equivalent, {st["pair_predictions"]/3/out["predict_all"][0]/1e6:.2f} Gpairs/s')
        r.rng_feed(rng.standard_normal(n))
    t0=time.time(); sg, _ = r.run(max_iter=2000); dt=time.time()-t0
    print(f'   run 2000 iterations: {dt*1e3:.1f} ms wall -> {2000/dt:.0f} it/s', sg['frame'], sg['stop_reason'])
    r.close()

if __name__ == '__main__':
    probe_parity()
    for n_axis in (20, 40, 100):
        probe_time(n_axis)
    probe_time(100, a=5.0)
