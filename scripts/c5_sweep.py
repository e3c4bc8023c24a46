"""SURVEY 8d C5 roofline sweep: synthetic GEL boxes N in {8k, 64k, 512k, 1M} x lattice a in {8, 5} A x step e = +0.2 / -0.2.
Per configuration: the full-system pass (PBC + LinkList + Predict(all) + CreateCBT) on the GPU with per-stage CUDA-event times and
the HBM roofline fractions of the stages, the device-resident event loop (us per iteration), and the CPU oracle port on one host
core for both (the port is pinned bit-exact to the reference at 10^6 atoms).  Writes a markdown table to stdout."""
import json, sys, time
sys.path.insert(0, ".")
import numpy as np
from sdmd_b200 import capi, synth
from oracle.oracle_py import Oracle

peak = json.load(open("MEASURED_PEAKS.json"))["hbm_gbs"] if __import__("os").path.exists("MEASURED_PEAKS.json") else 6650.0
rows = []
for n_axis in (20, 40, 80, 100):
    for a in (8.0, 5.0):
        for e in (0.2, -0.2):
            hv, rng = synth.gel_box(n_axis, a=a, jitter=0.3, step_e=e, seed=1575263273)
            n = hv.natom
            g = rng.standard_normal(n + 400000)
            r = capi.Replica(hv.copy(), device=0)
            r.set_timing(1)      # stage brackets only: per-kernel brackets (level 2) put event records between the kernels of a stage
            r.rng_feed(g)
            for it in range(8):
                if it == 3:
                    r.synchronize(); p0 = r.stats()["pair_predictions"]; r.reset_kernel_stats(); r.timer_begin()
                r.rng_seek(0)
                r.prepare_resident()       # PBC + LinkList + Predict(all) + CreateCBT, one call (PBC inside the K1 bracket)
            ms = r.timer_end() / 5
            pairs = (r.stats()["pair_predictions"] - p0) / 5
            st = {k: r.kernel_stats(k)[0] / 5 for k in ("cell_build", "predict_all", "predict_all_fixup", "cbt_build")}
            req = r.kernel_stats("predict_all_requeued")
            # loop
            r.run(max_iter=300)
            r.reset_kernel_stats()
            s0 = r.stats()
            s1, _ = r.run(max_iter=3000)
            run_ms, _ = r.kernel_stats("run")
            its = s1["iterations"] - s0["iterations"]
            us_it = 1e3 * run_ms / its
            ev_s = (s1["frame"] - s0["frame"]) / (run_ms * 1e-3)
            r.close()
            # CPU port, one core
            ho = hv.copy(); o = Oracle(ho); o.rng_feed(g)
            t0 = time.perf_counter(); o.pbc_full(); o.cell_build(); o.predict_all(); o.cbt_build(); t_pass = time.perf_counter() - t0
            cpairs = o.stats()["pair_predictions"]
            cpu_it = 200 if n > 100000 else 2000
            t0 = time.perf_counter(); so, _ = o.run(max_iter=cpu_it); t_loop = time.perf_counter() - t0
            o.close()
            assert cpairs == round(pairs), (cpairs, pairs)
            rows.append((n, a, e, ms, pairs / ms / 1e6, st, 88.0 * n / (st["predict_all"] * 1e-3) / 1e9 / peak, 144.0 * n / (st["cell_build"] * 1e-3) / 1e9 / peak,
                         req[0] / max(req[1], 1), us_it, ev_s, 88.0 * n / (us_it * 1e-6) / 1e9 / peak, pairs / t_pass / 1e6, so["frame"] / t_loop))
            print(f"# done N={n} a={a} e={e}: pass {ms:.3f} ms, loop {us_it:.2f} us/it", file=sys.stderr, flush=True)

print("| N | a (A) | e | pass ms | G pairs/s | K1 ms | K2 round 0 ms | fix-up ms | K5 ms | K2 frac of HBM | K1 frac of HBM | queued atoms | loop us/iter | events/s | sweep frac of HBM | CPU port M pairs/s (1 core) | CPU port events/s |")
print("|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|")
for (n, a, e, ms, gp, st, f2, f1, req, us_it, ev_s, fs, cp, cev) in rows:
    print(f"| {n} | {a:g} | {e:+g} | {ms:.3f} | {gp:.1f} | {st['cell_build']:.3f} | {st['predict_all']:.3f} | {st['predict_all_fixup']:.3f} | {st['cbt_build']:.3f} | "
          f"{f2:.3f} | {f1:.3f} | {req:.0f} | {us_it:.2f} | {ev_s:.0f} | {fs:.3f} | {cp:.1f} | {cev:.0f} |")
