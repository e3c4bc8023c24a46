"""Per-phase cycle counts of the one-block event loop on the 64 x Abeta16-22 box (start state from the reference harness)."""
import sys, time, tempfile, shutil
sys.path.insert(0, ".")
from pathlib import Path
import numpy as np
from sdmd_b200 import capi
from sdmd_b200.hostview import HostView
from oracle.make_golden import run_case

work = Path(tempfile.mkdtemp(prefix="s", dir="/tmp"))
snaps = run_case("abeta64", work, harness="ref_harness", coord="abeta64.gro", max_iter=10, n_gauss=400000, extra=[], gel_n=-8)
pre = HostView.from_snapshot(snaps[0])
shutil.rmtree(work, ignore_errors=True)
for mode in (1, 4, 2):
    r = capi.Replica(pre.copy(), device=0)
    r.set_loop_mode(mode)
    r.rng_feed(pre.extra["gauss"])
    r.pbc_full(); r.cell_build(); r.predict_all(); r.cbt_build()
    r.run(max_iter=5000)
    p0 = r.phase_cycles()
    t0 = time.time()
    st, _ = r.run(max_iter=40000)
    dt = time.time() - t0
    p1 = r.phase_cycles()
    it = max(p1["iterations"] - p0["iterations"], 1)
    print(f"mode {mode}: {40000/dt:.0f} it/s ({dt/40000*1e6:.1f} us/it)  cycles/iteration:",
          {k: int((p1[k] - p0[k]) / it) for k in p1 if k != "iterations" and not k.startswith("x")})
    r.close()
