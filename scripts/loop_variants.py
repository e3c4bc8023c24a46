"""A/B of build variants of the persistent large-system loop (SWEEP_UNROLL = atoms in flight per sweep thread): builds each
variant to build/variants/, runs scripts/loop_modes.py's mode-2 measurement with it."""
import os, subprocess, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
VDIR = ROOT / "build" / "variants"
CHILD = r"""
import sys, numpy as np
sys.path.insert(0, '.')
from sdmd_b200 import capi
from pathlib import Path
capi.LIB_PATH = Path(sys.argv[1])
from sdmd_b200 import synth
hv, rng = synth.gel_box(100, a=8.0, jitter=0.3, seed=1575263273)
r = capi.Replica(hv, device=0)
r.rng_feed(rng.standard_normal(hv.natom + 200000))
r.pbc_full(); r.cell_build(); r.predict_all(); r.cbt_build()
r.run(max_iter=500)
p0 = np.array(list(r.phase_cycles().values())); r.reset_kernel_stats()
st, _ = r.run(max_iter=4000)
d = np.array(list(r.phase_cycles().values())) - p0
ms, _ = r.kernel_stats('run')
print(f"{ms*1e3/max(d[7],1):.2f} us/iteration, wait for sweep {d[9]/max(d[7],1):.0f} cycles, event||gather {d[2]/max(d[7],1):.0f}, evaluate {d[4]/max(d[7],1):.0f}")
"""
if sys.argv[1] == "build":
    from sdmd_b200 import capi
    VDIR.mkdir(parents=True, exist_ok=True)
    for u in (1, 2, 3, 4, 6):
        capi.build(extra_flags=[f"-DSWEEP_UNROLL={u}"], out=VDIR / f"lib_sweep_u{u}.so")
else:
    for so in sorted(VDIR.glob("lib_sweep_u*.so")):
        r = subprocess.run([sys.executable, "-c", CHILD, str(so)], cwd=ROOT, capture_output=True, text=True, timeout=300)
        print(so.stem, r.stdout.strip(), r.stderr[-500:] if r.returncode else "", flush=True)
