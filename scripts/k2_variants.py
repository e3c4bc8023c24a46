"""Occupancy sweep of the two round-0 kernels of the full-system Predict (K2): the same source compiled with
different `__launch_bounds__` minimum-blocks (-DK2_MINBLOCKS / -DK2F_MINBLOCKS), each timed with
scripts/k2_breakdown.py's CUDA-event breakdown on the 1M-atom box.

Build the variants on the CPU box first (they travel with the snapshot under build/variants/):
    python scripts/k2_variants.py build            # or: build base=-DK2_DIRECT=0,-DK2_LEAN=0 d2=-DK2_DIRECT=2
then on the GPU:
    python scripts/k2_variants.py run > gpurun_out/k2_variants.txt
"""
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
VDIR = ROOT / "build" / "variants"
VARIANTS = [(4, 5), (5, 5), (6, 5), (8, 5), (6, 6), (6, 8), (5, 6)]   # (evaluation kernel, filter kernel)

CHILD = """
import sys
from pathlib import Path
sys.path.insert(0, '.')
from sdmd_b200 import capi
capi.LIB_PATH = Path(sys.argv[1])
sys.argv = ['k2_breakdown.py']
exec(open('scripts/k2_breakdown.py').read())
"""


def build():
    """No further arguments: the launch-bounds sweep.  Otherwise each argument is `name=-DFLAG=V,-DFLAG2=W` (one variant)."""
    sys.path.insert(0, str(ROOT))
    from sdmd_b200 import capi
    VDIR.mkdir(parents=True, exist_ok=True)
    jobs = [(f"k2_{k2}_f{k2f}", [f"-DK2_MINBLOCKS={k2}", f"-DK2F_MINBLOCKS={k2f}"]) for k2, k2f in VARIANTS]
    if len(sys.argv) > 2:
        jobs = [(a.split("=", 1)[0], [f for f in a.split("=", 1)[1].split(",") if f]) for a in sys.argv[2:]]
    for name, flags in jobs:      # every translation unit with the variant's flags (capi.build compiles them in parallel)
        capi.build(extra_flags=flags, out=VDIR / f"lib_{name}.so")


def run():
    for so in sorted(VDIR.glob("lib_*.so")):      # whatever was built: the sweep above or hand-built experiment variants
        print(f"=== {so.stem}", flush=True)
        r = subprocess.run([sys.executable, "-c", CHILD, str(so)], cwd=ROOT, capture_output=True, text=True, timeout=300)
        print(r.stdout + r.stderr[-2000:], flush=True)


if __name__ == "__main__":
    {"build": build, "run": run}[sys.argv[1]]()
