"""Variants of sdmd_k2_rows.cu (round-0 filter / evaluation for bond-free systems): only that translation unit is recompiled
with the variant's -D flags and linked against the cached objects of the shipped build (build/obj).
    python scripts/rows_variants.py build name=-DRF_B=4,-DRF_MINB=5 ...      (CPU box)
    python scripts/rows_variants.py run [n_axis] > gpurun_out/rows_variants.txt   (GPU box: scripts/k2_breakdown.py per variant)"""
import subprocess, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from sdmd_b200 import capi
VDIR = ROOT / "build" / "variants"
CHILD = """
import sys
from pathlib import Path
sys.path.insert(0, '.')
from sdmd_b200 import capi
capi.LIB_PATH = Path(sys.argv[1])
sys.argv = ['k2_breakdown.py'] + sys.argv[2:]
exec(open('scripts/k2_breakdown.py').read())
"""

def build():
    capi.build()
    VDIR.mkdir(parents=True, exist_ok=True)
    for a in sys.argv[2:]:
        name, flags = a.split("=", 1)
        obj = VDIR / f"rows_{name}.o"
        cmd = ["nvcc", *[f for f in capi.NVCC_FLAGS if f != "-shared"], "-Xptxas", "-v", *[f for f in flags.split(",") if f],
               f"-I{ROOT / 'include'}", f"-I{capi.CSRC}", "-c", str(capi.CSRC / "sdmd_k2_rows.cu"), "-o", str(obj)]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode: print(r.stderr); raise SystemExit(1)
        print(name, [ln.strip() for ln in r.stderr.splitlines() if "Used" in ln or "spill" in ln])
        objs = [str(capi.OBJ_DIR / (n + ".o")) if n != "sdmd_k2_rows.cu" else str(obj) for n in capi.SOURCES]
        r = subprocess.run(["nvcc", "-shared", "-gencode", "arch=compute_100a,code=sm_100a", *objs, "-o", str(VDIR / f"lib_{name}.so"), "-ldl"], capture_output=True, text=True)
        if r.returncode: print(r.stderr); raise SystemExit(1)
        obj.unlink()

def run():
    for so in sorted(VDIR.glob("lib_*.so")):
        print(f"=== {so.stem}", flush=True)
        r = subprocess.run([sys.executable, "-c", CHILD, str(so), *sys.argv[2:]], cwd=ROOT, capture_output=True, text=True, timeout=300)
        print("\n".join(l for l in r.stdout.splitlines() if l.startswith(("predict_all ", "k2_"))) + r.stderr[-1500:], flush=True)

if __name__ == "__main__":
    {"build": build, "run": run}[sys.argv[1]]()
