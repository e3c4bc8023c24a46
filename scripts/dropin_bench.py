"""Config 3 through the reference's own driver: stock sDMD (CPU, one core) vs the drop-in driver with the CUDA hot path
(sdmd_b200/host/_build/sDMD_cuda) on the 64 x Abeta16-22 box, same input files, whole-program wall clock.  Start-up
(input parsing, SDEnergyMin, CUDA initialisation) is removed by differencing two run lengths.

usage: python scripts/dropin_bench.py [t_short t_long]
"""
import re
import subprocess
import sys
import tempfile
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from oracle.make_golden import PEPTIDES, REFDIR, prepare_input, write_multi_ala10  # noqa: E402

t_short, t_long = (float(sys.argv[1]), float(sys.argv[2])) if len(sys.argv) > 2 else (2.0, 12.0)
KEYS = ("collision times", "bond times", "HB times", "HB Neighbor times", "thermostat times", "PBC&CC times", "Wall times")


def run(binary, tu):
    work = Path(tempfile.mkdtemp(prefix="d", dir="/tmp"))
    inp = prepare_input(work, time_units=tu)
    src, copies, spacing, box = PEPTIDES["abeta64"]
    write_multi_ala10(inp / "abeta64.gro", ROOT / "tests" / "golden" / src, copies_per_axis=copies, spacing_nm=spacing, box_nm=box)
    t0 = time.time()
    r = subprocess.run([str(binary), "-f", "abeta64.gro", "-i", str(inp), "-o", str(work / "out")], capture_output=True, text=True)
    dt = time.time() - t0
    assert r.returncode == 0, r.stdout[-2000:]
    out = r.stdout.replace("\r", "\n")
    ev = sum(int(re.search(re.escape(k) + r"\s*=\s*(\d+)", out).group(1)) for k in KEYS)
    recal = int(re.search(r"ReCalculate times\s*=\s*(\d+)", out).group(1))
    return dt, ev, recal, (work / "out" / "out_trj.gro").read_bytes()


for name, binary in (("reference sDMD (1 CPU core)", REFDIR / "sDMD_ref"), ("drop-in sDMD_cuda (B200)", ROOT / "sdmd_b200" / "host" / "_build" / "sDMD_cuda_fixedseed")):
    a = run(binary, t_short)
    b = run(binary, t_long)
    ev, it, dt = b[1] - a[1], (b[1] + b[2]) - (a[1] + a[2]), b[0] - a[0]
    print(f"{name:32s} N=4608  {t_long - t_short:g} time units: {ev} events, {it} loop iterations in {dt:.2f} s "
          f"-> {ev / dt:,.0f} events/s, {it / dt:,.0f} iterations/s   (start-up {a[0]:.1f} s)")
    if "reference" in name:
        ref_trj = b[3]
    else:
        print("trajectory identical to the reference's:", b[3] == ref_trj)
