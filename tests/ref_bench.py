"""TEST INFRASTRUCTURE: run the UNMODIFIED reference (oracle/_ref/ref_bench_gel, see oracle/ref_bench.c) on a GEL box of any
size — 10^6 atoms included — and hand back what its full-system pass started from and what it produced, so the oracle port and
the CUDA path can be compared with the reference ITSELF at the size the bench numbers are quoted on."""
from __future__ import annotations

import shutil
import subprocess
import tempfile
from pathlib import Path

import numpy as np

from sdmd_b200.hostview import HostView, PER_ATOM

ROOT = Path(__file__).resolve().parents[1]
REFDIR = ROOT / "oracle" / "_ref"
SLOT_DTYPE = np.dtype([("time", "<f8"), ("potential", "<f8"), ("partner", "<i4"), ("type", "<i4"), ("sub", "<i4"), ("pad", "<i4"),
                       ("pcounter", "<i8")])


def available() -> bool:
    return (REFDIR / "ref_bench_gel").exists() and (REFDIR / "ref_harness_gel").exists()


def write_jittered_gel_gro(path: Path, n: int, a_nm: float, jitter_nm: float, seed: int):
    """n^3 GEL atoms on a simple-cubic lattice (CreateGELCoordinate layout, ToolFunctions.c:775-798) with a uniform jitter;
    free-format serials so that more than 99 999 atoms parse (Initialization.c:906 reads "%i%5c%5c%s")."""
    rng = np.random.default_rng(seed)
    k = np.arange(n ** 3)
    xyz = np.stack([(k % n + 0.5) * a_nm, ((k // n) % n + 0.5) * a_nm, (k // (n * n) + 0.5) * a_nm], axis=1)
    if jitter_nm:
        xyz += rng.uniform(-jitter_nm, jitter_nm, xyz.shape)
    with open(path, "w") as f:
        f.write("Model\n%5i\n" % (n ** 3))
        lines = ["%d%-5s%-5s %d %8.3f%8.3f%8.3f\n" % (i + 1, "GEL", "GEL", i + 1, *xyz[i]) for i in range(n ** 3)]
        f.write("".join(lines))
        f.write("%10.5f%10.5f%10.5f\n" % (n * a_nm, n * a_nm, n * a_nm))


def run(n_axis: int, a_nm: float = 0.8, jitter_nm: float = 0.03, passes: int = 1, seed: int = 7, want_state: bool = True):
    """Returns (host view the pass starts from | None, gaussians | None, slots after the last pass, seconds per pass)."""
    from oracle.make_golden import prepare_input, run_case
    work = Path(tempfile.mkdtemp(prefix="s", dir="/tmp"))       # short path: the reference keeps file names in small buffers
    try:
        inp = prepare_input(work, time_units=10.0)
        gro = f"GEL{n_axis}.gro"
        write_jittered_gel_gro(inp / gro, n_axis, a_nm, jitter_nm, seed)
        (work / "out").mkdir(exist_ok=True)
        cmd = [str(REFDIR / "ref_bench_gel"), str(passes), str(work / "slots.bin")]
        if want_state:
            cmd.append(str(work / "state.bin"))
        cmd += ["--", "-f", gro, "-i", str(inp), "-o", str(work / "out")]
        r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE)
        if r.returncode != 0:
            raise RuntimeError(f"ref_bench_gel failed ({r.returncode}): {r.stderr[-2000:].decode(errors='replace')}")
        secs = [float(ln.split()[3]) for ln in r.stdout.decode(errors="replace").splitlines() if ln.startswith("pass ")]
        slots = np.fromfile(work / "slots.bin", dtype=SLOT_DTYPE)
        if not want_state:
            return None, None, slots, secs
        hv, gauss = view_from_state(work / "state.bin", work)
        return hv, gauss, slots, secs
    finally:
        shutil.rmtree(work, ignore_errors=True)


def view_from_state(state_path: Path, work: Path):
    """Host view + thermostat draws of the pass ref_bench_gel dumped (`work`: scratch directory for the template run)."""
    from oracle.make_golden import run_case
    raw = np.fromfile(state_path, dtype=np.uint8)
    n = int(raw[:8].view(np.int64)[0])
    f8 = raw[8:].view(np.float64)
    box = f8[:3].copy()
    cols = [f8[3 + k * n: 3 + (k + 1) * n] for k in range(7)]
    # static description (force-field tables, type maps, scalars) from the reference's own glue on a tiny box of the same kind
    snaps = run_case("gelT", work / "t", harness="ref_harness_gel", coord="GEL3.gro", max_iter=1, n_gauss=1, extra=[], gel_n=3,
                     gel_a_nm=0.9)
    tpl = HostView.from_snapshot(snaps[0])
    hv = HostView(n)
    for name in ("cutoffr", "thermoF", "target_temperature", "currenttime", "timestep", "thermostat_on", "method_ding",
                 "hydrogen_bond", "gel_mode", "frame"):
        setattr(hv, name, getattr(tpl, name))
    hv.hb_barrier = tpl.hb_barrier.copy()
    hv.box = box
    hv.tables, hv.coll_index, hv.hb_index = tpl.tables.copy(), tpl.coll_index.copy(), tpl.hb_index.copy()
    for name, _, k in PER_ATOM:                 # every atom is the same GEL particle: broadcast atom 1 of the template
        src = getattr(tpl, name)
        if len(src) >= 2 * k:
            getattr(hv, name).reshape(-1, k)[1:] = src.reshape(-1, k)[1]
    hv.aa_num[1:] = np.arange(1, n + 1)
    hv.protein_mass[1:] = hv.mass[1] * n        # one "protein" holding every atom
    for arr, col in zip((hv.x, hv.y, hv.z, hv.vx, hv.vy, hv.vz), cols[:6]):
        arr[1:] = col
    return hv, cols[6].copy()
