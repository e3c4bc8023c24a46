#!/usr/bin/env python
"""TEST INFRASTRUCTURE — regenerate golden vectors from the UNMODIFIED reference.

Runs oracle/_ref/ref_harness (reference objects + our dump harness, built by oracle/Makefile) on the
reference's own inputs and stores what the reference computed as compressed .npz fixtures:

  ala10        demo/Folding-style ala10.gro, T*=0.6, 3000 loop iterations from t=0
  ala10_hb     same run, snapshot taken after 170000 iterations, then 8000 traced iterations
               (contains HB form/break and HB-neighbour events)
  gel10        GEL step-potential spheres, 10^3 simple-cubic lattice at 8 A (reference CreateGELCoordinate
               layout, ToolFunctions.c:775-798), 3000 iterations

  abeta8       8 x Abeta16-22 (KLVFFAE, synthesised all-atom coordinates) on a 2x2x2 grid, snapshot after 150000
               iterations, then 8000 traced (BASELINE config 3 in small)

  ala10x8_wallp  8 ala10 chains between two smooth parallel plates (WallExist smooth / WallType parallel, plates normal
               to y, 38 A apart; Wall.txt atom type CA because the shipped PHO is only valid in CG mode): WallTime /
               WallEvent, PBC in x and z only
  ala10x8_obst   8 ala10 chains between two flat obstruction walls ("Obstruct? 2 y .. , y ..") and three hard spherical
               obstacles ("SphObst? 3 ..."), periodic box, snapshot after 20000 iterations, then 8000 traced
  ala10x4_tunnel 4 ala10 chains, two of them inside a tunnel between two holed obstruction walls ("Tunnel? 1 ..."), snapshot
               after 60000 iterations, then 8000 traced
  ala10_flowv / ala10_flowf  ala10 with "Flow velocity ..." / "Flow force ..." (drift terms of ThermostatEvent and CalKinetE)
  ala10x4_wallc  4 ala10 chains in a smooth cylinder along x (WallType cylinder, diameter 26 A), snapshot after 20000
               iterations, then 8000 traced

usage: python oracle/make_golden.py [--out tests/golden] [--work oracle/_ref/work] [--only NAME]
Also importable: run_case(name, workdir) returns the three snapshot paths for ad-hoc (uncommitted) runs.
"""
from __future__ import annotations

import argparse
import shutil
import subprocess
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from sdmd_b200.hostview import read_snapshot  # noqa: E402

REFDIR = ROOT / "oracle" / "_ref"

CASES = {
    # name: (harness, coordinate file, max_iter, n_gauss, extra harness args, gel lattice n)
    "ala10": ("ref_harness", "ala10.gro", 3000, 8000, [], 0),
    "ala10_hb": ("ref_harness", "ala10.gro", 8000, 16000, ["--warm", "170000"], 0),
    "gel10": ("ref_harness_gel", "GEL10.gro", 3000, 12000, [], 10),
    # 8 ala10 chains on a 2x2x2 grid (stand-in for the multi-peptide aggregation config: inter-chain collisions and
    # hydrogen bonds, several proteins, N = 472); 4000 iterations
    "ala10x8": ("ref_harness", "ala10x8.gro", 4000, 16000, [], -8),
    # smooth confining walls
    "ala10x8_wallp": ("ref_harness", "ala10x8w.gro", 8000, 40000, [], -8),
    "ala10x4_wallc": ("ref_harness", "ala10x4c.gro", 8000, 40000, ["--warm", "20000"], -8),
    "ala10x8_obst": ("ref_harness", "ala10x8o.gro", 8000, 40000, ["--warm", "20000"], -8),
    # BASELINE config 3 in small: 8 x Abeta16-22 (KLVFFAE; coordinates synthesised by oracle/make_peptide_gro.py, committed as
    # tests/golden/abeta16_22.gro): charged / aromatic / branched side chains, side-chain hydrogen bonds, CONSTRn tables
    "abeta8": ("ref_harness", "abeta8.gro", 8000, 40000, ["--warm", "150000"], -8),
    "ala10x4_tunnel": ("ref_harness", "ala10x4t.gro", 8000, 40000, ["--warm", "60000"], -8),
    # flow terms in ThermostatEvent / CalKinetE: constant drift velocity, constant force
    "ala10_flowv": ("ref_harness", "ala10.gro", 6000, 16000, [], 0),
    "ala10_flowf": ("ref_harness", "ala10.gro", 6000, 16000, ["--warm", "20000"], 0),
    # one charge source (+ its periodic image), -COM: ChargeTime / ChargeEvent, centre-of-mass term in ThermostatEvent
    "ala10_charge": ("ref_harness", "ala10.gro", 8000, 20000, ["--warm", "20000"], 0),
}
# "Flow" line of parameter.txt: constant drift velocity (A per time unit) / constant force (pN)
FLOWS = {"ala10_flowv": "velocity 0.02 0.0 0.01", "ala10_flowf": "force 5.0 0.0 0.0",
         "ala10_charge": "charge 1 : 100.0 75.0 75.0 0.05 0.2"}
SDMD_ARGS = {"ala10_charge": ["-COM"],      # extra command-line switches of the reference
             "ala10x8_wallp_compress": ["-Wsz", "5", "-Wrt", "4"], "ala10x8_wallp_pulse": ["-Wsz", "2", "-Wrt", "3"]}
# moving smooth wall ("WallDyn" line of parameter.txt; the wall travels -Wsz A in -Wrt time units): live cases only
WALLDYN = {"ala10x8_wallp_compress": "compress", "ala10x8_wallp_pulse": "pulse"}
# name: (single-chain .gro under tests/golden, chains per axis, spacing nm, box nm); abeta64 = SURVEY 8d C3 (live tests only)
PEPTIDES = {
    "abeta8": ("abeta16_22.gro", (2, 2, 2), 2.0, (6.0, 6.0, 6.0)),
    "abeta64": ("abeta16_22.gro", (4, 4, 4), 2.5, (10.0, 10.0, 10.0)),
}
# name: (WallType, chains per axis, spacing nm, box nm[, "Obstruct?" value, "SphObst?" value])
WALLS = {
    "ala10x8_wallp": ("parallel", (2, 2, 2), 1.9, (5.0, 3.8, 5.0)),
    "ala10x8_wallp_compress": ("parallel", (2, 2, 2), 1.9, (5.0, 3.8, 5.0)),
    "ala10x8_wallp_pulse": ("parallel", (2, 2, 2), 1.9, (5.0, 3.8, 5.0)),
    "ala10x4_wallc": ("cylinder", (4, 1, 1), 2.0, (8.0, 2.6, 2.6)),
    # two flat obstruction walls normal to y (1.5 A inside the periodic box faces) and three hard spheres of diameter 5 A,
    # each 0.4 A off the outermost atom of a chain (positions in A, parameter.txt lines 14 and 16); no smooth wall
    "ala10x8_obst": (None, (2, 2, 2), 1.9, (5.0, 4.2, 5.0), "2 y 1.5 , y 40.5",
                     "3 19.82 17.45 24.49 5.0 , 26.29 29.64 14.62 5.0 , 31.27 10.75 25.88 5.0"),
    # a tunnel of diameter 24 A along x between two obstruction walls (x = 20 A and 60 A, the walls have holes at its mouths):
    # two of the four chains sit inside the tunnel, two in the open space on either side
    "ala10x4_tunnel": (None, (4, 1, 1), 2.0, (8.0, 2.6, 2.6), "2 x 20.0 , x 60.0", None, "1 20.0 60.0 13.0 13.0 24.0"),
}


def write_gel_gro(path: Path, n: int, a_nm: float = 0.8):
    """Same text format as CreateGELCoordinate (ToolFunctions.c:775-798) for an n^3 lattice."""
    with open(path, "w") as f:
        f.write("Model\n%5i\n" % (n ** 3))
        k = 1
        for i in range(n):
            for j in range(n):
                for m in range(n):
                    if n ** 3 < 10000:
                        f.write("%5iGEL%5s%5i%8.3f%8.3f%8.3f\n" % (k, "GEL", k, (m + 0.5) * a_nm, (j + 0.5) * a_nm, (i + 0.5) * a_nm))
                    else:
                        # five-digit serials run into the atom-name column of the fixed-width layout; the reader itself
                        # (Initialization.c:906, "%i%5c%5c%s") only needs a free-format residue number, two 5-character name
                        # fields and a whitespace-separated serial
                        f.write("%d%-5s%-5s %d %8.3f%8.3f%8.3f\n" % (k, "GEL", "GEL", k, (m + 0.5) * a_nm, (j + 0.5) * a_nm, (i + 0.5) * a_nm))
                    k += 1
        f.write("%10.5f%10.5f%10.5f\n" % (n * a_nm, n * a_nm, n * a_nm))


def write_multi_ala10(path: Path, src: Path, copies_per_axis=2, spacing_nm: float = 2.4, box_nm=9.0):
    """nx*ny*nz translated copies of the reference's ala10.gro (a single int means a cube); residue numbers restart per
    chain, which is how the reference's .gro reader tells chains apart (Initialization.c:949-960)."""
    lines = src.read_text().splitlines()
    atoms = lines[2:2 + int(lines[1])]
    xyz = np.array([[float(a[20:28]), float(a[28:36]), float(a[36:44])] for a in atoms])
    xyz -= xyz.mean(axis=0)
    cp = np.array([copies_per_axis] * 3 if np.isscalar(copies_per_axis) else copies_per_axis, dtype=int)
    box = np.array([box_nm] * 3 if np.isscalar(box_nm) else box_nm, dtype=float)
    out, serial = [], 1
    for cz in range(cp[2]):
        for cy in range(cp[1]):
            for cx in range(cp[0]):
                shift = (np.array([cx, cy, cz]) + 0.5) * spacing_nm + (box - cp * spacing_nm) / 2
                for a, p in zip(atoms, xyz + shift):
                    out.append("%s%5d%8.3f%8.3f%8.3f" % (a[:15], serial, p[0], p[1], p[2]))
                    serial += 1
    with open(path, "w") as f:
        f.write("%d x ala10\n%5d\n" % (int(cp.prod()), len(out)))
        f.write("\n".join(out) + "\n")
        f.write("%10.5f%10.5f%10.5f\n" % tuple(box))


def prepare_input(work: Path, time_units: float = 1000.0, wall_type: str | None = None, obst: str | None = None,
                  sphobst: str | None = None, tunnel: str | None = None, flow: str | None = None, wall_dyn: str | None = None) -> Path:
    """Writable copy of the reference input directory with the run length (and a smooth wall) set."""
    src = REFDIR / "input"
    if not src.exists():
        raise FileNotFoundError(f"{src} missing: run `make -C oracle ref` where /root/reference exists")
    inp = work / "input"
    if inp.exists():
        shutil.rmtree(inp)
    shutil.copytree(src, inp)
    p = inp / "parameter.txt"
    lines = p.read_text().splitlines()
    lines[1] = "Time               %g" % time_units
    if wall_type:
        lines[9] = "WallExist?         smooth"
        lines[10] = "WallType           %s" % wall_type
    if wall_dyn:
        lines[11] = "WallDyn            %s" % wall_dyn
    if obst:
        lines[13] = "Obstruct?          %s" % obst
    if sphobst:
        lines[15] = "SphObst?           %s" % sphobst
    if tunnel:
        lines[14] = "Tunnel?            %s" % tunnel
    if flow:
        lines[12] = "Flow               %s" % flow
    if wall_type or obst:
        # the wall's atom type must be a valid AtomModel name (Models.c:129-209); the shipped PHO is CG-only
        (inp / "Library_Ding" / "Wall.txt").write_text("Wall\n   ATOM CA      CA    12.0  N N\n")
    p.write_text("\n".join(lines) + "\n")
    (work / "out").mkdir(parents=True, exist_ok=True)
    return inp


def run_case(name: str, work: Path, harness=None, coord=None, max_iter=None, n_gauss=None, extra=None, gel_n=None, gel_a_nm=0.8):
    spec = CASES.get(name)
    if spec is not None:
        harness, coord, max_iter, n_gauss, extra, gel_n = [s if o is None else o for s, o in
                                                           zip(spec, (harness, coord, max_iter, n_gauss, extra, gel_n))]
    work.mkdir(parents=True, exist_ok=True)
    wall = WALLS.get(name)
    inp = prepare_input(work, wall_type=wall[0] if wall else None, obst=wall[4] if wall and len(wall) > 4 else None,
                        sphobst=wall[5] if wall and len(wall) > 5 else None, tunnel=wall[6] if wall and len(wall) > 6 else None,
                        flow=FLOWS.get(name), wall_dyn=WALLDYN.get(name))
    if wall:
        write_multi_ala10(inp / coord, inp / "ala10.gro", copies_per_axis=wall[1], spacing_nm=wall[2], box_nm=wall[3])
    elif name in PEPTIDES:
        src, copies, spacing, box = PEPTIDES[name]
        write_multi_ala10(inp / coord, ROOT / "tests" / "golden" / src, copies_per_axis=copies, spacing_nm=spacing, box_nm=box)
    elif gel_n and gel_n > 0:
        write_gel_gro(inp / coord, gel_n, gel_a_nm)
    elif gel_n and gel_n == -80:
        # 4 x 4 x 5 chains, N = 4720: past the one-thread-block limit of the device loop (live test only, not a fixture)
        write_multi_ala10(inp / coord, inp / "ala10.gro", copies_per_axis=(4, 4, 5), box_nm=(10.0, 10.0, 12.5))
    elif gel_n and gel_n < 0:
        write_multi_ala10(inp / coord, inp / "ala10.gro")
    prefix = work / name
    cmd = [str(REFDIR / harness), str(prefix), str(max_iter), str(n_gauss), *extra, "--",
           "-f", coord, "-i", str(inp), "-o", str(work / "out"), *SDMD_ARGS.get(name, [])]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"{' '.join(cmd)} failed ({r.returncode}):\n{r.stdout[-2000:]}\n{r.stderr[-2000:]}")
    return tuple(Path(f"{prefix}_{k}.snap") for k in ("pre", "post", "final"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=str(ROOT / "tests" / "golden"))
    ap.add_argument("--work", default=str(REFDIR / "work" / "golden"))
    ap.add_argument("--only", default=None)
    a = ap.parse_args()
    out = Path(a.out)
    out.mkdir(parents=True, exist_ok=True)
    for name in CASES:
        if a.only and a.only != name:
            continue
        snaps = run_case(name, Path(a.work))
        for tag, snap in zip(("pre", "post", "final"), snaps):
            d = read_snapshot(snap)
            if tag != "pre":
                # tables / index maps are identical to the pre snapshot: keep fixtures small
                for k in ("tables", "coll_index", "hb_index"):
                    d[k] = d[k][:0]
                d.pop("gauss", None)
            np.savez_compressed(out / f"{name}_{tag}.npz", **d)
        print(name, "->", [f"{name}_{t}.npz" for t in ("pre", "post", "final")])


if __name__ == "__main__":
    main()
