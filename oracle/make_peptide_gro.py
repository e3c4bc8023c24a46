#!/usr/bin/env python
"""TEST INFRASTRUCTURE — synthesise an all-atom .gro of a short peptide for the reference's force field.

BASELINE config 3 is a box of 64 x Abeta16-22 (KLVFFAE); the reference ships no such coordinate file, so
one is built here by distance geometry from the reference's own residue library (Library_Ding/AA/*.txt, read from the
writable copy under oracle/_ref/input): the backbone (N H CA CB C O) of residues 2..8 of the shipped ala10.gro is kept,
and the side-chain atoms beyond CB are placed by minimising the violation of the library's BOND / ANGLE distances, the
[dmin, dmax] ranges of its CONSTRn rows and a soft-core distance to every other atom.  The result is written once to
tests/golden/abeta16_22.gro (our own coordinates, not reference content) so that tests do not depend on the optimiser.

usage: python oracle/make_peptide_gro.py [--seq KLVFFAE] [--out tests/golden/abeta16_22.gro]
"""
from __future__ import annotations

import argparse
import sys
from pathlib import Path

import numpy as np
from scipy.optimize import minimize

ROOT = Path(__file__).resolve().parents[1]
LIB = ROOT / "oracle" / "_ref" / "input"
ONE2THREE = {"A": "ALA", "R": "ARG", "N": "ASN", "D": "ASP", "C": "CYS", "Q": "GLN", "E": "GLU", "G": "GLY", "H": "HIS",
             "I": "ILE", "L": "LEU", "K": "LYS", "M": "MET", "F": "PHE", "S": "SER", "T": "THR", "W": "TRP", "Y": "TYR",
             "V": "VAL"}
BACKBONE = ("N", "H", "CA", "CB", "C", "O")


def read_residue(name):
    atoms, pairs, ranges = [], [], []
    for line in (LIB / "Library_Ding" / "AA" / f"{name}.txt").read_text().splitlines():
        w = line.split()
        if not w:
            continue
        if w[0] == "ATOM":
            atoms.append(w[1])
        elif w[0] in ("BOND", "ANGLE") and not w[2].startswith("+"):
            pairs.append((w[1], w[2], float(w[3]), w[0]))
        elif w[0].startswith("CONSTR") and not w[2].startswith("+"):
            ranges.append((w[1], w[2], float(w[3]), float(w[-1])))
    return atoms, pairs, ranges


def template_backbone(n_res):
    lines = (LIB / "ala10.gro").read_text().splitlines()
    out = []
    for a in lines[2:2 + int(lines[1])]:
        res, name = int(a[0:5]), a[10:15].strip()
        if 2 <= res < 2 + n_res and name in BACKBONE:
            out.append((res - 1, name, 10 * np.array([float(a[20:28]), float(a[28:36]), float(a[36:44])])))
    return out


def build(seq, seed=7):
    rng = np.random.default_rng(seed)
    placed = {}      # (res, atom) -> xyz in A
    for res, name, xyz in template_backbone(len(seq)):
        placed[(res, name)] = xyz
    order = []
    for r, one in enumerate(seq, start=1):
        resname = ONE2THREE[one]
        atoms, pairs, ranges = read_residue(resname)
        _, bb_pairs, _ = read_residue("ALA")
        side = [a for a in atoms if a not in BACKBONE]
        if resname == "GLY":
            side = []
            placed.pop((r, "CB"), None)
        if side:
            idx = {a: k for k, a in enumerate(side)}
            fixed_here = {a: placed[(r, a)] for a in BACKBONE if (r, a) in placed}
            others = np.array([v for (rr, a), v in placed.items() if not (rr == r and a in fixed_here)])
            bonded = {}
            for a, b, d, kind in pairs + bb_pairs:
                bonded.setdefault(a, set()).add(b)
                bonded.setdefault(b, set()).add(a)
            rng_pairs = {(a, b) for a, b, _, _ in ranges} | {(b, a) for a, b, _, _ in ranges}

            def pos(x, a):
                return x[3 * idx[a]:3 * idx[a] + 3] if a in idx else fixed_here[a]

            def cost(x):
                c = 0.0
                for a, b, d, _ in pairs:
                    if a in idx or b in idx:
                        c += 50 * (np.linalg.norm(pos(x, a) - pos(x, b)) - d) ** 2
                for a, b, lo, hi in ranges:
                    if a in idx or b in idx:
                        dist = np.linalg.norm(pos(x, a) - pos(x, b))
                        c += 20 * max(0.0, lo + 0.3 - dist) ** 2 + 20 * max(0.0, dist - hi + 0.3) ** 2
                names = list(idx) + list(fixed_here)
                for i, a in enumerate(idx):           # soft core inside the residue
                    for b in names:
                        if b == a or b in bonded.get(a, ()) or (a, b) in rng_pairs or (b in idx and idx[b] < idx[a]):
                            continue
                        if any(b in bonded.get(m, ()) for m in bonded.get(a, ())):
                            continue                  # 1-3 pairs are held by ANGLE rows
                        dist = np.linalg.norm(pos(x, a) - pos(x, b))
                        c += 10 * max(0.0, 3.3 - dist) ** 2
                if len(others):                       # soft core against the rest of the chain
                    for a in idx:
                        dd = np.linalg.norm(others - pos(x, a), axis=1)
                        c += 10 * np.sum(np.maximum(0.0, 3.6 - dd) ** 2)
                return c

            out_dir = fixed_here["CB"] - fixed_here["CA"]
            best = None
            for _ in range(24):
                x0 = np.concatenate([fixed_here["CB"] + out_dir * (1 + 0.8 * k) + rng.normal(0, 0.8, 3) for k in range(len(side))])
                res_ = minimize(cost, x0, method="L-BFGS-B", options={"maxiter": 4000})
                if best is None or res_.fun < best.fun:
                    best = res_
            print(f"residue {r} {resname}: side chain of {len(side)} atoms, residual {best.fun:.2e}", file=sys.stderr)
            for a in side:
                placed[(r, a)] = best.x[3 * idx[a]:3 * idx[a] + 3]
        for a in [b for b in ("N", "H", "CA", "CB") if (r, b) in placed] + side + ["C", "O"]:
            order.append((r, resname, a, placed[(r, a)]))
    return order


def write_gro(path, atoms, box_nm=15.0):
    xyz = np.array([a[3] for a in atoms]) / 10
    xyz += box_nm / 2 - xyz.mean(axis=0)
    with open(path, "w") as f:
        f.write("synthetic peptide (oracle/make_peptide_gro.py)\n%5d\n" % len(atoms))
        for k, ((r, resname, a, _), p) in enumerate(zip(atoms, xyz), start=1):
            f.write("%5d%-5s%5s%5d%8.3f%8.3f%8.3f\n" % (r, resname, a, k, p[0], p[1], p[2]))
        f.write("%10.5f%10.5f%10.5f\n" % (box_nm, box_nm, box_nm))


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--seq", default="KLVFFAE")
    ap.add_argument("--out", default=str(ROOT / "tests" / "golden" / "abeta16_22.gro"))
    a = ap.parse_args()
    write_gro(a.out, build(a.seq))
    print("wrote", a.out)
