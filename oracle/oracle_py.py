"""TEST INFRASTRUCTURE — ctypes binding of oracle/liboracle.so (the CPU restatement, sdmd_oracle.c).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this.
"""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

from sdmd_b200.hostview import CHostView, CRemdTemp, CRunStats, HostView, TRACE_DTYPE

HERE = Path(__file__).resolve().parent
_lib = None


class CWallDyn(C.Structure):      # struct sdmd_wall_dyn (include/sdmd_cuda.h): the reference's wallDyn fields + the box edges
    _fields_ = [("mode", C.c_int32), ("sign", C.c_int32), ("touch", C.c_int32), ("pad_", C.c_int32),
                ("size", C.c_double), ("rate", C.c_double), ("step", C.c_double), ("cur_time", C.c_double),
                ("orig_box", C.c_double * 3), ("box", C.c_double * 3)]


def build(force: bool = False) -> Path:
    so = HERE / "liboracle.so"
    src = HERE / "sdmd_oracle.c"
    if force or not so.exists() or so.stat().st_mtime < src.stat().st_mtime:
        subprocess.run(["make", "-C", str(HERE), "port"], check=True, capture_output=True)
    return so


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(str(build()))
        vp = C.c_void_p
        L.oracle_create.restype = vp
        L.oracle_create.argtypes = [C.POINTER(CHostView)]
        L.oracle_destroy.argtypes = [vp]
        L.oracle_rng_feed.argtypes = [vp, C.POINTER(C.c_double), C.c_int64]
        L.oracle_cell_grid.argtypes = [vp, C.POINTER(C.c_int), C.POINTER(C.c_double)]
        for f in ("oracle_pbc_full", "oracle_cell_build", "oracle_predict_all", "oracle_cbt_build"):
            getattr(L, f).argtypes = [vp]
            getattr(L, f).restype = None
        L.oracle_set_celllist.argtypes = [vp, C.POINTER(C.c_int32)]
        L.oracle_celllist.restype = C.POINTER(C.c_int32)
        L.oracle_celllist.argtypes = [vp]
        L.oracle_predict.argtypes = [vp, C.POINTER(C.c_int32)]
        L.oracle_cbt_update.argtypes = [vp, C.POINTER(C.c_int32)]
        L.oracle_cbt_node.restype = C.POINTER(C.c_int32)
        L.oracle_cbt_node.argtypes = [vp]
        L.oracle_cbt_leaf.restype = C.POINTER(C.c_int32)
        L.oracle_cbt_leaf.argtypes = [vp]
        L.oracle_advance.argtypes = [vp, C.c_double]
        L.oracle_run.restype = C.c_int
        L.oracle_run.argtypes = [vp, C.c_double, C.c_int64, C.c_void_p, C.c_int64, C.POINTER(CRunStats)]
        L.oracle_potential_energy.restype = C.c_double
        L.oracle_potential_energy.argtypes = [vp]
        L.oracle_kinetic_energy.restype = C.c_double
        L.oracle_kinetic_energy.argtypes = [vp]
        L.oracle_get_stats.argtypes = [vp, C.POINTER(CRunStats)]
        L.oracle_set_temperature.argtypes = [vp, C.c_double]
        L.oracle_set_wall_dyn.argtypes = [vp, C.POINTER(CWallDyn)]
        L.oracle_set_wall_dyn.restype = None
        L.oracle_get_wall_dyn.argtypes = [vp, C.POINTER(CWallDyn)]
        L.oracle_get_wall_dyn.restype = None
        L.oracle_remd_rearrange.restype = C.c_int
        L.oracle_remd_rearrange.argtypes = [C.POINTER(C.c_double), C.POINTER(CRemdTemp), C.c_int, C.POINTER(C.c_double),
                                            C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_double), C.POINTER(C.c_double)]
        L.oracle_gaussian_libc.restype = C.c_double
        _lib = L
    return _lib


class Oracle:
    """CPU oracle bound to a HostView (mutates its arrays in place)."""

    def __init__(self, hv: HostView):
        self.hv = hv
        self.L = lib()
        self._cs = hv.cstruct()
        self.h = self.L.oracle_create(C.byref(self._cs))
        self.n = hv.natom
        cn = (C.c_int * 3)()
        cs = (C.c_double * 3)()
        self.L.oracle_cell_grid(self.h, cn, cs)
        self.cellnum = np.array(cn[:], dtype=np.int32)
        self.cellsize = np.array(cs[:])
        self.ncell = int(np.prod(self.cellnum))
        if "wall_dyn_i4" in hv.extra and int(hv.extra["wall_dyn_i4"][0]):      # moving wall: wallDyn.* as the reference harness dumped it
            wi, wf = hv.extra["wall_dyn_i4"], hv.extra["wall_dyn_f8"]
            self.set_wall_dyn(int(wi[0]), int(wi[1]), int(wi[2]), *(float(v) for v in wf[:4]), [float(v) for v in wf[4:7]])

    def set_wall_dyn(self, mode, sign, touch, size, rate, step, cur_time, orig_box):
        wd = CWallDyn(mode=mode, sign=sign, touch=touch, size=size, rate=rate, step=step, cur_time=cur_time)
        wd.orig_box[:] = orig_box
        self.L.oracle_set_wall_dyn(self.h, C.byref(wd))

    def wall_dyn(self) -> dict:
        wd = CWallDyn()
        self.L.oracle_get_wall_dyn(self.h, C.byref(wd))
        return {"mode": wd.mode, "sign": wd.sign, "touch": wd.touch, "cur_time": wd.cur_time, "box": list(wd.box)}

    def close(self):
        if self.h:
            self.L.oracle_destroy(self.h)
            self.h = None

    def __del__(self):
        self.close()

    def rng_feed(self, g):
        g = np.ascontiguousarray(g, dtype=np.float64)
        self.L.oracle_rng_feed(self.h, g.ctypes.data_as(C.POINTER(C.c_double)), len(g))

    def pbc_full(self): self.L.oracle_pbc_full(self.h)
    def cell_build(self): self.L.oracle_cell_build(self.h)
    def predict_all(self): self.L.oracle_predict_all(self.h)
    def cbt_build(self): self.L.oracle_cbt_build(self.h)
    def advance(self, dt): self.L.oracle_advance(self.h, float(dt))

    def set_celllist(self, cl):
        cl = np.ascontiguousarray(cl, dtype=np.int32)
        assert len(cl) == self.n + self.ncell + 1
        self.L.oracle_set_celllist(self.h, cl.ctypes.data_as(C.POINTER(C.c_int32)))

    def celllist(self):
        return np.ctypeslib.as_array(self.L.oracle_celllist(self.h), shape=(self.n + self.ncell + 1,)).copy()

    @staticmethod
    def _renew(atoms):
        atoms = list(atoms)
        return np.array([len(atoms), *atoms, 0, 0], dtype=np.int32)

    def predict(self, atoms):
        r = self._renew(atoms)
        self.L.oracle_predict(self.h, r.ctypes.data_as(C.POINTER(C.c_int32)))

    def cbt_update(self, atoms):
        r = self._renew(atoms)
        self.L.oracle_cbt_update(self.h, r.ctypes.data_as(C.POINTER(C.c_int32)))

    def cbt(self):
        node = np.ctypeslib.as_array(self.L.oracle_cbt_node(self.h), shape=(2 * self.n,)).copy()
        leaf = np.ctypeslib.as_array(self.L.oracle_cbt_leaf(self.h), shape=(self.n + 1,)).copy()
        return node, leaf

    def run(self, t_until=1e300, max_iter=1 << 62, trace_cap=0):
        trace = np.zeros(trace_cap, dtype=TRACE_DTYPE)
        st = CRunStats()
        self.L.oracle_run(self.h, float(t_until), int(max_iter), trace.ctypes.data if trace_cap else None, trace_cap, C.byref(st))
        self.hv.sync_scalars_from_c()
        return st.as_dict(), trace[: min(trace_cap, st.iterations)]

    def stats(self):
        st = CRunStats()
        self.L.oracle_get_stats(self.h, C.byref(st))
        return st.as_dict()

    def potential_energy(self): return self.L.oracle_potential_energy(self.h)
    def kinetic_energy(self): return self.L.oracle_kinetic_energy(self.h)
    def set_temperature(self, t): self.L.oracle_set_temperature(self.h, float(t))


def remd_rearrange(epot, nums, temps, u01):
    """Server.c:145-192 on numpy inputs; returns (nums, temps, n_used, attempt, accept)."""
    n = len(epot)
    e = (C.c_double * n)(*[float(x) for x in epot])
    T = (CRemdTemp * n)()
    for i in range(n):
        T[i].num, T[i].T = int(nums[i]), float(temps[i])
    u = np.ascontiguousarray(u01, dtype=np.float64)
    used = C.c_int(0)
    att = (C.c_double * n)()
    acc = (C.c_double * n)()
    rc = lib().oracle_remd_rearrange(e, T, n, u.ctypes.data_as(C.POINTER(C.c_double)), len(u), C.byref(used), att, acc)
    if rc:
        raise RuntimeError(f"oracle_remd_rearrange rc={rc}")
    return [T[i].num for i in range(n)], [T[i].T for i in range(n)], used.value, list(att), list(acc)
