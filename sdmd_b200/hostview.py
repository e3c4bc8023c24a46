"""Host-side mirror of ``struct sdmd_host_view`` (include/sdmd_cuda.h) for Python callers.

A :class:`HostView` owns numpy arrays laid out exactly as the C-ABI expects (1-based atoms, natom+1
entries per per-atom array) and exposes a ctypes struct pointing into them.  It can be filled from a
snapshot written by the reference glue (``sdmd_b200/host/sdmd_refglue.c``; format: ``SDMDSNP1`` tagged
arrays) or from a synthetic generator (``sdmd_b200/synth.py``).
"""
from __future__ import annotations

import ctypes as C
import struct
from pathlib import Path

import numpy as np

ABI_VERSION = 8
MAX_CHARGE = 8
MAX_OBST = 4
NT = 33
NHBCLASS = 12
MAX_STEPS = 16
INFINIT = 1e8
INVALID = -111

STEP_TABLE_DTYPE = np.dtype(
    [("dmin2", "<f8"), ("dmax2", "<f8"), ("nstep", "<i4"), ("pad_", "<i4"),
     ("d2", "<f8", (MAX_STEPS,)), ("e", "<f8", (MAX_STEPS,)), ("acc", "<f8", (MAX_STEPS,))]
)
assert STEP_TABLE_DTYPE.itemsize == 408

TRACE_DTYPE = np.dtype(
    [("frame", "<i8"), ("target", "<i4"), ("partner", "<i4"), ("type", "<i4"), ("sub", "<i4"),
     ("executed", "<i4"), ("pad_", "<i4"), ("dt", "<f8")]
)
assert TRACE_DTYPE.itemsize == 40

_P = C.POINTER


class CHostView(C.Structure):
    _fields_ = [
        ("abi_version", C.c_int32), ("natom", C.c_int32),
        ("box", C.c_double * 3), ("cutoffr", C.c_double), ("thermoF", C.c_double),
        ("target_temperature", C.c_double), ("currenttime", C.c_double), ("timestep", C.c_double),
        ("thermostat_on", C.c_int32), ("method_ding", C.c_int32), ("hydrogen_bond", C.c_int32), ("gel_mode", C.c_int32),
        ("frame", C.c_int64),
        ("x", _P(C.c_double)), ("y", _P(C.c_double)), ("z", _P(C.c_double)),
        ("vx", _P(C.c_double)), ("vy", _P(C.c_double)), ("vz", _P(C.c_double)),
        ("mass", _P(C.c_double)), ("type", _P(C.c_int32)), ("aa_type", _P(C.c_int32)),
        ("protein_num", _P(C.c_int32)), ("aa_num", _P(C.c_int32)), ("is_cons", _P(C.c_uint8)),
        ("hb_bond", _P(C.c_int32)), ("hb_neighbor", _P(C.c_int32)), ("hb_role", _P(C.c_uint8)), ("hb_itype", _P(C.c_uint8)),
        ("hbn_partner", _P(C.c_int32)), ("hbn_partner_num", _P(C.c_int32)), ("hbn_status", _P(C.c_int32)),
        ("ev_time", _P(C.c_double)), ("ev_potential", _P(C.c_double)), ("ev_partner", _P(C.c_int32)),
        ("ev_counter", _P(C.c_int64)), ("ev_pcounter", _P(C.c_int64)), ("ev_type", _P(C.c_uint8)), ("ev_sub", _P(C.c_uint8)),
        ("cell_idx", _P(C.c_int32)),
        ("bond_off", _P(C.c_int32)), ("bond_j", _P(C.c_int32)), ("bond_dmin2", _P(C.c_double)), ("bond_dmax2", _P(C.c_double)),
        ("con_off", _P(C.c_int32)), ("con_j", _P(C.c_int32)), ("con_table", _P(C.c_int32)),
        ("cm_off", _P(C.c_int32)), ("cm_j", _P(C.c_int32)), ("cm_flags", _P(C.c_int32)),
        ("ntable", C.c_int32), ("tables", C.c_void_p),
        ("coll_index", _P(C.c_int32)), ("hb_index", _P(C.c_int32)),
        ("hb_barrier", C.c_double * 3),
        ("wall_kind", C.c_int32), ("wall_type", C.c_int32), ("wall_pos", C.c_double * 3),
        ("n_obst", C.c_int32), ("n_sphobst", C.c_int32),
        ("obst_pos", C.c_double * (3 * MAX_OBST)), ("obst_type", C.c_int32 * MAX_OBST),
        ("sphobst_pos", C.c_double * (3 * MAX_OBST)), ("sphobst_radius", C.c_double * MAX_OBST),
        ("n_tunnel", C.c_int32), ("tunnel_type", C.c_int32),
        ("tunnel_pos", C.c_double * (4 * MAX_OBST)), ("tunnel_diameter", C.c_double * MAX_OBST), ("obst_hole", C.c_int32 * MAX_OBST),
        ("flow_mark", C.c_int32), ("pad_flow_", C.c_int32), ("flow_v", C.c_double * 3),
        ("flow_a", _P(C.c_double)), ("flow_time_rec", _P(C.c_double)), ("protein_mass", _P(C.c_double)),
        ("charge", _P(C.c_double)), ("n_charge", C.c_int32), ("charge_com", C.c_int32),
        ("charge_pos", C.c_double * (3 * MAX_CHARGE)), ("charge_vel", C.c_double * (3 * MAX_CHARGE)),
        ("charge_potgrad", C.c_double * MAX_CHARGE), ("charge_gap", C.c_double * MAX_CHARGE),
        ("hb_name", _P(C.c_uint8)), ("hb_num_formed", C.c_int32), ("alpha_hb_formed", C.c_int32),
        ("wall_dyn", C.c_int32), ("wall_dyn_touch", C.c_int32), ("wall_dyn_step", C.c_double), ("wall_dyn_orig_box", C.c_double * 3),
    ]


class CRunStats(C.Structure):
    _fields_ = [
        ("frame", C.c_int64), ("iterations", C.c_int64), ("count_recal", C.c_int64),
        ("collision", C.c_int64), ("bond", C.c_int64), ("hb", C.c_int64), ("hb_neighbor", C.c_int64),
        ("thermostat", C.c_int64), ("pcc", C.c_int64), ("warnings", C.c_int64),
        ("pair_predictions", C.c_int64), ("gaussians_used", C.c_int64),
        ("currenttime", C.c_double), ("stop_reason", C.c_int32), ("error", C.c_int32),
        ("wall", C.c_int64), ("error_prints", C.c_int64),
        ("hb_num_formed", C.c_int64), ("alpha_hb_formed", C.c_int64),
    ]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


class CRemdTemp(C.Structure):
    _fields_ = [("num", C.c_int32), ("pad_", C.c_int32), ("T", C.c_double)]


# (name, dtype, entries-per-atom or None for non per-atom)
PER_ATOM = [
    ("x", "f8", 1), ("y", "f8", 1), ("z", "f8", 1), ("vx", "f8", 1), ("vy", "f8", 1), ("vz", "f8", 1),
    ("mass", "f8", 1), ("type", "i4", 1), ("aa_type", "i4", 1), ("protein_num", "i4", 1), ("aa_num", "i4", 1),
    ("is_cons", "u1", 1), ("hb_bond", "i4", 1), ("hb_neighbor", "i4", 1), ("hb_role", "u1", 1), ("hb_itype", "u1", 1),
    ("hbn_partner", "i4", 5), ("hbn_partner_num", "i4", 1), ("hbn_status", "i4", 1),
    ("ev_time", "f8", 1), ("ev_potential", "f8", 1), ("ev_partner", "i4", 1), ("ev_counter", "i8", 1),
    ("ev_pcounter", "i8", 1), ("ev_type", "u1", 1), ("ev_sub", "u1", 1), ("cell_idx", "i4", 3),
    ("flow_a", "f8", 3), ("flow_time_rec", "f8", 1), ("protein_mass", "f8", 1), ("charge", "f8", 1),
    ("hb_name", "u1", 1),
]
MUTABLE = ["x", "y", "z", "vx", "vy", "vz", "type", "hb_bond", "hb_itype", "hbn_partner", "hbn_partner_num",
           "hbn_status", "ev_time", "ev_potential", "ev_partner", "ev_counter", "ev_pcounter", "ev_type", "ev_sub",
           "cell_idx", "flow_time_rec"]
OTHER = [
    ("bond_off", "i4"), ("bond_j", "i4"), ("bond_dmin2", "f8"), ("bond_dmax2", "f8"),
    ("con_off", "i4"), ("con_j", "i4"), ("con_table", "i4"),
    ("cm_off", "i4"), ("cm_j", "i4"), ("cm_flags", "i4"),
    ("coll_index", "i4"), ("hb_index", "i4"),
]
_NP = {"f8": np.float64, "i4": np.int32, "i8": np.int64, "u1": np.uint8}
_CT = {"f8": C.c_double, "i4": C.c_int32, "i8": C.c_int64, "u1": C.c_uint8}


def read_snapshot(path) -> dict:
    """Parse a ``SDMDSNP1`` file into {name: ndarray}."""
    data = Path(path).read_bytes()
    if data[:8] != b"SDMDSNP1":
        raise ValueError(f"{path}: not a snapshot")
    out, pos = {}, 8
    while True:
        name = data[pos:pos + 24].split(b"\0")[0].decode()
        dt = data[pos + 24:pos + 32].split(b"\0")[0].decode()
        (count,) = struct.unpack_from("<Q", data, pos + 32)
        pos += 40
        if name == "END":
            break
        nbytes = np.dtype(_NP[dt]).itemsize * count
        out[name] = np.frombuffer(data, dtype=_NP[dt], count=count, offset=pos).copy()
        pos += (nbytes + 7) // 8 * 8
    return out


class HostView:
    """numpy-backed ``struct sdmd_host_view``."""

    def __init__(self, natom: int):
        self.natom = int(natom)
        n1 = self.natom + 1
        self.box = np.zeros(3)
        self.cutoffr = 9.0
        self.thermoF = 6.0
        self.target_temperature = 0.6
        self.currenttime = 0.0
        self.timestep = 0.0
        self.thermostat_on = 1
        self.method_ding = 1
        self.hydrogen_bond = 1
        self.gel_mode = 0
        self.frame = 0
        self.hb_barrier = np.zeros(3)
        self.wall_kind = 0          # enum sdmd_wall_kind: 0 none, 1 parallel plates, 2 cylinder, 3 sphere
        self.wall_type = 0
        self.wall_pos = np.zeros(3)
        # flat obstruction walls (position per dimension, negative = none; atom type) and hard spherical obstacles
        self.obst_pos = np.zeros((0, 3))
        self.obst_type = np.zeros(0, dtype=np.int32)
        self.sphobst_pos = np.zeros((0, 3))
        self.sphobst_radius = np.zeros(0)
        # tunnels through obstruction walls: (x0, x1, y, z) per tunnel, diameter, wall atom type; holed obstruction walls
        self.tunnel_pos = np.zeros((0, 4))
        self.tunnel_diameter = np.zeros(0)
        self.tunnel_type = 0
        self.obst_hole = np.zeros(0, dtype=np.int32)
        self.flow_mark = 0          # 0 none, 1 constant drift velocity, 2 constant force
        self.flow_v = np.zeros(3)
        self.charge_com = 0
        self.wall_dyn = 0           # moving wall (WallDyn): on / wallDyn.touch / wallDyn.step / wallDyn.origBoxDim
        self.wall_dyn_touch = 0
        self.wall_dyn_step = 0.0
        self.wall_dyn_orig_box = np.zeros(3)
        self.hb_num_formed = 0      # HBnumformed / alphaHBformed at the start of the run
        self.alpha_hb_formed = 0
        self.charge_src = np.zeros((0, 8))   # per charge source: position (3), velocity (3), potential gradient, gap
        for name, dt, k in PER_ATOM:
            setattr(self, name, np.zeros(n1 * k, dtype=_NP[dt]))
        for name in ("bond_off", "con_off", "cm_off"):
            setattr(self, name, np.zeros(n1 + 1, dtype=np.int32))
        for name, dt in OTHER:
            if not hasattr(self, name):
                setattr(self, name, np.zeros(1, dtype=_NP[dt]))
        self.coll_index = np.full(NT * NT, -1, dtype=np.int32)
        self.hb_index = np.full(NHBCLASS * NT * NT, -1, dtype=np.int32)
        self.tables = np.zeros(0, dtype=STEP_TABLE_DTYPE)
        self.extra = {}
        self._c = None

    # ------------------------------------------------------------------
    @classmethod
    def from_arrays(cls, d: dict) -> "HostView":
        si, sf = d["scalars_i4"], d["scalars_f8"]
        hv = cls(int(si[1]))
        hv.thermostat_on, hv.method_ding, hv.hydrogen_bond, hv.gel_mode = (int(v) for v in si[2:6])
        hv.box = sf[0:3].copy()
        hv.cutoffr, hv.thermoF, hv.target_temperature, hv.currenttime, hv.timestep = (float(v) for v in sf[3:8])
        hv.hb_barrier = sf[8:11].copy()
        hv.frame = int(d["scalars_i8"][0])
        if "wall_i4" in d:          # absent in snapshots written before the wall fields existed = no wall
            hv.wall_kind, hv.wall_type = (int(v) for v in d["wall_i4"][:2])
            hv.wall_pos = np.asarray(d["wall_f8"], dtype=np.float64)[:3].copy()
        if "obst_pos" in d:
            hv.obst_pos = np.asarray(d["obst_pos"], dtype=np.float64).reshape(-1, 3).copy()
            hv.obst_type = np.asarray(d["obst_type"], dtype=np.int32).copy()
            hv.sphobst_pos = np.asarray(d["sphobst_pos"], dtype=np.float64).reshape(-1, 3).copy()
            hv.sphobst_radius = np.asarray(d["sphobst_radius"], dtype=np.float64).copy()
        if "tunnel_pos" in d:
            hv.tunnel_pos = np.asarray(d["tunnel_pos"], dtype=np.float64).reshape(-1, 4).copy()
            hv.tunnel_diameter = np.asarray(d["tunnel_diameter"], dtype=np.float64).copy()
            hv.tunnel_type = int(d["tunnel_i4"][0])
            hv.obst_hole = np.asarray(d["obst_hole"], dtype=np.int32).copy()
        for name, dt, _ in PER_ATOM:
            if name in d:               # arrays added later (flow_*) are absent in older snapshots: keep the zeros
                setattr(hv, name, np.ascontiguousarray(d[name], dtype=_NP[dt]).copy())
        if "hb_i4" in d:
            hv.hb_num_formed, hv.alpha_hb_formed = (int(v) for v in d["hb_i4"][:2])
        if "wall_dyn_i4" in d:      # (written by oracle/ref_harness.c; the arrays also stay in `extra` for Oracle.set_wall_dyn)
            wi, wf = d["wall_dyn_i4"], d["wall_dyn_f8"]
            hv.wall_dyn = int(wi[0] != 0)
            hv.wall_dyn_touch = int(wi[2])
            hv.wall_dyn_step = float(wf[2])
            hv.wall_dyn_orig_box = np.asarray(wf[4:7], dtype=np.float64).copy()
        if "flow_f8" in d:
            hv.flow_mark = int(d["flow_i4"][0])
            hv.flow_v = np.asarray(d["flow_f8"], dtype=np.float64)[:3].copy()
        if "charge_src" in d:
            hv.charge_src = np.asarray(d["charge_src"], dtype=np.float64).reshape(-1, 8).copy()
            hv.charge_com = int(d["flow_i4"][1]) if len(d["flow_i4"]) > 1 else 0
        for name, dt in OTHER:
            arr = np.ascontiguousarray(d[name], dtype=_NP[dt]).copy()
            setattr(hv, name, arr if arr.size else np.zeros(1, dtype=_NP[dt]))
        hv.tables = np.frombuffer(d["tables"].tobytes(), dtype=STEP_TABLE_DTYPE).copy()
        known = {n for n, _, _ in PER_ATOM} | {n for n, _ in OTHER} | {"scalars_i4", "scalars_f8", "scalars_i8", "tables", "wall_i4", "wall_f8",
                                                                           "obst_pos", "obst_type", "sphobst_pos", "sphobst_radius",
                                                                           "tunnel_pos", "tunnel_diameter", "tunnel_i4", "obst_hole", "flow_i4", "flow_f8", "charge_src", "hb_i4"}
        hv.extra = {k: v for k, v in d.items() if k not in known}
        return hv

    @classmethod
    def from_snapshot(cls, path) -> "HostView":
        p = Path(path)
        if p.suffix == ".npz":
            with np.load(p) as z:
                return cls.from_arrays({k: z[k] for k in z.files})
        return cls.from_arrays(read_snapshot(p))

    def to_arrays(self) -> dict:
        d = {
            "scalars_i4": np.array([ABI_VERSION, self.natom, self.thermostat_on, self.method_ding, self.hydrogen_bond,
                                    self.gel_mode, len(self.tables), 0], dtype=np.int32),
            "scalars_f8": np.array([*self.box, self.cutoffr, self.thermoF, self.target_temperature, self.currenttime,
                                    self.timestep, *self.hb_barrier, 0.0]),
            "scalars_i8": np.array([self.frame], dtype=np.int64),
            "wall_i4": np.array([self.wall_kind, self.wall_type], dtype=np.int32),
            "wall_f8": np.array(self.wall_pos, dtype=np.float64),
            "obst_pos": np.asarray(self.obst_pos, dtype=np.float64).reshape(-1),
            "obst_type": np.asarray(self.obst_type, dtype=np.int32),
            "sphobst_pos": np.asarray(self.sphobst_pos, dtype=np.float64).reshape(-1),
            "sphobst_radius": np.asarray(self.sphobst_radius, dtype=np.float64),
            "tunnel_pos": np.asarray(self.tunnel_pos, dtype=np.float64).reshape(-1),
            "tunnel_diameter": np.asarray(self.tunnel_diameter, dtype=np.float64),
            "tunnel_i4": np.array([self.tunnel_type], dtype=np.int32),
            "obst_hole": np.asarray(self.obst_hole, dtype=np.int32),
            "flow_i4": np.array([self.flow_mark, self.charge_com], dtype=np.int32),
            "hb_i4": np.array([self.hb_num_formed, self.alpha_hb_formed], dtype=np.int32),
            "charge_src": np.asarray(self.charge_src, dtype=np.float64).reshape(-1),
            "flow_f8": np.asarray(self.flow_v, dtype=np.float64),
            "tables": np.frombuffer(self.tables.tobytes(), dtype=np.uint8),
        }
        for name, _, _ in PER_ATOM:
            d[name] = getattr(self, name)
        for name, _ in OTHER:
            d[name] = getattr(self, name)
        d.update(self.extra)
        return d

    def copy(self) -> "HostView":
        return HostView.from_arrays({k: np.array(v, copy=True) for k, v in self.to_arrays().items()})

    def reset_slots(self):
        """Event slots as the reference leaves them after calloc (Initialization.c:1012-1017)."""
        for name in ("ev_time", "ev_potential", "ev_partner", "ev_counter", "ev_pcounter", "ev_type", "ev_sub"):
            getattr(self, name)[:] = 0

    # ------------------------------------------------------------------
    def cstruct(self) -> CHostView:
        """(Re)build the ctypes struct; pointers alias the numpy arrays (keep ``self`` alive)."""
        c = CHostView()
        c.abi_version = ABI_VERSION
        c.natom = self.natom
        c.box[:] = [float(v) for v in self.box]
        c.cutoffr, c.thermoF = float(self.cutoffr), float(self.thermoF)
        c.target_temperature, c.currenttime, c.timestep = float(self.target_temperature), float(self.currenttime), float(self.timestep)
        c.thermostat_on, c.method_ding = int(self.thermostat_on), int(self.method_ding)
        c.hydrogen_bond, c.gel_mode = int(self.hydrogen_bond), int(self.gel_mode)
        c.frame = int(self.frame)
        c.hb_barrier[:] = [float(v) for v in self.hb_barrier]
        c.wall_kind, c.wall_type = int(self.wall_kind), int(self.wall_type)
        c.wall_pos[:] = [float(v) for v in self.wall_pos]
        c.n_obst, c.n_sphobst = len(self.obst_type), len(self.sphobst_radius)
        assert c.n_obst <= MAX_OBST and c.n_sphobst <= MAX_OBST
        for k, v in enumerate(np.asarray(self.obst_pos, dtype=np.float64).reshape(-1)):
            c.obst_pos[k] = float(v)
        for k, v in enumerate(self.obst_type):
            c.obst_type[k] = int(v)
        for k, v in enumerate(np.asarray(self.sphobst_pos, dtype=np.float64).reshape(-1)):
            c.sphobst_pos[k] = float(v)
        for k, v in enumerate(self.sphobst_radius):
            c.sphobst_radius[k] = float(v)
        c.flow_mark = int(self.flow_mark)
        c.hb_num_formed, c.alpha_hb_formed = int(self.hb_num_formed), int(self.alpha_hb_formed)
        c.wall_dyn, c.wall_dyn_touch, c.wall_dyn_step = int(self.wall_dyn), int(self.wall_dyn_touch), float(self.wall_dyn_step)
        c.wall_dyn_orig_box[:] = [float(v) for v in self.wall_dyn_orig_box]
        cs = np.asarray(self.charge_src, dtype=np.float64).reshape(-1, 8)
        assert len(cs) <= MAX_CHARGE
        c.n_charge, c.charge_com = len(cs), int(self.charge_com)
        for n, row in enumerate(cs):
            for k in range(3):
                c.charge_pos[3 * n + k] = float(row[k]); c.charge_vel[3 * n + k] = float(row[3 + k])
            c.charge_potgrad[n] = float(row[6]); c.charge_gap[n] = float(row[7])
        c.flow_v[:] = [float(v) for v in self.flow_v]
        c.n_tunnel, c.tunnel_type = len(self.tunnel_diameter), int(self.tunnel_type)
        assert c.n_tunnel <= MAX_OBST
        for k, v in enumerate(np.asarray(self.tunnel_pos, dtype=np.float64).reshape(-1)):
            c.tunnel_pos[k] = float(v)
        for k, v in enumerate(self.tunnel_diameter):
            c.tunnel_diameter[k] = float(v)
        for k, v in enumerate(self.obst_hole):
            c.obst_hole[k] = int(v)
        for name, dt, _ in PER_ATOM:
            arr = getattr(self, name)
            assert arr.flags["C_CONTIGUOUS"] and arr.dtype == _NP[dt], name
            setattr(c, name, arr.ctypes.data_as(_P(_CT[dt])))
        for name, dt in OTHER:
            arr = getattr(self, name)
            assert arr.flags["C_CONTIGUOUS"] and arr.dtype == _NP[dt], name
            setattr(c, name, arr.ctypes.data_as(_P(_CT[dt])))
        c.ntable = len(self.tables)
        c.tables = self.tables.ctypes.data if len(self.tables) else None
        self._c = c
        return c

    def sync_scalars_from_c(self):
        if self._c is not None:
            self.box = np.array(self._c.box[:], dtype=np.float64)      # (a moving wall changes the box edges)
            self.currenttime = self._c.currenttime
            self.frame = self._c.frame
            self.target_temperature = self._c.target_temperature

    def slots(self) -> dict:
        return {k: getattr(self, k).copy() for k in
                ("ev_time", "ev_potential", "ev_partner", "ev_pcounter", "ev_counter", "ev_type", "ev_sub")}
