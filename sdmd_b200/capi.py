"""ctypes binding of libsdmd_cuda.so — the call a Python user makes.

The class :class:`Replica` mirrors the reference's driver-level interface for the hot path with the same
names and argument meaning (``PBC("full")`` → :meth:`pbc_full`, ``LinkList("full")`` → :meth:`cell_build`,
``Predict`` → :meth:`predict` / :meth:`predict_all`, ``CreateCBT``/``UpdateCBT`` → :meth:`cbt_build` /
:meth:`cbt_update`, ``AssignJob``, ``UpdateData``+``TimeForward`` → :meth:`advance`, ``HazardCheck``,
``DoEvent``, ``CalPotenE``, ``SingleThreadRun`` → :meth:`run`).  Every call goes through the C-ABI of
``include/sdmd_cuda.h``; there is no CPU path — if the library or a CUDA device is missing this raises.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path

import numpy as np

from .hostview import CHostView, CRemdTemp, CRunStats, HostView, TRACE_DTYPE

PKG = Path(__file__).resolve().parent
ROOT = PKG.parent
LIB_PATH = PKG / "libsdmd_cuda.so"
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-fmad=false", "-lineinfo", "-O3", "-std=c++17",
              "-shared", "-Xcompiler", "-fPIC"]

# every symbol include/sdmd_cuda.h declares (checked by tests/test_abi.py)
ABI_SYMBOLS = [
    "sdmd_cuda_create", "sdmd_cuda_destroy", "sdmd_cuda_last_error", "sdmd_cuda_device_count", "sdmd_cuda_cell_grid",
    "sdmd_cuda_upload_coords", "sdmd_cuda_upload_celllist", "sdmd_cuda_download", "sdmd_cuda_download_slots",
    "sdmd_cuda_download_celllist", "sdmd_cuda_download_cbt",
    "sdmd_cuda_rng_feed", "sdmd_cuda_rng_fill_libc", "sdmd_cuda_rng_available", "sdmd_cuda_rng_seek", "sdmd_cuda_rng_reset", "sdmd_host_gaussian_libc",
    "sdmd_cuda_timer_begin", "sdmd_cuda_timer_end", "sdmd_cuda_synchronize", "sdmd_cuda_host_register", "sdmd_cuda_host_unregister",
    "sdmd_cuda_pbc_full", "sdmd_cuda_cell_build", "sdmd_cuda_predict_all", "sdmd_cuda_predict", "sdmd_cuda_cbt_build", "sdmd_cuda_prepare", "sdmd_cuda_prepare_resident",
    "sdmd_cuda_cbt_update", "sdmd_cuda_wall_dyn_step", "sdmd_cuda_wall_dyn_state", "sdmd_cuda_assign_job", "sdmd_cuda_advance", "sdmd_cuda_add_time", "sdmd_cuda_hazard_check",
    "sdmd_cuda_cell_move", "sdmd_cuda_do_event", "sdmd_cuda_potential_energy", "sdmd_cuda_kinetic_energy",
    "sdmd_cuda_run", "sdmd_cuda_run_many", "sdmd_cuda_potential_energy_many", "sdmd_cuda_get_stats", "sdmd_cuda_set_temperature",
    "sdmd_remd_rearrange", "sdmd_remd_nccl_unique_id", "sdmd_remd_nccl_init", "sdmd_remd_seed_server", "sdmd_remd_exchange", "sdmd_remd_exchange_energy", "sdmd_remd_finish", "sdmd_host_glibc_rand",
    "sdmd_cuda_kernel_ms", "sdmd_cuda_kernel_stats", "sdmd_cuda_reset_kernel_stats", "sdmd_cuda_set_timing", "sdmd_cuda_set_loop_mode",
    "sdmd_cuda_launch_count", "sdmd_cuda_stream", "sdmd_cuda_phase_cycles",
]


class SdmdError(RuntimeError):
    pass


CSRC = PKG / "csrc"
SOURCES = ["sdmd_cuda.cu", "sdmd_k2_tile.cu", "sdmd_k2_rows.cu"]            # one object per translation unit, rebuilt only when it (or a header) changed
HEADERS = [CSRC / "sdmd_dev.cuh", CSRC / "sdmd_dev_body.cuh", CSRC / "sdmd_host.h", ROOT / "include" / "sdmd_cuda.h"]
OBJ_DIR = ROOT / "build" / "obj"
PTXAS_LOG = ROOT / "profiles" / "ptxas"                  # registers / spills / shared memory per kernel, tracked (one file per TU)


def _compile_tu(name: str, extra_flags=()) -> tuple[str, int, str]:
    obj = OBJ_DIR / (name + ".o")
    cmd = ["nvcc", *[f for f in NVCC_FLAGS if f != "-shared"], "-Xptxas", "-v", *extra_flags, f"-I{ROOT / 'include'}", f"-I{CSRC}",
           "-c", str(CSRC / name), "-o", str(obj)]
    r = subprocess.run(cmd, capture_output=True, text=True)
    return " ".join(cmd), r.returncode, r.stdout + r.stderr


def build(force: bool = False, verbose: bool = False, extra_flags=(), out: Path | None = None) -> Path:
    """Compile libsdmd_cuda.so in-tree for sm_100a (nvcc cross-compiles without a GPU).  `extra_flags` / `out` build a named
    variant (scripts/k2_variants.py) without touching the shipped library or the object cache."""
    from concurrent.futures import ThreadPoolExecutor
    lib_path = out or LIB_PATH
    variant = bool(extra_flags) or out is not None
    OBJ_DIR.mkdir(parents=True, exist_ok=True)
    newest_hdr = max(h.stat().st_mtime for h in HEADERS if h.exists())
    stale = []
    for name in SOURCES:
        obj = OBJ_DIR / (name + ".o")
        if force or variant or not obj.exists() or obj.stat().st_mtime < max(newest_hdr, (CSRC / name).stat().st_mtime):
            stale.append(name)
    if not stale and lib_path.exists() and all(lib_path.stat().st_mtime >= (OBJ_DIR / (n + ".o")).stat().st_mtime for n in SOURCES):
        return lib_path
    with ThreadPoolExecutor(len(SOURCES)) as ex:
        results = list(ex.map(lambda n: (n, *_compile_tu(n, extra_flags)), stale))
    for name, cmd, rc, log in results:
        if verbose or rc:
            print(cmd)
            print(log)
        if rc:
            raise SdmdError(f"nvcc failed building {name}")
        if not variant:
            PTXAS_LOG.mkdir(parents=True, exist_ok=True)
            keep = [ln for ln in log.splitlines() if "ptxas info" in ln or "Compiling entry" in ln or "Function properties" in ln or "bytes stack frame" in ln]
            (PTXAS_LOG / (name + ".txt")).write_text("\n".join(keep) + "\n")
    cmd = ["nvcc", "-shared", "-gencode", "arch=compute_100a,code=sm_100a", *[str(OBJ_DIR / (n + ".o")) for n in SOURCES], "-o", str(lib_path), "-ldl"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or r.returncode:
        print(" ".join(cmd))
        print(r.stdout, r.stderr)
    if r.returncode:
        raise SdmdError("nvcc failed linking libsdmd_cuda.so")
    if variant:                                           # the cache now holds variant objects: force a rebuild next time
        for n in SOURCES:
            (OBJ_DIR / (n + ".o")).unlink(missing_ok=True)
    return lib_path


_lib = None


def lib():
    """Load the shared library; raises (never falls back) when it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise SdmdError(f"{LIB_PATH} not built: run `python -c 'import __graft_entry__ as g; g.build()'` (no CPU fallback exists)")
    L = C.CDLL(str(LIB_PATH), mode=C.RTLD_GLOBAL)
    vp, i32p, dp = C.c_void_p, C.POINTER(C.c_int32), C.POINTER(C.c_double)
    L.sdmd_cuda_last_error.restype = C.c_char_p
    L.sdmd_cuda_create.argtypes = [C.POINTER(CHostView), C.c_int, C.POINTER(vp)]
    L.sdmd_cuda_destroy.argtypes = [vp]
    L.sdmd_cuda_destroy.restype = None
    L.sdmd_cuda_cell_grid.argtypes = [vp, C.POINTER(C.c_int), dp]
    L.sdmd_cuda_upload_coords.argtypes = [vp] + [dp] * 6
    L.sdmd_cuda_upload_celllist.argtypes = [vp, i32p]
    L.sdmd_cuda_prepare.argtypes = [vp] + [dp] * 7 + [C.c_int64]
    L.sdmd_cuda_prepare_resident.argtypes = [vp]
    L.sdmd_cuda_download.argtypes = [vp, C.POINTER(CHostView)]
    L.sdmd_cuda_download_slots.argtypes = [vp, dp, dp, i32p, C.POINTER(C.c_uint8), C.POINTER(C.c_uint8)]
    L.sdmd_cuda_download_celllist.argtypes = [vp, i32p]
    L.sdmd_cuda_download_cbt.argtypes = [vp, i32p, i32p]
    L.sdmd_cuda_rng_feed.argtypes = [vp, dp, C.c_int64]
    L.sdmd_cuda_rng_fill_libc.argtypes = [vp, C.c_int64]
    L.sdmd_cuda_rng_available.argtypes = [vp]
    L.sdmd_cuda_rng_available.restype = C.c_int64
    L.sdmd_host_gaussian_libc.restype = C.c_double
    L.sdmd_cuda_rng_seek.argtypes = [vp, C.c_int64]
    L.sdmd_cuda_rng_reset.argtypes = [vp]
    L.sdmd_cuda_timer_begin.argtypes = [vp]
    L.sdmd_cuda_timer_end.argtypes = [vp, dp]
    L.sdmd_cuda_synchronize.argtypes = [vp]
    L.sdmd_cuda_host_register.argtypes = [vp, C.c_size_t]
    L.sdmd_cuda_host_unregister.argtypes = [vp]
    for f in ("pbc_full", "cell_build", "predict_all", "cbt_build", "do_event"):
        getattr(L, "sdmd_cuda_" + f).argtypes = [vp]
    L.sdmd_cuda_predict.argtypes = [vp, i32p]
    L.sdmd_cuda_cbt_update.argtypes = [vp, i32p]
    L.sdmd_cuda_assign_job.argtypes = [vp, i32p, i32p, dp, i32p]
    L.sdmd_cuda_advance.argtypes = [vp, C.c_double]
    L.sdmd_cuda_add_time.argtypes = [vp, C.c_double]
    L.sdmd_cuda_hazard_check.argtypes = [vp, C.c_int32, C.c_int32, i32p]
    L.sdmd_cuda_cell_move.argtypes = [vp, C.c_int32]
    L.sdmd_cuda_potential_energy.argtypes = [vp, dp]
    L.sdmd_cuda_kinetic_energy.argtypes = [vp, dp]
    L.sdmd_cuda_run.argtypes = [vp, C.c_double, C.c_int64, C.c_void_p, C.c_int64, C.POINTER(CRunStats)]
    L.sdmd_cuda_get_stats.argtypes = [vp, C.POINTER(CRunStats)]
    L.sdmd_cuda_run_many.argtypes = [C.POINTER(vp), C.c_int, dp, C.c_int64, C.POINTER(CRunStats)]
    L.sdmd_cuda_potential_energy_many.argtypes = [C.POINTER(vp), C.c_int, dp]
    L.sdmd_cuda_set_temperature.argtypes = [vp, C.c_double]
    L.sdmd_remd_rearrange.argtypes = [dp, C.POINTER(CRemdTemp), C.c_int, dp, C.c_int, C.POINTER(C.c_int), dp, dp]
    L.sdmd_remd_nccl_unique_id.argtypes = [C.c_char_p]
    L.sdmd_remd_nccl_init.argtypes = [vp, C.c_int, C.c_int, C.c_char_p]
    L.sdmd_remd_seed_server.argtypes = [vp, C.c_uint32]
    L.sdmd_remd_exchange.argtypes = [vp, C.c_double, C.POINTER(CRemdTemp), C.c_uint32]
    L.sdmd_remd_exchange_energy.argtypes = [vp, C.POINTER(CRemdTemp), dp]
    L.sdmd_remd_finish.argtypes = [vp]
    L.sdmd_host_glibc_rand.argtypes = [C.c_uint32, C.c_int64, C.POINTER(C.c_uint32)]
    L.sdmd_cuda_kernel_ms.argtypes = [vp, C.c_char_p]
    L.sdmd_cuda_kernel_ms.restype = C.c_double
    L.sdmd_cuda_kernel_stats.argtypes = [vp, C.c_char_p, dp, C.POINTER(C.c_int64)]
    L.sdmd_cuda_reset_kernel_stats.argtypes = [vp]
    L.sdmd_cuda_reset_kernel_stats.restype = None
    L.sdmd_cuda_set_timing.argtypes = [vp, C.c_int]
    L.sdmd_cuda_set_loop_mode.argtypes = [vp, C.c_int]
    L.sdmd_cuda_set_timing.restype = None
    L.sdmd_cuda_launch_count.argtypes = [vp]
    L.sdmd_cuda_launch_count.restype = C.c_int64
    L.sdmd_cuda_phase_cycles.argtypes = [vp, C.POINTER(C.c_int64)]
    L.sdmd_cuda_stream.argtypes = [vp]
    L.sdmd_cuda_stream.restype = vp
    _lib = L
    return L


def _check(rc, what):
    if rc != 0:
        raise SdmdError(f"{what} failed (rc={rc}): {lib().sdmd_cuda_last_error().decode()}")


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _ip(a):
    return a.ctypes.data_as(C.POINTER(C.c_int32))


def renew_list(atoms):
    atoms = [int(a) for a in atoms]
    return np.array([len(atoms), *atoms, 0, 0], dtype=np.int32)


class Replica:
    """One DMD system resident on one GPU (the reference's one-process-per-replica model)."""

    def __init__(self, hv: HostView, device: int | None = None):
        self.L = lib()
        self.hv = hv
        if device is None:
            device = int(os.environ.get("LOCAL_RANK", "0"))
        self._cs = hv.cstruct()
        h = C.c_void_p()
        _check(self.L.sdmd_cuda_create(C.byref(self._cs), int(device), C.byref(h)), "sdmd_cuda_create")
        self.h = h
        self.n = hv.natom
        cn, cs = (C.c_int * 3)(), (C.c_double * 3)()
        _check(self.L.sdmd_cuda_cell_grid(self.h, cn, cs), "cell_grid")
        self.cellnum = np.array(cn[:], dtype=np.int32)
        self.cellsize = np.array(cs[:])
        self.ncell = int(np.prod(self.cellnum.astype(np.int64)))

    def close(self):
        if getattr(self, "h", None):
            self.L.sdmd_cuda_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- transfers ----
    def upload_coords(self, hv: HostView | None = None):
        hv = hv or self.hv
        _check(self.L.sdmd_cuda_upload_coords(self.h, _dp(hv.x), _dp(hv.y), _dp(hv.z), _dp(hv.vx), _dp(hv.vy), _dp(hv.vz)), "upload_coords")

    def prepare(self, hv: HostView | None, gauss):
        """Prepare() from host inputs in one call (upload + PBC + LinkList + Predict(all) + CreateCBT)."""
        hv = hv or self.hv
        g = np.ascontiguousarray(gauss, dtype=np.float64)
        _check(self.L.sdmd_cuda_prepare(self.h, _dp(hv.x), _dp(hv.y), _dp(hv.z), _dp(hv.vx), _dp(hv.vy), _dp(hv.vz), _dp(g), len(g)), "prepare")

    def prepare_resident(self):
        """Prepare() on the state already in HBM: PBC + LinkList + Predict(all) + CreateCBT in one call."""
        _check(self.L.sdmd_cuda_prepare_resident(self.h), "prepare_resident")

    def upload_celllist(self, cl):
        cl = np.ascontiguousarray(cl, dtype=np.int32)
        assert len(cl) == self.n + self.ncell + 1
        _check(self.L.sdmd_cuda_upload_celllist(self.h, _ip(cl)), "upload_celllist")

    def download(self, hv: HostView | None = None) -> HostView:
        hv = hv or self.hv
        cs = hv.cstruct()
        _check(self.L.sdmd_cuda_download(self.h, C.byref(cs)), "download")
        hv.sync_scalars_from_c()
        return hv

    def download_slots(self):
        n1 = self.n + 1
        t, p = np.zeros(n1), np.zeros(n1)
        pa = np.zeros(n1, dtype=np.int32)
        ty, su = np.zeros(n1, dtype=np.uint8), np.zeros(n1, dtype=np.uint8)
        _check(self.L.sdmd_cuda_download_slots(self.h, _dp(t), _dp(p), _ip(pa), ty.ctypes.data_as(C.POINTER(C.c_uint8)),
                                               su.ctypes.data_as(C.POINTER(C.c_uint8))), "download_slots")
        return {"ev_time": t, "ev_potential": p, "ev_partner": pa, "ev_type": ty, "ev_sub": su}

    def celllist(self):
        cl = np.zeros(self.n + self.ncell + 1, dtype=np.int32)
        _check(self.L.sdmd_cuda_download_celllist(self.h, _ip(cl)), "download_celllist")
        return cl

    def cbt(self):
        node = np.zeros(2 * self.n, dtype=np.int32)
        leaf = np.zeros(self.n + 1, dtype=np.int32)
        _check(self.L.sdmd_cuda_download_cbt(self.h, _ip(node), _ip(leaf)), "download_cbt")
        return node, leaf

    # ---- thermostat stream ----
    def rng_feed(self, g):
        g = np.ascontiguousarray(g, dtype=np.float64)
        _check(self.L.sdmd_cuda_rng_feed(self.h, _dp(g), len(g)), "rng_feed")

    def rng_fill_libc(self, n):
        _check(self.L.sdmd_cuda_rng_fill_libc(self.h, int(n)), "rng_fill_libc")

    def rng_available(self):
        return int(self.L.sdmd_cuda_rng_available(self.h))

    def rng_seek(self, pos): _check(self.L.sdmd_cuda_rng_seek(self.h, int(pos)), "rng_seek")

    def rng_refeed(self, g):
        """Fresh stream: drop what was fed, then feed ``g``."""
        _check(self.L.sdmd_cuda_rng_reset(self.h), "rng_reset")
        self.rng_feed(g)
    def synchronize(self): _check(self.L.sdmd_cuda_synchronize(self.h), "synchronize")
    def timer_begin(self): _check(self.L.sdmd_cuda_timer_begin(self.h), "timer_begin")

    def timer_end(self) -> float:
        ms = C.c_double()
        _check(self.L.sdmd_cuda_timer_end(self.h, C.byref(ms)), "timer_end")
        return ms.value

    # ---- hot path, reference names ----
    def pbc_full(self): _check(self.L.sdmd_cuda_pbc_full(self.h), "pbc_full")
    def cell_build(self): _check(self.L.sdmd_cuda_cell_build(self.h), "cell_build")
    def predict_all(self): _check(self.L.sdmd_cuda_predict_all(self.h), "predict_all")
    def cbt_build(self): _check(self.L.sdmd_cuda_cbt_build(self.h), "cbt_build")
    def do_event(self): _check(self.L.sdmd_cuda_do_event(self.h), "do_event")

    def predict(self, atoms):
        r = renew_list(atoms)
        _check(self.L.sdmd_cuda_predict(self.h, _ip(r)), "predict")

    def cbt_update(self, atoms):
        r = renew_list(atoms)
        _check(self.L.sdmd_cuda_cbt_update(self.h, _ip(r)), "cbt_update")

    def assign_job(self):
        t, p, dt = C.c_int32(), C.c_int32(), C.c_double()
        r = np.zeros(4, dtype=np.int32)
        _check(self.L.sdmd_cuda_assign_job(self.h, C.byref(t), C.byref(p), C.byref(dt), _ip(r)), "assign_job")
        return t.value, p.value, dt.value, r

    def advance(self, dt): _check(self.L.sdmd_cuda_advance(self.h, float(dt)), "advance")
    def add_time(self, dt): _check(self.L.sdmd_cuda_add_time(self.h, float(dt)), "add_time")

    def hazard_check(self, target, partner):
        hz = C.c_int32()
        _check(self.L.sdmd_cuda_hazard_check(self.h, int(target), int(partner), C.byref(hz)), "hazard_check")
        return hz.value

    def cell_move(self, atom): _check(self.L.sdmd_cuda_cell_move(self.h, int(atom)), "cell_move")

    def potential_energy(self):
        e = C.c_double()
        _check(self.L.sdmd_cuda_potential_energy(self.h, C.byref(e)), "potential_energy")
        return e.value

    def kinetic_energy(self):
        e = C.c_double()
        _check(self.L.sdmd_cuda_kinetic_energy(self.h, C.byref(e)), "kinetic_energy")
        return e.value

    def run(self, t_until=1e300, max_iter=1 << 62, trace_cap=0):
        trace = np.zeros(max(trace_cap, 0), dtype=TRACE_DTYPE)
        st = CRunStats()
        rc = self.L.sdmd_cuda_run(self.h, float(t_until), int(max_iter), trace.ctypes.data if trace_cap else None,
                                  int(trace_cap), C.byref(st))
        _check(rc, "run")
        d = st.as_dict()
        return d, trace[: min(trace_cap, d["iterations"])] if trace_cap else trace

    def stats(self):
        st = CRunStats()
        _check(self.L.sdmd_cuda_get_stats(self.h, C.byref(st)), "get_stats")
        return st.as_dict()

    def set_temperature(self, t): _check(self.L.sdmd_cuda_set_temperature(self.h, float(t)), "set_temperature")

    # ---- measurement ----
    def kernel_stats(self, name):
        ms, n = C.c_double(), C.c_int64()
        self.L.sdmd_cuda_kernel_stats(self.h, name.encode(), C.byref(ms), C.byref(n))
        return ms.value, n.value

    def reset_kernel_stats(self): self.L.sdmd_cuda_reset_kernel_stats(self.h)
    def set_timing(self, on): self.L.sdmd_cuda_set_timing(self.h, int(on))   # 0 off, 1 stages, 2 + fix-up sub-stages
    def set_loop_mode(self, mode): _check(self.L.sdmd_cuda_set_loop_mode(self.h, int(mode)), "set_loop_mode")   # 0 auto, 1 one CTA, 2 sweep
    def launch_count(self): return int(self.L.sdmd_cuda_launch_count(self.h))
    def stream(self): return self.L.sdmd_cuda_stream(self.h)

    def phase_cycles(self):
        out = (C.c_int64 * 12)()
        _check(self.L.sdmd_cuda_phase_cycles(self.h, out), "phase_cycles")
        names = ["head", "advance", "event", "gather", "eval", "fold", "cbt", "iterations", "event_thread0", "wait_sweep",
                 "x10", "x11"]
        return dict(zip(names, out[:12]))

    # ---- REMD ----
    def remd_init(self, nranks, rank, unique_id: bytes, server_seed: int):
        _check(self.L.sdmd_remd_nccl_init(self.h, int(nranks), int(rank), unique_id), "remd_nccl_init")
        _check(self.L.sdmd_remd_seed_server(self.h, int(server_seed)), "remd_seed_server")

    def remd_seed_server(self, seed): _check(self.L.sdmd_remd_seed_server(self.h, int(seed)), "remd_seed_server")

    def remd_exchange_energy(self, num, T):
        """CalPotenE + exchange in one call (energy computed on the device, written into the all-gather's send record).
        Returns (num, T, this replica's energy)."""
        t = CRemdTemp(int(num), 0, float(T))
        e = C.c_double()
        _check(self.L.sdmd_remd_exchange_energy(self.h, C.byref(t), C.byref(e)), "remd_exchange_energy")
        return t.num, t.T, e.value

    def remd_finish(self): _check(self.L.sdmd_remd_finish(self.h), "remd_finish")

    def remd_exchange(self, epot, num, T):
        t = CRemdTemp(int(num), 0, float(T))
        _check(self.L.sdmd_remd_exchange(self.h, float(epot), C.byref(t), 0), "remd_exchange")
        return t.num, t.T


def run_many(reps, t_until, max_iter=1 << 62):
    """Run several replicas resident on one GPU concurrently (one persistent thread block each)."""
    n = len(reps)
    hs = (C.c_void_p * n)(*[r.h for r in reps])
    tu = np.ascontiguousarray(np.broadcast_to(np.asarray(t_until, dtype=np.float64), (n,)).copy())
    st = (CRunStats * n)()
    _check(lib().sdmd_cuda_run_many(hs, n, _dp(tu), int(max_iter), st), "run_many")
    return [s.as_dict() for s in st]


def potential_energy_many(reps):
    n = len(reps)
    hs = (C.c_void_p * n)(*[r.h for r in reps])
    e = np.zeros(n)
    _check(lib().sdmd_cuda_potential_energy_many(hs, n, _dp(e)), "potential_energy_many")
    return e.tolist()


def host_register(arr: np.ndarray):
    _check(lib().sdmd_cuda_host_register(arr.ctypes.data, arr.nbytes), "host_register")


def host_unregister(arr: np.ndarray):
    lib().sdmd_cuda_host_unregister(arr.ctypes.data)


def glibc_rand(seed: int, n: int) -> np.ndarray:
    """n outputs of glibc rand() after srand(seed) (the library's restatement; no process-global state touched)."""
    out = np.zeros(n, dtype=np.uint32)
    _check(lib().sdmd_host_glibc_rand(int(seed), int(n), out.ctypes.data_as(C.POINTER(C.c_uint32))), "glibc_rand")
    return out


def nccl_unique_id() -> bytes:
    buf = C.create_string_buffer(128)
    _check(lib().sdmd_remd_nccl_unique_id(buf), "nccl_unique_id")
    return buf.raw


def remd_rearrange(epot, nums, temps, u01):
    """Host-side sweep of Server.c:145-192 through the C-ABI."""
    n = len(epot)
    e = (C.c_double * n)(*[float(x) for x in epot])
    T = (CRemdTemp * n)()
    for i in range(n):
        T[i].num, T[i].T = int(nums[i]), float(temps[i])
    u = np.ascontiguousarray(u01, dtype=np.float64)
    used = C.c_int(0)
    att, acc = (C.c_double * n)(), (C.c_double * n)()
    _check(lib().sdmd_remd_rearrange(e, T, n, _dp(u), len(u), C.byref(used), att, acc), "remd_rearrange")
    return [T[i].num for i in range(n)], [T[i].T for i in range(n)], used.value, list(att), list(acc)
