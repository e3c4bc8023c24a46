#!/usr/bin/env python
"""bench.py — sDMD hot path on B200: one JSON line per run (see the contract in the task statement).

Workloads (``--workload``):
  pass1m  (default) BASELINE.json configs[4] / BASELINE.md C5(i): synthetic 1M-particle GEL box (hard core 4.0 A +
          step at 6.5 A, simple-cubic 8 A lattice + 0.3 A jitter, T*=0.6, Andersen 6.0).  One step = one
          full-system pass of the hot path: PBC("full") + LinkList("full") + Predict(all atoms) + CreateCBT.
          metric = pair time-predictions/s (CalculateParameters-equivalent pair evaluations).
  loop    C5(ii): the device-resident event loop on the same box; one step = 256 loop iterations;
          metric = DMD events/s (committed events).
  remd    configs[1]: ala10 replica exchange, one replica per GPU, temperatures of demo/REMD/REMDConfig.txt,
          ExchangeRate 50; one step = run to the next exchange boundary + CalPotenE + exchange (NCCL all-gather);
          metric = REMD replica-events/s (sum over replicas).

The default line also carries two sub-records measured in the same run: ``loop`` (events/s of the device-resident loop on
the same 1M box) and ``remd`` (ala10, one replica per rank, temperatures of demo/REMD/REMDConfig.txt, NCCL exchange every
50 tu: replica-events/s, events per rank, microseconds per exchange and a parity bit — the (num, T) walk of every rank
against the same ensemble run in ONE process with the Server.c sweep on the host).

``--impl reference`` times the reference's CPU implementation of the same path on the host cores, one worker process per
rank of the launch (rank 0 starts them all): the UNMODIFIED reference itself — oracle/_ref/ref_bench_gel for the 1M box
(its dense connectionMap answered by a zero window, see oracle/ref_bench.c), sDMD_ref + sServer for the remd sub-record —
or, where oracle/_ref is absent, the oracle port (oracle/sdmd_oracle.c, pinned bit-exact against the reference at 10^6 atoms).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

REMD_TEMPS_K = [236.5, 250, 263.5, 277.5, 292.5, 308, 324.5, 341]  # first 8 of demo/REMD/REMDConfig.txt
KB = 0.0019872041  # REMDExec.c:19


def reduced_T(kelvin: float) -> float:
    return float("%lf" % (kelvin * KB))  # sREMD prints "%lf", the replica parses it with atof


# ----------------------------------------------------------------------------------------------------
REAL_STDOUT_FD = None      # the process's real stdout while fd 1 is pointed at stderr (main)


_T0 = time.perf_counter()


def mark(dist, what):
    """Progress marker on stderr (every rank): where a run is, should it ever sit somewhere."""
    sys.stderr.write(f"[bench r{dist.rank}/{dist.world} +{time.perf_counter() - _T0:6.1f}s] {what}\n")
    sys.stderr.flush()


def ensure_built(capi, dist):
    """The shared library is built by __graft_entry__.build() and travels with the tree; a bench rank never recompiles behind the
    back of the others (time stamps of a copied tree are no basis for N ranks to start N nvcc runs on one box).  Only when the
    library is missing does local rank 0 build it, the others wait."""
    if capi.LIB_PATH.exists() and os.environ.get("SDMD_BENCH_BUILD", "") != "1":
        return
    if dist.local_rank == 0:
        capi.build()
    dist.barrier()


def run_guarded(fn, limit_s, line, key, why):
    """Run `fn()`; if it has not returned after `limit_s` seconds, print `line` (the bench line built so far, rank 0 only; None on
    the other ranks) with `line[key] = {"unavailable": why}` on the process's real stdout and leave the process with exit code 0."""
    def give_up():
        if line is not None:
            line[key] = {"unavailable": why}
            os.write(REAL_STDOUT_FD if REAL_STDOUT_FD is not None else 1, (json.dumps(line) + "\n").encode())   # (fd 1 may point at stderr, see main)
        try:
            sys.stdout.flush(); sys.stderr.flush()
        finally:
            os._exit(0)

    guard = threading.Timer(limit_s, give_up)
    guard.daemon = True
    guard.start()
    try:
        return fn()
    finally:
        guard.cancel()


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, gpu_index: int, active: bool = True):
        super().__init__(daemon=True)
        self.gpu = gpu_index
        self.samples = []
        self.stop_flag = threading.Event()
        self.proc = None
        # only the rank that prints the line samples: one nvidia-smi loop per box, not one per rank (every query takes the
        # driver's global lock; eight of them polling at 100 ms on an 8-GPU board are a load on every CUDA call of every rank)
        self.active = active

    def run(self):
        if not self.active:
            return
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.gpu)], stdout=subprocess.PIPE, text=True)
            for line in self.proc.stdout:
                self.samples.append([s.strip() for s in line.split(",")])
                if self.stop_flag.is_set():
                    break
        except Exception:
            pass

    def finish(self):
        self.stop_flag.set()
        if self.proc:
            self.proc.terminate()
        sm = [float(s[1]) for s in self.samples if len(s) > 7 and s[1].replace(".", "").isdigit()]
        mx = [float(s[2]) for s in self.samples if len(s) > 7 and s[2].replace(".", "").isdigit()]
        reasons = set()
        for s in self.samples:
            if len(s) > 7:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), s[4:8]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def load_peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class Dist:
    """torch.distributed plumbing: rendezvous, barrier, max-over-ranks."""

    def __init__(self, want_gpus: int, backend: str | None = None):
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        self.torch = None
        if self.world > 1:
            import torch
            import torch.distributed as dist
            self.torch, self.dist = torch, dist
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            # the collectives of this file are barriers and 32-byte exchanges: NVLink-SHARP (multicast) set-up buys them nothing, and
            # two communicators (torch's and the library's own) both setting it up on an 8-GPU box is start-up time and risk for free
            os.environ.setdefault("NCCL_NVLS_ENABLE", "0")
            os.environ.setdefault("NCCL_DEBUG", "WARN")      # (NCCL's own warnings and errors on stderr, next to the progress markers)
            if backend is None:
                backend = "nccl" if torch.cuda.is_available() else "gloo"
            if backend == "nccl":
                torch.cuda.set_device(self.local_rank)
            dist.init_process_group(backend=backend)
            self.backend = backend

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()

    def _tensor(self, vals):
        t = self.torch.tensor(vals, dtype=self.torch.float64)
        return t.cuda() if self.backend == "nccl" else t

    def max(self, x: float) -> float:
        if self.world == 1:
            return x
        t = self._tensor([x])
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def sum(self, x: float) -> float:
        if self.world == 1:
            return x
        t = self._tensor([x])
        self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        return float(t.item())

    def gather_vec(self, vals, world: int):
        """Every rank contributes a short float vector; returns the list of all of them (all-reduce of a zero-padded matrix)."""
        if self.world == 1:
            return [list(vals)]
        m = [[0.0] * len(vals) for _ in range(world)]
        m[self.rank] = [float(v) for v in vals]
        t = self._tensor(m)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        return [[float(x) for x in row] for row in t.cpu().tolist()]

    def bcast_bytes(self, b: bytes | None, n: int) -> bytes:
        if self.world == 1:
            return b
        t = self.torch.zeros(n, dtype=self.torch.uint8)
        if self.rank == 0:
            t = self.torch.frombuffer(bytearray(b), dtype=self.torch.uint8).clone()
        if self.backend == "nccl":
            t = t.cuda()
        self.dist.broadcast(t, src=0)
        return bytes(t.cpu().numpy().tobytes())

    def close(self):
        if self.world > 1:
            self.dist.destroy_process_group()


# ----------------------------------------------------------------------------------------------------
# workload: pass1m
# ----------------------------------------------------------------------------------------------------
def make_box(n_axis, rank):
    from sdmd_b200.synth import gel_box
    hv, rng = gel_box(n_axis, a=8.0, jitter=0.3, seed=1575263273 + rank)
    gauss = rng.standard_normal(hv.natom)
    return hv, gauss


def cuda_pass(args, dist):
    from sdmd_b200 import capi
    ensure_built(capi, dist)
    n_axis = args.n_axis
    hv, gauss = make_box(n_axis, dist.rank)
    n = hv.natom
    rep = capi.Replica(hv, device=dist.local_rank)
    rep.rng_feed(gauss)
    mark(dist, "replica created")

    def one_pass():     # Prepare() of the reference on the resident state: PBC + LinkList + Predict(all) + CreateCBT, one C-ABI call
        rep.rng_seek(0)
        rep.prepare_resident()

    # -------- device-resident timing (`value`) --------
    sampler = ClockSampler(dist.local_rank, active=dist.rank == 0)
    sampler.start()
    t_w = time.perf_counter()
    nwarm = 0
    while nwarm < args.warmup or time.perf_counter() - t_w < 0.5:   # >= W warm-up passes, and long enough for clock samples
        one_pass()
        nwarm += 1
    rep.synchronize()
    pairs0 = rep.stats()["pair_predictions"]
    rep.reset_kernel_stats()
    dist.barrier()
    rep.timer_begin()
    for _ in range(args.steps):
        one_pass()
    ms = rep.timer_end()
    mark(dist, "device-resident passes timed")
    dist.barrier()
    ms = dist.max(ms)
    launches = rep.launch_count()
    pairs = rep.stats()["pair_predictions"] - pairs0
    total_pairs = dist.sum(float(pairs))
    kst = {k: rep.kernel_stats(k) for k in ("predict_all", "predict_all_fixup", "cell_build", "cbt_build", "predict_all_rounds", "predict_all_requeued")}
    t_k = time.perf_counter()
    while time.perf_counter() - t_k < 0.3:      # keep the same load up while the 100 ms sampler catches the region
        one_pass()
    rep.synchronize()
    clocks = sampler.finish()
    clocks["note"] = "sampled every 100 ms from warm-up through the timed region and 0.3 s of identical load after it"

    # -------- end to end through the C-ABI with HOST buffers (`e2e`) --------
    host = [np.ascontiguousarray(a.copy()) for a in (hv.x, hv.y, hv.z, hv.vx, hv.vy, hv.vz)]
    g_host = np.ascontiguousarray(gauss.copy())
    for a in host + [g_host]:
        capi.host_register(a)
    hv_e2e = hv  # arrays used by upload_coords
    hv_e2e.x, hv_e2e.y, hv_e2e.z, hv_e2e.vx, hv_e2e.vy, hv_e2e.vz = host
    h2d = sum(a.nbytes for a in host) + g_host.nbytes
    d2h = 120 + 24  # sdmd_run_stats + AssignJob result

    def one_e2e():
        rep.prepare(hv_e2e, g_host)     # upload + PBC + LinkList + Predict(all) + CreateCBT, one C-ABI call
        rep.assign_job()
        return rep.stats()

    for _ in range(max(1, args.warmup // 2)):
        one_e2e()
    pairs0 = rep.stats()["pair_predictions"]
    dist.barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        one_e2e()
    rep.synchronize()
    e2e_serial_s = dist.max(time.perf_counter() - t0)
    mark(dist, "end-to-end passes (one at a time) timed")
    e2e_serial_pairs = dist.sum(float(rep.stats()["pair_predictions"] - pairs0))

    # The same K steps with two of them in flight: a second replica context (own stream, own device buffers) driven by a
    # second host thread, so the upload of one step's 56 MB overlaps the kernels of the other.  Every step still uploads its
    # inputs from page-locked host memory and reads its result back inside the timed region; the link is the limit then.
    import threading
    rep2 = capi.Replica(hv, device=dist.local_rank)
    rep2.rng_feed(gauss)

    def worker(r, k):
        for _ in range(k):
            r.prepare(hv_e2e, g_host)
            r.assign_job()
            r.stats()

    for r in (rep, rep2):
        worker(r, 2)
    p0 = rep.stats()["pair_predictions"] + rep2.stats()["pair_predictions"]
    dist.barrier()
    t0 = time.perf_counter()
    th = [threading.Thread(target=worker, args=(rep, (args.steps + 1) // 2)), threading.Thread(target=worker, args=(rep2, args.steps // 2))]
    for t in th:
        t.start()
    for t in th:
        t.join()
    rep.synchronize(); rep2.synchronize()
    e2e_s = dist.max(time.perf_counter() - t0)
    mark(dist, "end-to-end passes (two in flight) timed")
    e2e_pairs = dist.sum(float(rep.stats()["pair_predictions"] + rep2.stats()["pair_predictions"] - p0))
    rep2.close()
    for a in host + [g_host]:
        capi.host_unregister(a)

    # -------- extras: advance roofline + event loop on the same box --------
    rep.reset_kernel_stats()
    for _ in range(50):
        rep.advance(1e-7)
    adv_ms, adv_n = rep.kernel_stats("advance")
    extra = {"advance_us": 1e3 * adv_ms / adv_n, "advance_GBps": 88.0 * n / (adv_ms / adv_n) / 1e6}
    rep.rng_refeed(np.random.default_rng(7).standard_normal(n + 200000))
    rep.cell_build(); rep.predict_all(); rep.cbt_build()
    rep.run(max_iter=256)            # first call allocates the second coordinate buffer of the overlapped sweep
    st0 = rep.stats()
    t0 = time.perf_counter()
    st1, _ = rep.run(max_iter=2048)
    loop_s = time.perf_counter() - t0
    extra["loop_events_per_s_1M"] = (st1["frame"] - st0["frame"]) / loop_s
    extra["loop_iterations_per_s_1M"] = (st1["iterations"] - st0["iterations"]) / loop_s

    peak, peak_src = load_peaks()
    pa_ms, pa_n = kst["predict_all"]
    per_launch_ms = pa_ms / max(pa_n, 1)
    alg_bytes = 88.0 * n  # 56 B own state read + 32 B slot written per atom
    achieved = alg_bytes / (per_launch_ms * 1e-3) / 1e9
    rows = os.environ.get("SDMD_K2_ROWS", "1") != "0" and os.environ.get("SDMD_K2_TILE", "") not in ("1", "2")
    roof = {"bound": "hbm", "kernel": ("k_r0_filter_rows + k_r0_eval_rows" if rows else "k_predict_all_filter + k_predict_all_round") + " (round 0 of the full-system Predict)",
            "achieved": achieved, "peak": peak, "unit": "GB/s",
            "frac": achieved / peak, "traffic": None, "peak_source": peak_src, "algorithmic_bytes_per_launch": alg_bytes,
            "kernel_ms": per_launch_ms, "kernel_share_of_step": pa_ms / ms if ms else None,
            "pairs_per_s_kernel": (pairs / max(pa_n, 1)) / (per_launch_ms * 1e-3),
            "note": "HBM fraction as north_star asks; the stage is not HBM bound: ~41 candidate distance tests (single-precision "
                    "superset filter, instruction-issue bound: issue slots 69 % busy) + ~6 double-precision root solves per atom "
                    "(evaluation kernel: issue 44 %, FP64 pipe 23 %, warps wait on their first loads); see profiles/r2_summary.md"}
    # dram__bytes_read+write per launch of the two kernels from the committed `ncu --set full` capture, if recorded
    tf = ROOT / "profiles" / "ncu_traffic.json"
    if tf.exists():
        try:
            t = json.loads(tf.read_text())
            tile = os.environ.get("SDMD_K2_TILE", "") == "1"
            names = ["k_predict_all_tile"] if tile else (["k_r0_filter_rows", "k_r0_eval_rows<1>"] if rows else ["k_predict_all_filter", "k_predict_all_round<1>"])
            if int(t.get("natom", 0)) == n and all(k in t["kernels"] for k in names):
                roof["traffic"] = sum(float(t["kernels"][k]["dram_bytes"]) for k in names)
                roof["traffic_source"] = t.get("source", "profiles/ncu_traffic.json")
                roof["traffic_kernels"] = names
        except Exception:
            pass
    step_ms = ms / args.steps
    out = {
        "metric": "pair time-predictions/s", "value": total_pairs / (ms * 1e-3), "unit": "pairs/s", "n_gpus": dist.world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": step_ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "C5(i) synthetic 1M-particle GEL hard-sphere+step box: full-system pass PBC+LinkList+Predict(all)+CreateCBT",
                   "natom_per_gpu": n, "lattice": f"{n_axis}^3 simple cubic, a=8.0 A, jitter 0.3 A", "T_reduced": 0.6,
                   "cutoff_A": 9.0, "parallelism": f"replicas x{dist.world} (one independent box per GPU, no collective)",
                   "l2": "working set ~190 MB per pass > 126 MB L2 (no explicit flush)"},
        "e2e": {"value": e2e_pairs / e2e_s, "unit": "pairs/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": 1e3 * e2e_s / args.steps, "steps_in_flight": 2,
                "one_step_at_a_time": {"value": e2e_serial_pairs / e2e_serial_s, "ms_per_step": 1e3 * e2e_serial_s / args.steps},
                "note": "sdmd_cuda_prepare + assign_job + stats per step from page-locked host arrays; two replica contexts on two host threads keep two steps in flight (upload of one under the kernels of the other)"},
        "gpu_launches": launches, "clocks": clocks, "roofline": roof,
        "kernels_ms_per_step": {k: v[0] / args.steps for k, v in kst.items() if not k.startswith("predict_all_r")},
        "predict_all_requeued_atoms": kst["predict_all_requeued"][0] / max(kst["predict_all_requeued"][1], 1),
        "predict_all_rounds": kst["predict_all_rounds"][0] / max(kst["predict_all_rounds"][1], 1),
        "extra": extra,
    }
    out["loop"] = {"metric": "DMD events/s", "value": extra["loop_events_per_s_1M"], "unit": "events/s",
                   "iterations_per_s": extra["loop_iterations_per_s_1M"], "natom": n,
                   "sample": "2048 loop iterations of the device-resident event loop on the same box (sdmd_cuda_run, host wall clock "
                             "around the call, stats read back); `--workload loop` times it on its own with the sweep's roofline"}
    rep.close()
    mark(dist, "extras done, pass replica closed")
    if dist.rank == 0 and not args.no_cpu_baseline:
        out["cpu_baseline"] = cpu_pass_baseline(args, gpu_pairs_per_pass=int(round(pairs / max(args.steps, 1))))
        mark(dist, "CPU baseline done")
    if not args.no_remd:
        # The sub-record is an extra of the line, not its metric: it must not be able to take the line with it.  (An 8-rank run of
        # this round sat for six minutes somewhere behind the NCCL start-up and was killed by the box's limit before printing
        # anything; 1, 2 and 4 ranks finish the sub-record in ~15 s.)  If it has not finished in time, rank 0 prints the line
        # without it and every rank leaves.
        limit = float(os.environ.get("SDMD_BENCH_REMD_LIMIT_S", "150"))
        r = run_guarded(lambda: remd_subrecord(args, dist), limit, out if dist.rank == 0 else None, "remd",
                        f"the remd sub-record did not finish within {limit:.0f} s on {dist.world} ranks and was abandoned")
        if dist.rank == 0:
            out["remd"] = r
    return out


def port_pass(n_axis, rank=0, steps=1):
    """The oracle port (reference algorithm, plain C, 1 thread) on the bench box of `rank`: (pairs per pass, best seconds)."""
    from oracle.oracle_py import Oracle
    hv, gauss = make_box(n_axis, rank)
    o = Oracle(hv)
    t_best, pairs = None, 0
    for _ in range(steps):
        hv.reset_slots()
        o.rng_feed(gauss)
        p0 = o.stats()["pair_predictions"]
        t0 = time.perf_counter()
        o.pbc_full(); o.cell_build(); o.predict_all(); o.cbt_build()
        dt = time.perf_counter() - t0
        pairs = o.stats()["pair_predictions"] - p0
        t_best = dt if t_best is None else min(t_best, dt)
    o.close()
    return pairs, t_best


REF_BENCH = ROOT / "oracle" / "_ref" / "ref_bench_gel"


def reference_pass_workers(n_axis, workers, passes, warm=1):
    """`workers` processes of the UNMODIFIED reference (oracle/_ref/ref_bench_gel), each timing `warm + passes` full-system
    passes over the same n_axis^3 GEL box (8 A lattice, 0.3 A jitter; the reference draws its own velocities).  Returns
    (seconds per timed pass for every worker, pair evaluations per pass counted by the pinned port on worker 0's state)."""
    import shutil
    import tempfile
    sys.path.insert(0, str(ROOT))
    from oracle.make_golden import prepare_input
    from oracle.oracle_py import Oracle
    from tests import ref_bench as rb             # test infrastructure around oracle/_ref, used by this CPU leg only
    write_jittered_gel_gro = rb.write_jittered_gel_gro
    work = Path(tempfile.mkdtemp(prefix="s", dir="/tmp"))
    try:
        inp = prepare_input(work, time_units=10.0)
        gro = f"GEL{n_axis}.gro"
        write_jittered_gel_gro(inp / gro, n_axis, 0.8, 0.03, 1575263273)
        procs = []
        for w in range(workers):
            (work / f"out{w}").mkdir()
            cmd = [str(REF_BENCH), str(warm + passes), "-"] + ([str(work / "state.bin")] if w == 0 else []) + \
                  ["--", "-f", gro, "-i", str(inp), "-o", str(work / f"out{w}")]
            # stdout to a file: the reference prints megabytes while reading a 10^6-atom box, and a worker blocked on a full pipe
            # would not run side by side with the others
            procs.append(subprocess.Popen(cmd, stdout=open(work / f"stdout{w}.txt", "wb"), stderr=subprocess.DEVNULL, cwd=work))
        secs = []
        for w, pr in enumerate(procs):
            if pr.wait(timeout=1800) != 0:
                raise RuntimeError("ref_bench_gel failed")
            out = (work / f"stdout{w}.txt").read_bytes().decode(errors="replace")
            s = [float(ln.split()[3]) for ln in out.splitlines() if ln.startswith("pass ")]
            secs.append(s[warm:])
        # the reference keeps no evaluation counter: the port (pinned bit-exact to the reference at this size) counts the pair
        # evaluations of the very pass worker 0 ran, from the state that worker dumped
        hv, g = rb.view_from_state(work / "state.bin", work)
        o = Oracle(hv); o.rng_feed(g); o.pbc_full(); o.cell_build(); o.predict_all()
        pairs = o.stats()["pair_predictions"]
        o.close()
    finally:
        shutil.rmtree(work, ignore_errors=True)
    return secs, pairs


def cpu_pass_baseline(args, gpu_pairs_per_pass=None):
    """CPU baseline of the default workload on one host core: the unmodified reference when oracle/_ref holds it, and the
    port on the very box the GPU timed (which also pins the pair count of the GPU arm)."""
    pairs, t_port = port_pass(args.n_axis, 0, 1)
    port = {"value": pairs / t_port, "unit": "pairs/s", "cores": 1, "kind": "port", "seconds": t_port, "pairs_per_pass": pairs}
    if gpu_pairs_per_pass is not None:
        port["gpu_pairs_per_pass"] = gpu_pairs_per_pass
        port["pair_count_equal"] = bool(gpu_pairs_per_pass == pairs)
        if gpu_pairs_per_pass != pairs:
            raise SystemExit(f"bench: GPU pair count {gpu_pairs_per_pass} != CPU port {pairs} on the same box")
    if REF_BENCH.exists():
        secs, rpairs = reference_pass_workers(args.n_axis, 1, passes=3, warm=1)
        t = min(secs[0])
        return {"value": rpairs / t, "unit": "pairs/s", "cores": 1, "kind": "reference", "seconds": t,
                "sample": f"3 full-system passes (after 1 warm-up pass) of the unmodified -DGEL reference over a {args.n_axis}^3-atom box of the same "
                          f"lattice (oracle/_ref/ref_bench_gel: PBC+LinkList+Predict(all)+CreateCBT, {rpairs} pair evaluations per pass, best "
                          f"{t:.2f} s); its dense connectionMap is answered by a zero window (oracle/ref_bench.c)",
                "port": port}
    port["sample"] = (f"one full-system pass over the same {args.n_axis ** 3}-atom box ({pairs} pair evaluations, {t_port:.2f} s) with the oracle "
                      "port, pinned bit-exact to the reference at this size; oracle/_ref/ref_bench_gel is absent")
    return port


def reference_pass(args, dist):
    """--impl reference: one CPU worker per rank of the launch (rank 0 starts them all), each on its own host core."""
    if dist.rank != 0:
        return None
    W = dist.world
    cores = os.cpu_count() or 1
    cfg = {"workload": "C5(i) synthetic 1M-particle GEL hard-sphere+step box: full-system pass PBC+LinkList+Predict(all)+CreateCBT",
           "natom_per_gpu": args.n_axis ** 3, "lattice": f"{args.n_axis}^3 simple cubic, a=8.0 A, jitter 0.3 A", "T_reduced": 0.6,
           "cutoff_A": 9.0, "parallelism": f"replicas x{W} (one independent box per CPU worker, no exchange)",
           "l2": "n/a (host cores)"}
    if REF_BENCH.exists():
        secs, pairs = reference_pass_workers(args.n_axis, W, passes=args.steps, warm=max(1, min(args.warmup, 2)))
        per_worker = [sum(s) for s in secs]
        tot = max(per_worker)
        nsteps = len(secs[0])
        v = W * pairs * nsteps / tot
        kind, sample = "reference", (f"{W} process(es) of the unmodified -DGEL reference (oracle/_ref/ref_bench_gel), {nsteps} timed passes each "
                                     f"over a {args.n_axis}^3-atom box; {pairs} pair evaluations per pass (counted by the pinned port)")
    else:
        import multiprocessing as mp
        with mp.get_context("spawn").Pool(W) as pool:
            res = pool.starmap(port_pass, [(args.n_axis, r, args.steps) for r in range(W)])
        pairs = res[0][0]
        tot = max(t for _, t in res) * args.steps
        nsteps = args.steps
        v = sum(p for p, _ in res) * nsteps / tot
        kind, sample = "port", f"{W} process(es) of the oracle port, best of {nsteps} passes each (oracle/_ref absent)"
    out = {
        "impl": "reference", "metric": "pair time-predictions/s", "value": v, "unit": "pairs/s", "n_gpus": W,
        "steps": nsteps, "warmup": args.warmup, "ms_per_step": 1e3 * tot / nsteps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": cfg,
        "cpu_baseline": {"value": v, "unit": "pairs/s", "cores": min(W, cores), "kind": kind, "sample": sample, "host_cores": cores},
        "e2e": {"value": v, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    if not args.no_remd:
        out["remd"] = reference_remd_server(args, W)
    return out


# ----------------------------------------------------------------------------------------------------
# workload: loop (C5(ii): device-resident event loop on the synthetic box)
# ----------------------------------------------------------------------------------------------------
LOOP_ITERS = 256


def cuda_loop(args, dist):
    from sdmd_b200 import capi
    ensure_built(capi, dist)
    hv, gauss = make_box(args.n_axis, dist.rank)
    n = hv.natom
    rng = np.random.default_rng(99 + dist.rank)
    rep = capi.Replica(hv, device=dist.local_rank)
    rep.rng_feed(gauss)
    rep.pbc_full(); rep.cell_build(); rep.predict_all(); rep.cbt_build()
    g_step = np.ascontiguousarray(rng.standard_normal(8 * LOOP_ITERS))
    capi.host_register(g_step)

    def one_step():
        rep.rng_feed(g_step)                       # host -> device: this step's thermostat draws
        st, _ = rep.run(max_iter=LOOP_ITERS)       # device -> host: sdmd_run_stats
        return st

    sampler = ClockSampler(dist.local_rank, active=dist.rank == 0)
    sampler.start()
    for _ in range(args.warmup):
        one_step()
    st0 = rep.stats()
    rep.reset_kernel_stats()
    dist.barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        st = one_step()
    rep.synchronize()
    dt = dist.max(time.perf_counter() - t0)
    dist.barrier()
    run_ms, _ = rep.kernel_stats("run")          # CUDA events around the launches of sdmd_cuda_run
    run_ms = dist.max(run_ms)
    clocks = sampler.finish()
    ev = dist.sum(float(st["frame"] - st0["frame"]))
    its = dist.sum(float(st["iterations"] - st0["iterations"]))
    launches = rep.launch_count()
    rep.reset_kernel_stats()
    for _ in range(50):
        rep.advance(1e-9)
    adv_ms, adv_n = rep.kernel_stats("advance")
    adv = adv_ms / adv_n
    peak, peak_src = load_peaks()
    alg = 88.0 * n
    per_it_ms = run_ms / max(its / dist.world, 1)
    out = {
        "metric": "DMD events/s", "value": ev / (run_ms * 1e-3), "unit": "events/s", "n_gpus": dist.world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": run_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"C5(ii) device-resident event loop on the synthetic {n}-atom GEL box, {LOOP_ITERS} loop iterations per step",
                   "natom_per_gpu": n, "parallelism": f"replicas x{dist.world} (one independent box per GPU, no collective)",
                   "l2": "per-event working set 88 MB > L2 is streamed (evict-first) by the sweep blocks"},
        "e2e": {"value": ev / dt, "unit": "events/s", "h2d_bytes_per_step": g_step.nbytes, "d2h_bytes_per_step": 120},
        "iterations_per_s": its / (run_ms * 1e-3), "gpu_launches": launches, "clocks": clocks,
        # one launch per iteration (k_sweep_event): the sweep role streams the 88 B/atom while the event block works; the
        # launch lasts as long as the slower of the two, so the achieved figure is algorithmic bytes over the iteration time
        "roofline": {"bound": "hbm", "kernel": ("k_sweep_event (UpdateData+TimeForward sweep overlapped with the event block, one launch per iteration)"
                                                if n > 98304 else "k_run_cluster (one persistent cluster of 16 CTAs: block 0 runs the events, all advance)"
                                                if n > 2048 else "k_run_cta (one persistent CTA)"),
                     "achieved": alg / (per_it_ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                     "frac": alg / (per_it_ms * 1e-3) / 1e9 / peak, "traffic": None, "peak_source": peak_src,
                     "algorithmic_bytes_per_launch": alg, "kernel_ms": per_it_ms, "kernel_share_of_step": 1.0,
                     "sweep_alone_ms": adv, "sweep_alone_frac": alg / (adv * 1e-3) / 1e9 / peak,
                     "note": "the iteration is bound by the latency of the single event block (dependent loads under a "
                             "bandwidth-saturating sweep), not by the sweep: k_advance alone reaches sweep_alone_frac"},
    }
    if dist.rank == 0 and not args.no_cpu_baseline:
        from oracle.oracle_py import Oracle
        ho, go = make_box(args.n_axis, 0)
        o = Oracle(ho)
        o.rng_feed(np.concatenate([go, np.random.default_rng(5).standard_normal(4096)]))
        o.pbc_full(); o.cell_build(); o.predict_all(); o.cbt_build()
        t1 = time.perf_counter()
        so, _ = o.run(max_iter=200)
        t2 = time.perf_counter() - t1
        out["cpu_baseline"] = {"value": so["frame"] / t2, "unit": "events/s", "cores": 1, "kind": "port",
                               "sample": f"200 loop iterations of the same box with the oracle port ({t2:.2f} s); the O(N) UpdateData/TimeForward sweeps dominate"}
    rep.close()
    return out


# ----------------------------------------------------------------------------------------------------
# workload: remd (ala10, one replica per GPU)
# ----------------------------------------------------------------------------------------------------
def ala10_view(rank):
    from sdmd_b200.hostview import HostView
    pre = HostView.from_snapshot(ROOT / "tests" / "golden" / "ala10_pre.npz")
    hv = pre.copy()
    hv.target_temperature = reduced_T(REMD_TEMPS_K[rank % len(REMD_TEMPS_K)])
    hv.timestep = 1e12
    return hv


def remd_ladder_K(n):
    """Temperature ladder: the demo's own first 8 values, or a geometric ladder over the same span for larger ensembles."""
    if n <= len(REMD_TEMPS_K):
        return list(REMD_TEMPS_K[:n])
    return [float("%.2f" % t) for t in np.geomspace(REMD_TEMPS_K[0], REMD_TEMPS_K[-1], n)]


REMD_SERVER_SEED = 12345
REMD_GAUSS = 400000          # thermostat draws fed to every replica up front (a 50 tu period of ala10 uses a few thousand)


def _ala10_replica(capi, replica_id, device):
    """ala10 replica `replica_id` ready to run: temperature of the demo ladder, its own glibc rand() thermostat stream."""
    import ctypes
    hv = ala10_view(replica_id)
    rep = capi.Replica(hv, device=device)
    rep.set_timing(0)
    ctypes.CDLL("libc.so.6").srand(1575263273 + replica_id)
    rep.rng_fill_libc(hv.natom + REMD_GAUSS)
    rep.pbc_full(); rep.cell_build(); rep.predict_all(); rep.cbt_build()
    return hv, rep


def remd_subrecord(args, dist):
    """BASELINE configs[1] inside the default run: one ala10 replica per rank, exchange every --exchange-rate tu through
    sdmd_remd_exchange_energy (CalPotenE on the device -> NCCL all-gather of 32 B per rank -> replicated Server.c sweep).
    The walk of (num, T) over all ranks is compared with the same ensemble run in one process on rank 0's GPU."""
    from sdmd_b200 import capi
    from sdmd_b200.remd import RemdEnsemble
    W, periods, rate = dist.world, args.remd_periods, float(args.exchange_rate)
    hv, rep = _ala10_replica(capi, dist.rank, dist.local_rank)
    if W > 1:
        uid = capi.nccl_unique_id() if dist.rank == 0 else None
        uid = dist.bcast_bytes(uid, 128)
        rep.remd_init(W, dist.rank, uid, server_seed=REMD_SERVER_SEED)
        mark(dist, "remd: NCCL communicator of the library initialised")
    else:
        rep.remd_seed_server(REMD_SERVER_SEED)
    state = {"num": dist.rank + 1, "T": hv.target_temperature, "next": rate}
    walk, xchg_s = [], []

    def period():
        st, _ = rep.run(t_until=state["next"])
        if st["stop_reason"] != 1:
            raise SystemExit(f"remd sub-record: unexpected stop {st['stop_reason']} (thermostat stream exhausted?)")
        t0 = time.perf_counter()
        state["num"], state["T"], _e = rep.remd_exchange_energy(state["num"], state["T"])
        xchg_s.append(time.perf_counter() - t0)
        state["next"] += rate
        walk.append((state["num"], state["T"]))
        return st

    period()                                            # warm-up period (also the first, lazily connecting, collective)
    mark(dist, "remd: warm-up period and first exchange done")
    f0 = rep.stats()["frame"]
    xchg_s.clear()
    dist.barrier()
    t0 = time.perf_counter()
    for _ in range(periods):
        st = period()
    rep.synchronize()
    dt = time.perf_counter() - t0
    dist.barrier()
    dt = dist.max(dt)
    mine = float(st["frame"] - f0)
    per_rank = dist.gather_vec([mine], W)               # events of every rank during the timed periods
    walks = dist.gather_vec([v for nt in walk for v in nt], W)
    xs = dist.max(float(np.mean(xchg_s)))
    rep.close()
    mark(dist, "remd: timed periods done")
    out = {"metric": "REMD replica-events/s", "value": sum(r[0] for r in per_rank) / dt, "unit": "events/s", "replicas": W,
           "periods_timed": periods, "exchange_rate_tu": rate, "temperatures_K": REMD_TEMPS_K[:W], "seconds": dt,
           "events_per_rank": [int(r[0]) for r in per_rank], "us_per_exchange_call": 1e6 * xs,
           "us_per_exchange_note": "mean time inside sdmd_remd_exchange_energy on the slowest rank's clock: CalPotenE kernel + collective + sweep, "
                                   "INCLUDING the wait for the last replica to reach the boundary",
           "exchange": "sdmd_remd_exchange_energy: CalPotenE on the device written into the send record, ncclAllGather of 32 B per rank, "
                       "replicated Server.c:145-192 sweep" if W > 1 else "single replica: CalPotenE + sweep over one entry, no collective",
           "limit": "the period lasts as long as the hottest replica (most events per 50 tu) needs"}
    if dist.rank == 0:
        # parity: the same replicas in ONE process, coupled by the host-side sweep (tests pin that ensemble to the reference's sServer)
        reps = []
        for r in range(W):
            _, rr = _ala10_replica(capi, r, dist.local_rank)
            reps.append(rr)
        ens = RemdEnsemble(reps, [reduced_T(REMD_TEMPS_K[r]) for r in range(W)], rate, server_seed=REMD_SERVER_SEED)
        for _ in range(periods + 1):
            ens.step()
        want = [[v for rec in ens.log[r] for v in (rec[2], rec[1])] for r in range(W)]
        got = [list(w) for w in walks]
        out["parity"] = bool(all(len(a) == len(b) and all(float(x) == float(y) for x, y in zip(a, b)) for a, b in zip(got, want)))
        out["walk_num"] = [[int(w[2 * k]) for k in range(periods + 1)] for w in got]
        out["swaps_seen"] = int(sum(1 for r in range(W) for k in range(periods + 1) if int(got[r][2 * k]) != r + 1))
        for rr in reps:
            rr.close()
    return out


def reference_remd_server(args, W):
    """Reference arm of the remd sub-record: W unmodified sDMD processes coupled through the reference's own sServer, on the
    host cores; start-up (parsing, SDEnergyMin, first Predict) is measured by a 0.05-time-unit run and subtracted."""
    ref = ROOT / "oracle" / "_ref"
    if not (ref / "sDMD_ref").exists() or not (ref / "sServer").exists():
        return {"unavailable": "oracle/_ref/sDMD_ref or sServer not built"}
    import shutil
    import socket
    import tempfile
    rate = float(args.exchange_rate)
    tu_full = rate * (args.remd_periods + 1) + 1.0

    def run_all(time_units):
        s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
        server = subprocess.Popen([str(ref / "sServer"), str(port), str(W), str(REMD_SERVER_SEED)], stdout=subprocess.DEVNULL,
                                  stderr=subprocess.DEVNULL)
        time.sleep(0.2)
        procs, dirs = [], []
        t0 = time.perf_counter()
        for r in range(W):
            d = Path(tempfile.mkdtemp(prefix=f"sr{r}_", dir="/tmp"))
            shutil.copytree(ref / "input", d / "input")
            lines = (ref / "demo_REMD" / "parameter.txt").read_text().splitlines()
            lines[1] = "Time               %g" % time_units
            (d / "input" / "parameter.txt").write_text("\n".join(lines) + "\n")
            (d / "out").mkdir()
            # stdout to a FILE: the replicas wait for each other at every exchange, so one of them blocked on a full pipe
            # (the reference prints its progress continuously) would stall all of them
            procs.append(subprocess.Popen([str(ref / "sDMD_ref"), "-i", str(d / "input"), "-f", "ala10.gro", "-o", str(d / "out"), "-REMD", "localhost",
                                           str(port), "%lf" % reduced_T(REMD_TEMPS_K[r % len(REMD_TEMPS_K)]), str(r + 1), str(int(rate)), str(r)],
                                          stdout=open(d / "stdout.txt", "w"), stderr=subprocess.DEVNULL))
            dirs.append(d)
        events = 0
        for pr, d in zip(procs, dirs):
            pr.wait(timeout=900)
            for line in (d / "stdout.txt").read_text(errors="replace").replace("\r", "\n").splitlines():
                if line.startswith("Calculate event times:"):
                    events += int(line.split(":")[1])
        dt = time.perf_counter() - t0
        try:
            server.wait(timeout=20)
        except subprocess.TimeoutExpired:
            server.kill()
        for d in dirs:
            shutil.rmtree(d, ignore_errors=True)
        return events, dt

    ev0, t_start = run_all(0.05)
    ev1, t_full = run_all(tu_full)
    dt = max(t_full - t_start, 1e-9)
    return {"metric": "REMD replica-events/s", "value": (ev1 - ev0) / dt, "unit": "events/s", "replicas": W, "seconds": dt,
            "cores": min(W, os.cpu_count() or 1), "kind": "reference", "exchange_rate_tu": rate, "temperatures_K": REMD_TEMPS_K[:W],
            "sample": f"{W} unmodified sDMD_ref processes + the reference's sServer, {tu_full:g} time units each; start-up ({t_start:.2f} s) "
                      "measured by a 0.05-time-unit run and subtracted"}


def cuda_remd_many(args, dist):
    """One REMD ensemble of R ala10 replicas per GPU: every replica is a persistent thread block (sdmd_cuda_run_many),
    CalPotenE for all of them in one launch, the Server.c sweep on the host.  GPUs hold independent ensembles."""
    from sdmd_b200 import capi
    from sdmd_b200.remd import RemdEnsemble
    ensure_built(capi, dist)
    R = args.replicas_per_gpu
    temps_K = remd_ladder_K(R)
    rate = float(args.exchange_rate)
    per_step = 260000                                   # gaussians one replica may consume per exchange interval (generous)
    rng = np.random.default_rng(1575263273 + dist.rank)
    pool = np.ascontiguousarray(rng.standard_normal(per_step * 4 + 8 * 1024 * 1024))   # synthetic thermostat stream, host memory
    capi.host_register(pool)
    reps = []
    for k in range(R):
        hv = ala10_view(0)
        hv.target_temperature = reduced_T(temps_K[k])
        rep = capi.Replica(hv, device=dist.local_rank)
        rep.set_timing(0)
        off = int(rng.integers(0, 8 * 1024 * 1024))
        rep.rng_feed(pool[off:off + per_step * 2])
        rep.pbc_full(); rep.cell_build(); rep.predict_all(); rep.cbt_build()
        reps.append(rep)
    fed = [0]

    def refill(k, rep):
        if rep.rng_available() < per_step:
            off = int(rng.integers(0, 8 * 1024 * 1024))
            rep.rng_feed(pool[off:off + per_step])     # host -> device inside the timed region
            fed[0] += per_step * 8

    ens = RemdEnsemble(reps, [reduced_T(t) for t in temps_K], rate, server_seed=12345 + dist.rank, refill=refill)

    def totals():
        st = [r.stats() for r in reps]
        return sum(s["frame"] for s in st), sum(s["iterations"] for s in st)

    for _ in range(args.warmup):
        ens.step()
    f0, it0 = totals()
    fed[0] = 0
    l0 = sum(r.launch_count() for r in reps)
    sampler = ClockSampler(dist.local_rank, active=dist.rank == 0)
    sampler.start()
    dist.barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        ens.step()
    reps[0].synchronize()
    dt = time.perf_counter() - t0
    dist.barrier()
    clocks = sampler.finish()
    dt = dist.max(dt)
    f1, it1 = totals()
    ev = dist.sum(float(f1 - f0))
    its = dist.sum(float(it1 - it0))
    launches = sum(r.launch_count() for r in reps) - l0
    peak, peak_src = load_peaks()
    out = {
        "metric": "REMD replica-events/s", "value": ev / dt, "unit": "events/s", "n_gpus": dist.world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "ala10 (tests/golden/ala10_pre.npz, from the reference's input/ala10.gro); thermostat stream synthetic N(0,1)",
        "config": {"workload": f"ala10 REMD, {R} replicas per GPU (one persistent thread block each), one ensemble per GPU",
                   "replicas_per_gpu": R, "temperatures_K": [temps_K[0], temps_K[-1]], "exchange_rate_tu": rate, "natom": 59,
                   "parallelism": f"replica-parallel {R} x {dist.world} GPUs, host-side Server.c sweep per GPU, no collective"},
        "e2e": {"value": ev / dt, "unit": "events/s", "h2d_bytes_per_step": fed[0] // max(args.steps, 1),
                "d2h_bytes_per_step": R * (120 + 8)},
        "iterations_per_s": its / dt, "gpu_launches": launches, "clocks": clocks,
        "roofline": {"bound": "hbm", "kernel": "k_run_cta_multi (one persistent CTA per replica, latency bound per CTA)", "achieved": None,
                     "peak": peak, "unit": "GB/s", "frac": None, "traffic": None, "peak_source": peak_src,
                     "us_per_iteration_per_replica": 1e6 * dt * R * dist.world / max(its, 1)},
    }
    capi.host_unregister(pool)
    for r in reps:
        r.close()
    return out


def cuda_remd(args, dist):
    if args.replicas_per_gpu > 1:
        return cuda_remd_many(args, dist)
    from sdmd_b200 import capi
    ensure_built(capi, dist)
    hv = ala10_view(dist.rank)
    rep = capi.Replica(hv, device=dist.local_rank)
    if dist.world > 1:
        uid = capi.nccl_unique_id() if dist.rank == 0 else None
        uid = dist.bcast_bytes(uid, 128)
        rep.remd_init(dist.world, dist.rank, uid, server_seed=12345)
    else:
        rep.remd_seed_server(12345)
    np.random.seed(1 + dist.rank)
    import ctypes
    libc = ctypes.CDLL("libc.so.6")
    libc.srand(1575263273 + dist.rank)
    rep.rng_fill_libc(hv.natom + 400000)
    rep.pbc_full(); rep.cell_build(); rep.predict_all(); rep.cbt_build()
    num, T = dist.rank + 1, hv.target_temperature
    rate = float(args.exchange_rate)
    next_x = rate
    state = {"num": num, "T": T, "next": next_x}

    def one_step():
        if rep.rng_available() < 300000:
            rep.rng_fill_libc(300000)
        st, _ = rep.run(t_until=state["next"])
        e = rep.potential_energy()
        state["num"], state["T"] = rep.remd_exchange(e, state["num"], state["T"])
        state["next"] += rate
        return st

    for _ in range(args.warmup):
        one_step()
    rep.reset_kernel_stats()
    f0 = rep.stats()["frame"]
    it0 = rep.stats()["iterations"]
    sampler = ClockSampler(dist.local_rank, active=dist.rank == 0)
    sampler.start()
    dist.barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        st = one_step()
    rep.synchronize()
    dt = time.perf_counter() - t0
    dist.barrier()
    clocks = sampler.finish()
    dt = dist.max(dt)
    ev = dist.sum(float(st["frame"] - f0))
    its = dist.sum(float(st["iterations"] - it0))
    run_ms, run_n = rep.kernel_stats("run")
    peak, peak_src = load_peaks()
    out = {
        "metric": "REMD replica-events/s", "value": ev / dt, "unit": "events/s", "n_gpus": dist.world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "ala10 (tests/golden/ala10_pre.npz, from the reference's input/ala10.gro)",
        "config": {"workload": "ala10 REMD, one replica per GPU", "temperatures_K": REMD_TEMPS_K[:dist.world],
                   "exchange_rate_tu": rate, "natom": hv.natom, "parallelism": f"replica-parallel x{dist.world}, NCCL all-gather exchange"},
        "e2e": {"value": ev / dt, "unit": "events/s", "h2d_bytes_per_step": 8 * 300000, "d2h_bytes_per_step": 120 + 8},
        "iterations_per_s": its / dt, "gpu_launches": rep.launch_count(), "clocks": clocks,
        "roofline": {"bound": "hbm", "kernel": "k_run_cta (persistent single-CTA event loop, latency bound)", "achieved": None,
                     "peak": peak, "unit": "GB/s", "frac": None, "traffic": None, "peak_source": peak_src,
                     "us_per_iteration": 1e3 * run_ms / max(its / dist.world, 1)},
    }
    rep.close()
    return out


def reference_remd(args, dist):
    """The unmodified reference (oracle/_ref/sDMD_ref) on ala10, one process per replica on the host cores."""
    if dist.rank != 0:
        return None
    ref = ROOT / "oracle" / "_ref" / "sDMD_ref"
    if not ref.exists():
        return {"impl": "reference", "unavailable": "oracle/_ref/sDMD_ref not built"}
    import shutil
    import tempfile
    R = dist.world * max(1, args.replicas_per_gpu)
    ladder = remd_ladder_K(args.replicas_per_gpu) if args.replicas_per_gpu > 1 else REMD_TEMPS_K
    tu = float(args.exchange_rate) * (args.steps + args.warmup)

    def run_all(time_units):
        procs, dirs = [], []
        t0 = time.perf_counter()
        for r in range(R):
            d = Path(tempfile.mkdtemp(prefix=f"sdmd_ref_{r}_"))
            shutil.copytree(ROOT / "oracle" / "_ref" / "input", d / "input")
            p = d / "input" / "parameter.txt"
            lines = p.read_text().splitlines()
            lines[1] = "Time               %g" % time_units
            lines[2] = "T(reduced)         %g" % reduced_T(ladder[r % len(ladder)])
            p.write_text("\n".join(lines) + "\n")
            (d / "out").mkdir()
            # stdout to a file: a process blocked on a full pipe (the reference prints its progress continuously) would not run
            # side by side with the others
            procs.append(subprocess.Popen([str(ref), "-f", "ala10.gro", "-i", str(d / "input"), "-o", str(d / "out")],
                                          stdout=open(d / "stdout.txt", "w"), stderr=subprocess.DEVNULL))
            dirs.append(d)
        events = 0
        for p, d in zip(procs, dirs):
            p.wait(timeout=1800)
            out = (d / "stdout.txt").read_text(errors="replace")
            for line in out.replace("\r", "\n").splitlines():
                if line.startswith("Calculate event times:"):
                    events += int(line.split(":")[1])
        dt = time.perf_counter() - t0
        for d in dirs:
            shutil.rmtree(d, ignore_errors=True)
        return events, dt

    # start-up (parsing, SDEnergyMin, first full prediction) is timed separately and taken out, so the figure is the
    # reference's steady-state event rate
    ev0, t_start = run_all(0.05)
    ev1, t_full = run_all(tu)
    events = ev1 - ev0
    dt = max(t_full - t_start, 1e-9)
    v = events / dt
    return {"impl": "reference", "metric": "REMD replica-events/s", "value": v, "unit": "events/s", "n_gpus": dist.world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / max(args.steps + args.warmup, 1),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "ala10",
            "config": {"workload": "ala10 REMD, one replica per GPU" if args.replicas_per_gpu <= 1 else
                       f"ala10 REMD, {args.replicas_per_gpu} replicas per GPU (one persistent thread block each), one ensemble per GPU",
                       "temperatures_K": ladder[:R] if args.replicas_per_gpu <= 1 else [ladder[0], ladder[-1]]},
            "cpu_baseline": {"value": v, "unit": "events/s", "cores": min(R, os.cpu_count() or 1), "kind": "reference",
                             "sample": f"{R} unmodified sDMD processes x {tu:g} time units, start-up ({t_start:.2f} s: parsing + SDEnergyMin) "
                                       f"measured by a 0.05-time-unit run and subtracted; no exchange server"},
            "e2e": {"value": v, "unit": "events/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}


# ----------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="cuda", choices=["cuda", "reference"])
    ap.add_argument("--workload", default="pass1m", choices=["pass1m", "loop", "remd"])
    ap.add_argument("--n-axis", type=int, default=100, help="lattice points per axis of the synthetic box (100 -> 1M atoms)")
    ap.add_argument("--exchange-rate", type=float, default=50.0)
    ap.add_argument("--replicas-per-gpu", type=int, default=1, help="remd workload: independent REMD ensembles packed on each GPU")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-remd", action="store_true", help="default workload: skip the remd sub-record")
    ap.add_argument("--remd-periods", type=int, default=3, help="timed exchange periods of the remd sub-record (one more runs as warm-up)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "cuda" else args.warmup

    # the contract is ONE JSON line on stdout: anything a library prints there (NCCL's version banner, the reference's own
    # progress output) goes to stderr instead; the real stdout comes back for the final line
    sys.stdout.flush()
    global REAL_STDOUT_FD
    real_stdout = REAL_STDOUT_FD = os.dup(1)
    os.dup2(2, 1)
    dist = Dist(args.gpus, backend="gloo" if args.impl == "reference" else None)
    try:
        if args.impl == "reference":
            out = reference_pass(args, dist) if args.workload == "pass1m" else reference_remd(args, dist)
        else:
            out = {"pass1m": cuda_pass, "loop": cuda_loop, "remd": cuda_remd}[args.workload](args, dist)
        if dist.rank == 0 and out is not None:
            sys.stdout.flush()
            os.dup2(real_stdout, 1)
            print(json.dumps(out))
            sys.stdout.flush()
    finally:
        dist.close()


if __name__ == "__main__":
    main()
