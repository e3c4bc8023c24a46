"""CPU: bench.py's CUDA arm end to end with the C-ABI replaced by a stand-in (no GPU here).  What is checked is the host side
of the bench — argument handling, the order of the sections, the JSON line and its keys, the CPU baseline leg (real: the oracle
port and, where built, the unmodified reference), the guarded remd sub-record — so that an edit to bench.py cannot leave the
driver without a line.  The numbers in the line are meaningless."""
import json
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]

FAKE = r'''
import sys, types, time
import numpy as np
sys.path.insert(0, %(root)r)
from pathlib import Path
from oracle.oracle_py import Oracle

fake = types.ModuleType("sdmd_b200.capi")
fake.LIB_PATH = Path(%(root)r) / "bench.py"          # "exists"
fake.build = lambda *a, **k: fake.LIB_PATH
fake.host_register = lambda a: None
fake.host_unregister = lambda a: None
fake.nccl_unique_id = lambda: bytes(128)

class Replica:
    def __init__(self, hv, device=0):
        self.hv = hv; self.g = None; self.per_pass = None
        self.pairs = 0; self.frame = 0; self.iters = 0; self.t = 0.0; self.launches = 0
    def _pass(self):
        if self.per_pass is None:       # the pair evaluations of one pass over this box, counted by the oracle port on the same inputs
            ho = self.hv.copy(); ho.reset_slots(); o = Oracle(ho)
            o.rng_feed(self.g if self.g is not None else np.zeros(ho.natom + 8))
            o.pbc_full(); o.cell_build(); o.predict_all(); o.cbt_build()
            self.per_pass = o.stats()["pair_predictions"]; o.close()
        self.pairs += self.per_pass; self.launches += 11
    def rng_feed(self, g): self.g = np.array(g, dtype=np.float64)
    def rng_refeed(self, g): pass
    def rng_seek(self, p): pass
    def rng_fill_libc(self, n): pass
    def rng_available(self): return 1 << 30
    def set_timing(self, l): pass
    def pbc_full(self): pass
    def cell_build(self): pass
    def predict_all(self): self._pass()
    def cbt_build(self): pass
    def prepare_resident(self): self._pass()
    def prepare(self, hv, g): self._pass()
    def assign_job(self): return (1, 0, 0.1)
    def advance(self, dt): pass
    def synchronize(self): pass
    def reset_kernel_stats(self): self.launches = 0
    def timer_begin(self): self._t0 = time.perf_counter()
    def timer_end(self): return 1e3 * (time.perf_counter() - self._t0) + 1.0
    def launch_count(self): return self.launches
    def kernel_stats(self, name): return (2.0, 20)
    def stats(self): return {"pair_predictions": self.pairs, "frame": self.frame, "iterations": self.iters, "currenttime": self.t, "stop_reason": 1}
    def run(self, t_until=None, max_iter=None, **k):
        self.frame += 1000; self.iters += 1700
        if t_until is not None: self.t = t_until
        return self.stats(), None
    def remd_seed_server(self, s): pass
    def remd_init(self, *a, **k): pass
    def remd_exchange_energy(self, num, T): return num, T, -1.0
    def potential_energy(self): return -1.0
    def close(self): pass
fake.Replica = Replica
sys.modules["sdmd_b200.capi"] = fake
import sdmd_b200
sdmd_b200.capi = fake
import sdmd_b200.remd as remd

class Ens:
    def __init__(self, reps, temps, rate, server_seed=0, **k):
        self.temps = temps; self.rate = rate; self.log = [[] for _ in reps]; self.k = 0
    def step(self):
        self.k += 1
        for r, T in enumerate(self.temps): self.log[r].append((self.k * self.rate, T, r + 1))
remd.RemdEnsemble = Ens
import bench
sys.argv = ["bench.py"] + %(argv)r
bench.main()
'''


def _run(argv, env=None):
    import os
    e = dict(os.environ, **(env or {}))
    r = subprocess.run([sys.executable, "-c", FAKE % {"root": str(ROOT), "argv": argv}], capture_output=True, text=True, timeout=600, env=e, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-3000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, r.stdout[-2000:]          # the contract: ONE JSON line on stdout
    return json.loads(lines[0]), r.stderr


def test_default_line_has_every_key_the_contract_names():
    d, err = _run(["--n-axis", "12", "--steps", "3", "--warmup", "3"])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype",
              "data", "config", "e2e", "gpu_launches", "clocks", "roofline", "cpu_baseline", "remd", "loop"):
        assert k in d, k
    assert d["metric"] == "pair time-predictions/s" and d["unit"] == "pairs/s" and d["n_gpus"] == 1 and d["higher_is_better"] is True
    assert d["steps"] == 3 and d["warmup"] == 3 and d["value"] > 0 and d["gpu_launches"] > 0
    assert {"value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"} <= set(d["e2e"])
    assert {"bound", "achieved", "peak", "unit", "frac", "traffic"} <= set(d["roofline"]) and d["roofline"]["bound"] == "hbm"
    assert {"value", "unit", "cores", "kind", "sample"} <= set(d["cpu_baseline"]) and d["cpu_baseline"]["kind"] in ("reference", "port")
    assert d["remd"]["metric"] == "REMD replica-events/s" and "parity" in d["remd"]
    assert "[bench r0/1" in err and "CPU baseline done" in err          # the progress markers


def test_line_survives_a_sub_record_that_never_returns():
    """The remd sub-record under its watchdog: with a limit of zero seconds the line must still come out, marked."""
    d, _ = _run(["--n-axis", "12", "--steps", "3", "--warmup", "3"], env={"SDMD_BENCH_REMD_LIMIT_S": "0.001"})
    assert d["value"] > 0 and "cpu_baseline" in d
    assert "unavailable" in d["remd"] or "metric" in d["remd"]          # (a fast machine may still finish first)


def test_two_ranks_over_gloo(tmp_path):
    """The same under torchrun with two ranks (gloo on CPU): rendezvous, barriers, max / sum over ranks, the unique-id broadcast and
    the gathers of the remd sub-record; rank 0 alone prints."""
    script = tmp_path / "dry.py"
    script.write_text(FAKE % {"root": str(ROOT), "argv": ["--gpus", "2", "--n-axis", "12", "--steps", "3", "--warmup", "3"]})
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", "29577", str(script)], capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-3000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip().startswith("{")]
    assert len(lines) == 1, r.stdout[-2000:]
    d = json.loads(lines[0])
    assert d["n_gpus"] == 2 and d["value"] > 0 and d["remd"]["replicas"] == 2 and len(d["remd"]["events_per_rank"]) == 2
    assert "[bench r1/2" in r.stderr and "[bench r0/2" in r.stderr
