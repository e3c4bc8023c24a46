"""REMD: the exchange protocol (REMD.c / Server.c) restated in sdmd_b200/remd.py + the C-ABI sweep, checked against the
reference's own sServer + sDMD processes (CPU, oracle replicas), and the rank-parallel path under gloo."""
import os
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

from sdmd_b200 import capi
from sdmd_b200.remd import RemdEnsemble, ServerStream, reduced_temperature
from tests.remd_ref import REF, TEMPS_K, harness_snapshots, run_reference_remd

ROOT = Path(__file__).resolve().parents[1]


def test_glibc_rand_restatement_matches_libc():
    import ctypes
    capi.build()
    libc = ctypes.CDLL("libc.so.6")
    for seed in (1, 777, 1575263273):
        libc.srand(seed)
        want = np.array([libc.rand() for _ in range(300000)], dtype=np.uint32)
        assert np.array_equal(capi.glibc_rand(seed, 300000), want)


def _oracle_ensemble(views, rate, seed):
    from oracle.oracle_py import Oracle
    reps = []
    for hv in views:
        o = Oracle(hv)
        o.rng_feed(hv.extra["gauss"])
        o.set_celllist(hv.extra["celllist"])
        o.predict_all()
        o.cbt_build()
        reps.append(o)
    temps = [reduced_temperature(TEMPS_K[k]) for k in range(len(views))]
    return RemdEnsemble(reps, temps, rate, seed)


@pytest.mark.refbin
def test_remd_exchange_sequence_matches_reference_server(tmp_path):
    """3 replicas, exchange every 5 time units for 60: temperatures/ranks after every exchange equal the files the
    reference replicas write (out_REMD<id>.txt) when coupled through the reference's sServer."""
    if not (REF / "sServer").exists():
        pytest.skip("oracle/_ref not built")
    capi.build()
    nrep, rate, tu, seed = 3, 5.0, 60.0, 4242
    logs, server_out = run_reference_remd(tmp_path, nrep, rate, tu, seed)
    assert "connection 3 established" in server_out
    views = harness_snapshots(tmp_path, nrep, rate, tu, n_gauss=200000)
    for hv in views:
        hv.timestep = 1e12
    ens = _oracle_ensemble(views, rate, seed)
    n_x = len(logs[0].splitlines())
    assert n_x >= 10
    for _ in range(n_x):
        ens.step()
    swaps = 0
    for k in range(nrep):
        assert ens.log_text(k) == logs[k], (k, ens.log_text(k)[:200], logs[k][:200])
        swaps += sum(1 for rec in ens.log[k] if rec[2] != k + 1)
    assert swaps > 0   # the run really exchanged temperatures


def test_remd_two_ranks_gloo(tmp_path):
    """world_size 2 over gloo: each rank owns one (oracle) replica; the replicated sweep must give both ranks a
    consistent permutation identical to the single-process ensemble."""
    script = tmp_path / "rank.py"
    script.write_text('''
import os, sys, json
sys.path.insert(0, %r)
import numpy as np, torch.distributed as dist
from sdmd_b200.hostview import HostView
from sdmd_b200.remd import ServerStream, exchange_distributed, reduced_temperature, RemdEnsemble
from oracle.oracle_py import Oracle
TEMPS_K = [236.5, 250]
def make(k):
    pre = HostView.from_snapshot(%r)
    hv = pre.copy(); hv.timestep = 1e12
    o = Oracle(hv); o.rng_feed(np.random.default_rng(100 + k).standard_normal(200000))
    o.set_celllist(pre.extra["celllist"]); o.predict_all(); o.cbt_build()
    return o
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
rep = make(rank)
T, num = reduced_temperature(TEMPS_K[rank]), rank + 1
rep.set_temperature(T)
stream = ServerStream(99)
log = []
for x in range(1, 9):
    st, _ = rep.run(t_until=2.0 * x)
    e = rep.potential_energy()
    num, T = exchange_distributed(dist, rank, world, e, num, T, stream)
    rep.set_temperature(T)
    log.append((st["frame"], num, T))
if rank == 0:
    ens = RemdEnsemble([make(0), make(1)], [reduced_temperature(t) for t in TEMPS_K], 2.0, 99)
    for _ in range(8): ens.step()
    single = [(n, t) for (_, t, n) in ens.log[0]]
    assert single == [(n, t) for (_, n, t) in log], (single, log)
gathered = [None] * world
dist.all_gather_object(gathered, log)
if rank == 0:
    for x in range(8):
        assert sorted(g[x][1] for g in gathered) == [1, 2]
    print("REMD_GLOO_OK", json.dumps(log[-1]))
dist.destroy_process_group()
''' % (str(ROOT), str(ROOT / "tests" / "golden" / "ala10_pre.npz")))
    capi.build()
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                        "--master-port", "29611", str(script)], capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0 and "REMD_GLOO_OK" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]
