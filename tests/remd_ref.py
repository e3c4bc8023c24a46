"""Test helper: run the reference's own REMD (sServer + N x sDMD_ref -REMD ...) and the harness snapshots."""
from __future__ import annotations

import shutil
import socket
import subprocess
import time
from pathlib import Path

from sdmd_b200.hostview import HostView
from sdmd_b200.remd import reduced_temperature

ROOT = Path(__file__).resolve().parents[1]
REF = ROOT / "oracle" / "_ref"
TEMPS_K = [236.5, 250, 263.5, 277.5, 292.5, 308, 324.5, 341]


def free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _input_dir(work: Path, time_units: float) -> Path:
    inp = work / "input"
    shutil.copytree(REF / "input", inp)
    lines = (REF / "demo_REMD" / "parameter.txt").read_text().splitlines()
    lines[1] = "Time               %g" % time_units
    (inp / "parameter.txt").write_text("\n".join(lines) + "\n")
    (work / "out").mkdir(parents=True, exist_ok=True)
    return inp


def remd_args(port, T, num, rate, rid):
    return ["-REMD", "localhost", str(port), "%lf" % T, str(num), str(int(rate)), str(rid)]


def _wait_connected(proc, port, timeout=180.0):
    import psutil
    t0 = time.time()
    while time.time() - t0 < timeout:
        if proc.poll() is not None:
            return
        try:
            for c in psutil.Process(proc.pid).net_connections(kind="tcp"):
                if c.status == psutil.CONN_ESTABLISHED and c.raddr and c.raddr.port == port:
                    return
        except (psutil.Error, OSError):
            pass
        time.sleep(0.02)
    raise AssertionError("replica did not connect to the exchange server")


def run_reference_remd(work: Path, nrep: int, rate: float, time_units: float, server_seed: int, binary=None):
    """Returns the text of out_REMD<id>.txt per replica (SaveRE, DataSave.c:687-700)."""
    port = free_port()
    server = subprocess.Popen([str(REF / "sServer"), str(port), str(nrep), str(server_seed)], stdout=subprocess.PIPE,
                              stderr=subprocess.STDOUT, text=True)
    time.sleep(0.3)
    procs = []
    for r in range(nrep):
        d = work / f"ref{r}"
        inp = _input_dir(d, time_units)
        T = reduced_temperature(TEMPS_K[r])
        procs.append(subprocess.Popen([str(binary or REF / "sDMD_ref"), "-i", str(inp), "-f", "ala10.gro", "-o", str(d / "out"),
                                       *remd_args(port, T, r + 1, rate, r)], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL))
        # sREMD staggers the replicas by sleep(1 + id) (REMDExec.c:133) so that they connect — lazily, at their first
        # exchange (REMD.c:150-153) — in replica-id order, which is the order the server sweeps them in.  A fixed sleep is
        # a race on a busy or cold machine (start-up, CUDA initialisation), so wait for this replica's connection itself.
        _wait_connected(procs[-1], port)
    for p in procs:
        p.wait(timeout=600)
    try:
        out = server.communicate(timeout=30)[0]
    except subprocess.TimeoutExpired:
        server.kill()
        out = ""
    logs = []
    for r in range(nrep):
        f = list((work / f"ref{r}" / "out").glob("out_REMD*"))
        assert f, list((work / f"ref{r}" / "out").iterdir())
        logs.append(f[0].read_text())
    return logs, out


def harness_snapshots(work: Path, nrep: int, rate: float, time_units: float, n_gauss: int):
    """Per-replica start state (after SDEnergyMin + PBC + LinkList, SGThread.c:42-45) and gaussian stream."""
    views = []
    for r in range(nrep):
        d = work / f"snap{r}"
        inp = _input_dir(d, time_units)
        T = reduced_temperature(TEMPS_K[r])
        cmd = [str(REF / "ref_harness"), str(d / "rep"), "0", str(n_gauss), "--", "-i", str(inp), "-f", "ala10.gro",
               "-o", str(d / "out"), *remd_args(1, T, r + 1, rate, r)]
        res = subprocess.run(cmd, capture_output=True, text=True)
        assert res.returncode == 0, res.stdout[-1500:] + res.stderr[-1500:]
        views.append(HostView.from_snapshot(d / "rep_pre.snap"))
    return views
