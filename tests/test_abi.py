"""CPU: the C-ABI library builds for sm_100a, loads, and exports every symbol include/sdmd_cuda.h declares."""
import ctypes
import re
from pathlib import Path

import pytest

from sdmd_b200 import capi

ROOT = Path(__file__).resolve().parents[1]


@pytest.fixture(scope="module")
def lib_path():
    return capi.build()


def declared_symbols():
    text = (ROOT / "include" / "sdmd_cuda.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(sdmd_(?:cuda|remd|host)_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_exported(lib_path):
    L = ctypes.CDLL(str(lib_path))
    syms = declared_symbols()
    assert len(syms) >= 40
    missing = [s for s in syms if not hasattr(L, s)]
    assert not missing, missing
    assert sorted(set(capi.ABI_SYMBOLS)) == syms  # the Python binding covers the whole header


def test_no_device_fails_loudly(lib_path):
    """Without a GPU the product path must fail, never fall back to a CPU implementation."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from sdmd_b200.synth import gel_box
    hv, _ = gel_box(3)
    with pytest.raises(capi.SdmdError):
        capi.Replica(hv, device=0)


def test_product_does_not_import_oracle():
    """The oracle is test infrastructure: nothing under sdmd_b200/ may import, include, link or load it."""
    pat = re.compile(r"(^\s*(from|import)\s+oracle\b)|(#include\s*[\"<][^\">]*oracle)|liboracle|oracle_py|sdmd_oracle", re.M)
    for p in (ROOT / "sdmd_b200").rglob("*"):
        if p.suffix in (".py", ".cu", ".cuh", ".c", ".h"):
            assert not pat.search(p.read_text()), p


def test_remd_rearrange_matches_oracle(lib_path):
    import numpy as np
    from oracle.oracle_py import remd_rearrange as oracle_rr
    rng = np.random.default_rng(5)
    for trial in range(200):
        n = int(rng.integers(2, 13))
        e = rng.normal(-30, 8, n)
        temps = np.sort(rng.uniform(0.4, 0.9, n))
        order = rng.permutation(n)
        nums = (order + 1).tolist()
        T = temps[order].tolist()
        if trial % 7 == 0:
            T[int(rng.integers(n))] = 0.0  # a finished replica
        u = rng.uniform(0, 1, 2 * n + 2)
        assert capi.remd_rearrange(e, nums, T, u) == oracle_rr(e, nums, T, u)


def test_bench_guard_prints_the_line_and_leaves(tmp_path):
    """bench.run_guarded: a sub-record that does not come back must not take the bench line with it — the line is printed on the
    real stdout with the sub-record marked unavailable, exit code 0; a sub-record that does come back is returned untouched."""
    import json
    import subprocess
    import sys
    from pathlib import Path
    root = Path(__file__).resolve().parents[1]
    prog = (
        "import os, sys, time, json\n"
        f"sys.path.insert(0, {str(root)!r})\n"
        "import bench\n"
        "bench.REAL_STDOUT_FD = os.dup(1)\n"
        "os.dup2(2, 1)\n"                                      # like bench.main: library chatter goes to stderr
        "line = {'metric': 'm', 'value': 1.0}\n"
        "assert bench.run_guarded(lambda: {'ok': 1}, 5.0, line, 'remd', 'late') == {'ok': 1}\n"
        "bench.run_guarded(lambda: time.sleep(30), 0.3, line, 'remd', 'did not finish')\n"
        "os.write(bench.REAL_STDOUT_FD, b'NOT REACHED\\n')\n")
    r = subprocess.run([sys.executable, "-c", prog], capture_output=True, text=True, timeout=60)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, r.stdout
    d = json.loads(lines[0])
    assert d["value"] == 1.0 and d["remd"] == {"unavailable": "did not finish"}
