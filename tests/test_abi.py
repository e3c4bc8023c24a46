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
