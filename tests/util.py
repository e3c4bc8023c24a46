"""Shared helpers for the parity tests."""
from __future__ import annotations

from pathlib import Path

import numpy as np

from sdmd_b200.hostview import HostView, TRACE_DTYPE

ROOT = Path(__file__).resolve().parents[1]
GOLDEN = ROOT / "tests" / "golden"
REFDIR = ROOT / "oracle" / "_ref"

STATE_FIELDS = ("x", "y", "z", "vx", "vy", "vz", "ev_time", "ev_potential", "ev_partner", "ev_pcounter", "ev_counter",
                "ev_type", "ev_sub", "type", "hb_bond", "hb_itype", "hbn_partner", "hbn_status", "hbn_partner_num",
                "cell_idx")
SLOT_FIELDS = ("ev_time", "ev_potential", "ev_partner", "ev_pcounter", "ev_counter", "ev_type", "ev_sub", "hb_itype",
               "hbn_partner_num")


def load_case(name, src=GOLDEN, ext="npz"):
    pre = HostView.from_snapshot(Path(src) / f"{name}_pre.{ext}")
    post = HostView.from_snapshot(Path(src) / f"{name}_post.{ext}")
    fin = HostView.from_snapshot(Path(src) / f"{name}_final.{ext}")
    return pre, post, fin


def ref_trace(fin: HostView):
    return np.frombuffer(fin.extra["trace"].tobytes(), dtype=TRACE_DTYPE)


def bits_equal(a, b):
    a, b = np.ascontiguousarray(a), np.ascontiguousarray(b)
    return a.shape == b.shape and a.dtype == b.dtype and a.tobytes() == b.tobytes()


def assert_fields_equal(got: HostView, want: HostView, fields, what=""):
    """Bit-exact comparison of per-atom arrays (entry 0 is unused by the reference and ignored)."""
    n1 = got.natom + 1
    for f in fields:
        a, b = getattr(got, f), getattr(want, f)
        k = len(a) // n1
        a, b = a[k:], b[k:]
        if not bits_equal(a, b):
            bad = np.nonzero(a != b)[0]
            raise AssertionError(f"{what}{f}: {len(bad)} of {len(a)} entries differ, first at flat index {bad[:5]}: "
                                 f"got {a[bad[:5]]} want {b[bad[:5]]}")


def assert_trace_equal(got, want):
    assert len(got) == len(want), (len(got), len(want))
    for f in TRACE_DTYPE.names:
        if f == "pad_":
            continue
        a, b = got[f], want[f]
        if f == "dt":
            a, b = a.view(np.int64), b.view(np.int64)
        if not np.array_equal(a, b):
            i = int(np.nonzero(a != b)[0][0])
            raise AssertionError(f"trace field {f} first differs at iteration {i}: got {got[i]} want {want[i]}")


def trace_times_close(got, want, rtol=1e-12):
    """north_star tolerance for times: 1e-12 relative (the CUDA path is expected to be bit-equal anyway)."""
    a, b = got["dt"], want["dt"]
    return np.all(np.abs(a - b) <= rtol * np.maximum(np.abs(a), np.abs(b)))
