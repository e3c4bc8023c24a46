"""CPU: pin the oracle restatement (oracle/sdmd_oracle.c) against dumps of the unmodified reference.

The committed fixtures under tests/golden/ were produced by oracle/make_golden.py from the reference
itself (oracle/_ref/ref_harness).  When the compiled reference is present the longest case is also
re-generated on this machine and compared, so the pin does not depend on the fixtures alone.
"""
import os

import numpy as np
import pytest

# the live harness runs of this module flatten the reference's globals through the glue (sdmd_refglue.c): have it cross-check its
# sparse connection rows against the dense connectionMap scan every time
os.environ["SDMD_GLUE_CHECK_CM"] = "1"

from oracle.oracle_py import Oracle, remd_rearrange
from tests.util import (REFDIR, SLOT_FIELDS, STATE_FIELDS, assert_fields_equal, assert_trace_equal, bits_equal,
                        load_case, ref_trace)


def _prepare(pre):
    hv = pre.copy()
    o = Oracle(hv)
    o.rng_feed(pre.extra["gauss"])
    return hv, o


@pytest.mark.parametrize("case", ["ala10", "gel10", "ala10x8", "ala10x8_wallp"])
def test_cell_grid_and_link_list(case):
    pre, post, fin = load_case(case)
    hv, o = _prepare(pre)
    assert np.array_equal(o.cellnum, pre.extra["cellnum"])
    assert bits_equal(o.cellsize, pre.extra["cellsize"])
    hv.cell_idx[:] = 0
    o.pbc_full()
    o.cell_build()
    assert np.array_equal(o.celllist(), pre.extra["celllist"])
    assert_fields_equal(hv, pre, ("cell_idx", "x", "y", "z"))


@pytest.mark.parametrize("case", ["ala10", "gel10", "ala10x8", "ala10x8_wallp"])
def test_predict_all_and_cbt(case):
    pre, post, fin = load_case(case)
    hv, o = _prepare(pre)
    o.set_celllist(pre.extra["celllist"])
    o.predict_all()
    assert_fields_equal(hv, post, SLOT_FIELDS, what="after Predict(all): ")
    o.cbt_build()
    node, leaf = o.cbt()
    assert np.array_equal(node[1:], post.extra["cbt_node"][1:])
    assert np.array_equal(leaf[1:], post.extra["cbt_leaf"][1:])
    assert o.potential_energy() == post.extra["epot"][0]
    assert o.kinetic_energy() == post.extra["ekin"][0]


@pytest.mark.parametrize("case", ["ala10", "ala10_hb", "gel10", "ala10x8", "ala10x8_wallp", "ala10x4_wallc", "ala10x8_obst", "abeta8", "ala10x4_tunnel", "ala10_flowv", "ala10_flowf", "ala10_charge"])
def test_event_loop_trace(case):
    pre, post, fin = load_case(case)
    hv, o = _prepare(pre)
    o.set_celllist(pre.extra["celllist"])
    if not int(pre.extra.get("warm", [0])[0]):
        o.predict_all()
    o.cbt_build()
    want = ref_trace(fin)
    st, tr = o.run(max_iter=len(want), trace_cap=len(want))
    assert st["error"] == 0
    assert_trace_equal(tr, want)
    assert_fields_equal(hv, fin, STATE_FIELDS, what="final state: ")
    assert np.array_equal(o.celllist(), fin.extra["celllist"])
    assert np.array_equal(o.cbt()[0][1:], fin.extra["cbt_node"][1:])
    ref = fin.extra["stats"]
    got = [st[k] for k in ("frame", "iterations", "count_recal", "collision", "bond", "hb", "hb_neighbor", "thermostat", "pcc")]
    got[0] = st["frame"]  # absolute frame counter
    assert got == [int(v) for v in ref[:9]]
    assert st["warnings"] == int(ref[9])              # warningsum (0 in every shipped fixture; see test_warningsum_against_live_reference)
    assert o.potential_energy() == fin.extra["epot"][0]
    if case == "ala10_hb":
        assert st["hb"] > 0 and st["hb_neighbor"] > 0  # the fixture really exercises the HB paths
    if "wall" in case:
        # smooth confining wall: WallTime runs in every prediction, WallEvent a few times
        assert pre.wall_kind in (1, 2) and st["wall"] == int(ref[11]) and st["wall"] > 0
        assert (want["type"] == 5).sum() >= st["wall"]
    if "charge" in case:
        # charge source + periodic image, -COM: ChargeTime / ChargeEvent (partner code -100 n), centre-of-mass term in ThermostatEvent
        assert pre.flow_mark == 3 and pre.charge_com == 1 and len(pre.charge_src) == 2
        assert ((want["type"] == 8) & (want["executed"] == 1)).sum() > 0
        assert o.kinetic_energy() == fin.extra["ekin"][0]
    if "flow" in case:
        # drift velocity / constant force: extra terms where the thermostat redraws a velocity, CalKinetE relative to the drift
        # (flowv) or to the chain's centre of mass (flowf), timeRec bookkeeping
        assert pre.flow_mark == (1 if case.endswith("v") else 2) and st["thermostat"] > 10
        assert o.kinetic_energy() == fin.extra["ekin"][0]
        assert_fields_equal(hv, fin, ("flow_time_rec",))
    if "tunnel" in case:
        # tunnel between two holed obstruction walls: TunnelTime / TunnelEvent, EnterTunnel in ObstTime
        assert len(pre.tunnel_diameter) == 1 and list(pre.obst_hole) == [1, 1] and st["wall"] == int(ref[11])
        assert ((want["type"] == 7) & (want["executed"] == 1)).sum() > 0
    if "obst" in case:
        # flat obstruction walls + hard spherical obstacles: both event types occur, partner codes -dim-10n / -1000n
        assert len(pre.obst_type) == 2 and len(pre.sphobst_radius) == 3 and st["wall"] == int(ref[11])
        assert ((want["type"] == 6) & (want["executed"] == 1)).sum() > 0 and ((want["type"] == 9) & (want["executed"] == 1)).sum() > 0


@pytest.mark.refbin
def test_long_run_against_live_reference(tmp_path):
    """300k iterations of ala10 regenerated with the reference binary on this machine."""
    if not (REFDIR / "ref_harness").exists():
        pytest.skip("oracle/_ref/ref_harness not built (needs /root/reference at build time)")
    from oracle.make_golden import run_case
    from sdmd_b200.hostview import HostView
    snaps = run_case("ala10_live", tmp_path, harness="ref_harness", coord="ala10.gro", max_iter=300000, n_gauss=600000,
                     extra=[], gel_n=0)
    pre, post, fin = (HostView.from_snapshot(p) for p in snaps)
    hv, o = _prepare(pre)
    o.set_celllist(pre.extra["celllist"])
    o.predict_all()
    o.cbt_build()
    want = ref_trace(fin)
    st, tr = o.run(max_iter=len(want), trace_cap=len(want))
    assert_trace_equal(tr, want)
    assert_fields_equal(hv, fin, STATE_FIELDS)
    assert st["hb"] > 0


@pytest.mark.refbin
def test_multichain_protein_against_live_reference(tmp_path):
    """80 ala10 chains (N = 4720, several proteins, inter-chain collisions) regenerated with the reference binary here."""
    if not (REFDIR / "ref_harness").exists():
        pytest.skip("oracle/_ref/ref_harness not built (needs /root/reference at build time)")
    from oracle.make_golden import run_case
    from sdmd_b200.hostview import HostView
    snaps = run_case("ala10x80", tmp_path, harness="ref_harness", coord="ala10x80.gro", max_iter=5000, n_gauss=40000,
                     extra=[], gel_n=-80)
    pre, post, fin = (HostView.from_snapshot(p) for p in snaps)
    assert pre.natom == 4720
    hv, o = _prepare(pre)
    o.pbc_full(); o.cell_build()
    assert np.array_equal(o.celllist(), pre.extra["celllist"])
    o.predict_all()
    assert_fields_equal(hv, post, SLOT_FIELDS, what="after Predict(all): ")
    o.cbt_build()
    want = ref_trace(fin)
    st, tr = o.run(max_iter=len(want), trace_cap=len(want))
    assert_trace_equal(tr, want)
    assert_fields_equal(hv, fin, STATE_FIELDS)
    assert o.potential_energy() == fin.extra["epot"][0]


@pytest.mark.refbin
def test_abeta64_box_against_live_reference(tmp_path):
    """BASELINE config 3: 64 x Abeta16-22 on a 4x4x4 grid, 25 A spacing, 100 A box (N = 4608; the chain's
    coordinates are the committed tests/golden/abeta16_22.gro), run by the reference binary here: link list, Predict(all),
    20 000 iterations."""
    if not (REFDIR / "ref_harness").exists():
        pytest.skip("oracle/_ref/ref_harness not built (needs /root/reference at build time)")
    from oracle.make_golden import run_case
    from sdmd_b200.hostview import HostView
    snaps = run_case("abeta64", tmp_path, harness="ref_harness", coord="abeta64.gro", max_iter=20000, n_gauss=120000,
                     extra=[], gel_n=-8)
    pre, post, fin = (HostView.from_snapshot(p) for p in snaps)
    assert pre.natom == 64 * 72 and len(set(pre.type[1:].tolist())) >= 14
    hv, o = _prepare(pre)
    o.pbc_full(); o.cell_build()
    assert np.array_equal(o.celllist(), pre.extra["celllist"])
    o.predict_all()
    assert_fields_equal(hv, post, SLOT_FIELDS, what="after Predict(all): ")
    o.cbt_build()
    want = ref_trace(fin)
    st, tr = o.run(max_iter=len(want), trace_cap=len(want))
    assert_trace_equal(tr, want)
    assert_fields_equal(hv, fin, STATE_FIELDS)
    assert o.potential_energy() == fin.extra["epot"][0]


@pytest.mark.refbin
def test_gel_42875_against_live_reference(tmp_path):
    """35^3 GEL atoms: the largest system the reference's int-sized (N+1)^2 connectionMap allows (N <= 46339), run here by the
    -DGEL reference binary; link list, Predict(all) and 5000 iterations."""
    if not (REFDIR / "ref_harness_gel").exists():
        pytest.skip("oracle/_ref/ref_harness_gel not built (needs /root/reference at build time)")
    from oracle.make_golden import run_case
    from sdmd_b200.hostview import HostView
    import shutil
    import tempfile
    from pathlib import Path
    work = Path(tempfile.mkdtemp(prefix="s", dir="/tmp"))   # short path: the reference keeps file names in small fixed buffers
    try:
        snaps = run_case("gel35", work, harness="ref_harness_gel", coord="GEL35.gro", max_iter=5000, n_gauss=80000,
                         extra=[], gel_n=35)
        pre, post, fin = (HostView.from_snapshot(p) for p in snaps)
    finally:
        shutil.rmtree(work, ignore_errors=True)
    assert pre.natom == 42875
    hv, o = _prepare(pre)
    o.pbc_full(); o.cell_build()
    assert np.array_equal(o.celllist(), pre.extra["celllist"])
    o.predict_all()
    assert_fields_equal(hv, post, SLOT_FIELDS, what="after Predict(all): ")
    o.cbt_build()
    want = ref_trace(fin)
    st, tr = o.run(max_iter=len(want), trace_cap=len(want))
    assert_trace_equal(tr, want)
    assert_fields_equal(hv, fin, STATE_FIELDS)


@pytest.mark.refbin
@pytest.mark.parametrize("n_axis", [20, 100])
def test_full_system_pass_against_the_reference_itself(n_axis):
    """The oracle port against the UNMODIFIED reference at the bench size (10^6 atoms; 8000 as the quick case): ref_bench_gel
    runs the -DGEL reference's own PBC + LinkList + Predict(all) + CreateCBT (dense connectionMap answered by a zero window,
    oracle/ref_bench.c); every event slot of the port's pass over the same start state must equal the reference's, bit for bit.
    This is what pins the port — bench.py's pair counter and the GPU tests' oracle — at the size the numbers are quoted on."""
    from tests import ref_bench
    if not ref_bench.available():
        pytest.skip("oracle/_ref/ref_bench_gel not built (needs /root/reference at build time)")
    hv, gauss, slots, _ = ref_bench.run(n_axis)
    o = Oracle(hv)
    o.rng_feed(gauss)
    o.pbc_full(); o.cell_build(); o.predict_all(); o.cbt_build()
    assert np.array_equal(hv.ev_time[1:].view(np.int64), slots["time"].view(np.int64))
    assert np.array_equal(hv.ev_potential[1:].view(np.int64), slots["potential"].view(np.int64))
    assert np.array_equal(hv.ev_partner[1:], slots["partner"])
    assert np.array_equal(hv.ev_type[1:].astype(np.int32), slots["type"]) and np.array_equal(hv.ev_sub[1:].astype(np.int32), slots["sub"])
    assert np.array_equal(hv.ev_pcounter[1:], slots["pcounter"])
    assert (slots["partner"] > 0).sum() > 0.1 * hv.natom


@pytest.mark.refbin
def test_warningsum_against_live_reference():
    """`warningsum` (DMD.c:61): an 8^3 GEL lattice at 3.95 A puts every nearest neighbour inside the 4.0 A hard core, so
    FindPair's too-small warning (ToolFunctions.c:119-139) fires for each of them in Predict(all) and again in every
    re-prediction; the oracle's `warnings` must equal the reference's count, and its negative-time / bond-error prints
    (`error_prints`, "!!ERROR!!" lines that never touch warningsum) stay out of it."""
    if not (REFDIR / "ref_harness_gel").exists():
        pytest.skip("oracle/_ref/ref_harness_gel not built (needs /root/reference at build time)")
    from oracle.make_golden import run_case
    from sdmd_b200.hostview import HostView
    import shutil
    import tempfile
    from pathlib import Path
    work = Path(tempfile.mkdtemp(prefix="s", dir="/tmp"))
    try:
        snaps = run_case("gel8w", work, harness="ref_harness_gel", coord="GEL8.gro", max_iter=2000, n_gauss=20000,
                         extra=[], gel_n=8, gel_a_nm=0.395)
        pre, post, fin = (HostView.from_snapshot(p) for p in snaps)
    finally:
        shutil.rmtree(work, ignore_errors=True)
    hv, o = _prepare(pre)
    o.pbc_full(); o.cell_build(); o.predict_all()
    assert_fields_equal(hv, post, SLOT_FIELDS, what="after Predict(all): ")
    assert o.stats()["warnings"] == 6 * 512
    assert o.potential_energy() == post.extra["epot"][0]   # the harness calls CalPotenE here too: 3 * 512 more warnings (ToolFunctions.c:684-752)
    o.cbt_build()
    want = ref_trace(fin)
    st, tr = o.run(max_iter=len(want), trace_cap=len(want))
    assert_trace_equal(tr, want)
    assert_fields_equal(hv, fin, STATE_FIELDS)
    assert st["warnings"] == int(fin.extra["stats"][9]) > 6 * 512


def test_remd_rearrange_rule():
    """Server.c:145-192: sequential, order dependent, pair (1,2) tried from both sides (hand-computed cases)."""
    # i=0 (num 1) meets num 2: dE=(2-1/.6)(-20+10)<0 -> swap.  i=1 (now num 1) meets j=0 (num 2): dE=+3.33 -> one
    # draw, 0.99 > exp(-3.33) -> reject.  i=2 (num 3) meets j=0 (num 2): dE=(1/.7-1/.6)(-10+30)<0 -> swap.
    nums, temps, used, att, acc = remd_rearrange([-10.0, -20.0, -30.0], [1, 2, 3], [0.5, 0.6, 0.7], [0.99] * 8)
    assert used == 1
    assert nums == [3, 1, 2] and temps == [0.7, 0.5, 0.6]
    assert att[:2] == [2.0, 1.0] and acc[:2] == [1.0, 1.0]
    # same start, but the draw is small enough to accept the reverse swap
    nums2, temps2, used2, _, _ = remd_rearrange([-10.0, -20.0], [1, 2], [0.5, 0.6], [0.01, 0.99])
    assert used2 == 1 and nums2 == [1, 2] and temps2 == [0.5, 0.6]
    # a finished replica (T == 0) is skipped entirely
    nums3, _, used3, _, _ = remd_rearrange([-10.0, 0.0, -30.0], [1, 2, 3], [0.5, 0.0, 0.7], [0.5])
    assert nums3 == [1, 2, 3] and used3 == 0


@pytest.mark.refbin
@pytest.mark.parametrize("case", ["ala10x8_wallp_compress", "ala10x8_wallp_pulse"])
def test_moving_wall_against_live_reference(tmp_path, case):
    """"WallDyn compress" / "WallDyn pulse" (8 ala10 chains between smooth plates that move): DoWallDyn (Events.c:502-540) once per
    time unit inside the loop, WallTime's retry (TimePrediction.c:675-697) when an atom ends up outside — event trace, final
    state, final box edges and wallDyn bookkeeping of the oracle against the reference harness run here."""
    if not (REFDIR / "ref_harness").exists():
        pytest.skip("oracle/_ref/ref_harness not built (needs /root/reference at build time)")
    from oracle.make_golden import run_case
    from sdmd_b200.hostview import HostView
    snaps = run_case(case, tmp_path, harness="ref_harness", coord="ala10x8w.gro", max_iter=40000, n_gauss=200000, extra=[], gel_n=-8)
    pre, post, fin = (HostView.from_snapshot(p) for p in snaps)
    assert pre.wall_dyn == 1 and pre.extra["wall_dyn_i4"][0] in (1, 3)
    hv, o = _prepare(pre)
    o.pbc_full(); o.cell_build()
    o.predict_all()
    assert_fields_equal(hv, post, SLOT_FIELDS, what="after Predict(all): ")
    o.cbt_build()
    want = ref_trace(fin)
    st, tr = o.run(max_iter=len(want), trace_cap=len(want))
    assert_trace_equal(tr, want)
    assert_fields_equal(hv, fin, STATE_FIELDS)
    wd = o.wall_dyn()
    wi, wf = fin.extra["wall_dyn_i4"], fin.extra["wall_dyn_f8"]
    assert bits_equal(np.array(wd["box"]), np.asarray(wf[7:10]))                # boxDimension after the run
    assert (wd["mode"], wd["sign"], wd["touch"]) == tuple(int(v) for v in wi[:3]) and wd["cur_time"] == float(wf[3])
    moved = abs(wf[8] - pre.extra["wall_dyn_f8"][8])
    # DoWallDyn ran several times, the wall was hit, and either it travelled or WallTime's retry pushed it back out (the chains fill
    # the slit, so a compression step of 1.25 A leaves atoms outside: every step is undone by a retry, touch counts them)
    assert wd["cur_time"] >= 3 and st["wall"] > 0 and (moved > 0.5 or wd["touch"] > 0), (wd, moved, st["wall"])


@pytest.mark.refbin
@pytest.mark.parametrize("case,kw", [
    ("ala10", dict(harness="ref_harness", coord="ala10.gro", max_iter=500, n_gauss=4000, extra=[], gel_n=0)),
    ("ala10x8", dict(harness="ref_harness", coord="ala10x8.gro", max_iter=500, n_gauss=8000, extra=[], gel_n=-8)),
    ("ala10x80", dict(harness="ref_harness", coord="ala10x80.gro", max_iter=500, n_gauss=40000, extra=[], gel_n=-80)),
], ids=["ala10", "ala10x8", "ala10x80"])
def test_grid_sdenergymin_equals_the_reference(tmp_path, case, kw, monkeypatch):
    """SDEnergyMin with the linear-time neighbour search (sdmd_b200/host/sdmd_sdmin.c) against the reference's own
    (ToolFunctions.c:325-491): the same harness built with either, same input — every dump (relaxed coordinates, the rand()-drawn
    thermostat stream that follows the relaxation's draws, slots, trace, final state) must be byte-identical, and so must the
    MiniEnergy.gro the function writes."""
    if not (REFDIR / "ref_harness_sdmin").exists():
        pytest.skip("oracle/_ref/ref_harness_sdmin not built (needs /root/reference at build time)")
    from oracle.make_golden import run_case
    monkeypatch.setenv("SDMD_FAST_SDMIN", "1")          # the grid search also below its automatic threshold of 8192 atoms
    want = run_case(case, tmp_path / "ref", **kw)
    got = run_case(case, tmp_path / "grid", **dict(kw, harness="ref_harness_sdmin"))
    for a, b in zip(want, got):
        assert a.read_bytes() == b.read_bytes(), a.name
    mini = [p.read_bytes() for p in (tmp_path / "ref" / "input" / "MiniEnergy.gro", tmp_path / "grid" / "input" / "MiniEnergy.gro")]
    assert len(mini[0]) > 1000 and mini[0] == mini[1]
