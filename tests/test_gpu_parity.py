"""GPU: the CUDA path, called through the C-ABI, against the reference's golden vectors and the oracle."""
import numpy as np
import pytest

from tests.util import (SLOT_FIELDS, STATE_FIELDS, assert_fields_equal, assert_trace_equal, bits_equal, load_case,
                        ref_trace, trace_times_close)

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def capi():
    from sdmd_b200 import capi as m
    m.build()
    return m


def _replica(capi, pre):
    hv = pre.copy()
    r = capi.Replica(hv, device=0)
    r.rng_feed(pre.extra["gauss"])
    return hv, r


@pytest.mark.parametrize("case", ["ala10", "gel10"])
def test_cell_build_matches_reference(capi, case):
    pre, post, fin = load_case(case)
    hv, r = _replica(capi, pre)
    assert np.array_equal(r.cellnum, pre.extra["cellnum"]) and bits_equal(r.cellsize, pre.extra["cellsize"])
    r.pbc_full()
    r.cell_build()
    assert np.array_equal(r.celllist(), pre.extra["celllist"])
    r.download(hv)
    assert_fields_equal(hv, pre, ("cell_idx", "x", "y", "z"))


@pytest.mark.parametrize("case", ["ala10", "gel10"])
def test_predict_all_and_cbt_match_reference(capi, case):
    pre, post, fin = load_case(case)
    hv, r = _replica(capi, pre)
    r.cell_build()
    r.predict_all()
    r.cbt_build()
    r.download(hv)
    assert_fields_equal(hv, post, SLOT_FIELDS, what="after Predict(all): ")
    node, leaf = r.cbt()
    assert np.array_equal(node[1:], post.extra["cbt_node"][1:])
    assert np.array_equal(leaf[1:], post.extra["cbt_leaf"][1:])
    assert r.potential_energy() == post.extra["epot"][0]
    assert r.kinetic_energy() == post.extra["ekin"][0]


@pytest.mark.parametrize("case", ["ala10", "ala10_hb", "gel10"])
def test_device_event_loop_matches_reference(capi, case):
    pre, post, fin = load_case(case)
    hv, r = _replica(capi, pre)
    if int(pre.extra.get("warm", [0])[0]):
        r.upload_celllist(pre.extra["celllist"])
    else:
        r.cell_build()
        r.predict_all()
    r.cbt_build()
    want = ref_trace(fin)
    st, tr = r.run(max_iter=len(want), trace_cap=len(want))
    assert st["error"] == 0 and st["iterations"] == len(want)
    assert trace_times_close(tr, want, rtol=1e-12)   # north_star tolerance ...
    assert_trace_equal(tr, want)                      # ... and in fact bit-equal
    r.download(hv)
    assert_fields_equal(hv, fin, STATE_FIELDS, what="final state: ")
    assert np.array_equal(r.celllist(), fin.extra["celllist"])
    assert np.array_equal(r.cbt()[0][1:], fin.extra["cbt_node"][1:])
    ref = [int(v) for v in fin.extra["stats"][:9]]
    got = [st[k] for k in ("frame", "iterations", "count_recal", "collision", "bond", "hb", "hb_neighbor", "thermostat", "pcc")]
    assert got == ref
    assert r.potential_energy() == fin.extra["epot"][0]


def test_stepwise_api_matches_run(capi):
    """Drive the loop one reference call at a time (AssignJob, UpdateData/TimeForward, HazardCheck, DoEvent,
    Predict, UpdateCBT) and compare with the golden trace."""
    pre, post, fin = load_case("ala10")
    hv, r = _replica(capi, pre)
    r.cell_build(); r.predict_all(); r.cbt_build()
    want = ref_trace(fin)[:300]
    for k, w in enumerate(want):
        target, partner, dt, renew = r.assign_job()
        assert (target, np.float64(dt).tobytes()) == (int(w["target"]), np.float64(w["dt"]).tobytes()), k
        r.advance(dt)
        r.add_time(dt)
        hz = r.hazard_check(target, partner)
        assert (hz == 0) == bool(w["executed"]), k
        if hz == 0:
            r.do_event()
            atoms = [target] + ([partner] if partner else [])
        else:
            atoms = [target]
        r.predict(atoms)
        r.cbt_update(atoms)


def test_large_gel_box_against_oracle(capi):
    """Synthetic GEL box (N = 32768, 8 A lattice) — CUDA vs the CPU oracle on identical inputs: cell list,
    full-system prediction (incl. the duplicate-copy fix-point), CBT, 2000 events through the multi-kernel loop."""
    from oracle.oracle_py import Oracle
    from sdmd_b200.synth import gel_box
    hv0, rng = gel_box(32, jitter=0.3)
    gauss = rng.standard_normal(hv0.natom + 20000)
    ho = hv0.copy(); o = Oracle(ho); o.rng_feed(gauss)
    hg = hv0.copy(); r = capi.Replica(hg, device=0); r.rng_feed(gauss)
    o.pbc_full(); o.cell_build(); o.predict_all(); o.cbt_build()
    r.pbc_full(); r.cell_build(); r.predict_all(); r.cbt_build()
    assert np.array_equal(r.celllist(), o.celllist())
    r.download(hg)
    assert_fields_equal(hg, ho, SLOT_FIELDS + ("cell_idx",))
    assert np.array_equal(r.cbt()[0][1:], o.cbt()[0][1:])
    assert r.stats()["pair_predictions"] == o.stats()["pair_predictions"]
    so, to = o.run(max_iter=2000, trace_cap=2000)
    sg, tg = r.run(max_iter=2000, trace_cap=2000)
    assert_trace_equal(tg, to)
    r.download(hg)
    assert_fields_equal(hg, ho, STATE_FIELDS)
    assert np.array_equal(r.celllist(), o.celllist())
    for k in ("frame", "count_recal", "collision", "thermostat", "pcc", "pair_predictions", "gaussians_used"):
        assert sg[k] == so[k], k


def test_round_trip_properties_1m(capi):
    """Full size (N = 10^6): size-independent properties instead of an element-wise oracle comparison —
    link list is a permutation with every atom in the cell its coordinates say, CBT root is the global
    argmin, both atoms of a predicted pair carry the bit-identical time, advance(dt) then advance(-dt)..."""
    from sdmd_b200.synth import gel_box
    hv, rng = gel_box(100, jitter=0.3)
    n = hv.natom
    r = capi.Replica(hv, device=0)
    r.rng_feed(rng.standard_normal(n))
    r.pbc_full(); r.cell_build(); r.predict_all(); r.cbt_build()
    r.download(hv)
    cl = r.celllist()
    # every atom appears exactly once as a successor/head
    succ = np.concatenate([cl[1:n + 1], cl[n + 1:]])
    seen = succ[succ > 0]
    assert len(seen) == n and len(np.unique(seen)) == n
    # cell index consistent with coordinates
    ci = hv.cell_idx.reshape(-1, 3)[1:]
    for k, arr in enumerate((hv.x, hv.y, hv.z)):
        want = np.minimum((arr[1:] / r.cellsize[k]).astype(np.int32), r.cellnum[k] - 1)
        assert np.array_equal(ci[:, k], want)
    # heads: descending order inside each cell (fresh list)
    heads = cl[n + 1:]
    nxt = cl[heads[heads > 0]]
    assert np.all((nxt == 0) | (nxt < heads[heads > 0]))
    # tournament root = earliest event, ties resolved to SOME atom with the minimum time
    node, leaf = r.cbt()
    t = hv.ev_time
    assert t[node[1]] == t[1:].min()
    # pair symmetry: if j's slot points at i and i's at j with the same type, times are bit-identical
    p = hv.ev_partner
    i = np.arange(n + 1)
    mutual = (p > 0)
    mutual[mutual] &= (p[p[mutual]] == i[mutual])
    assert mutual.sum() > 1000
    assert np.array_equal(t[mutual].view(np.int64), t[p[mutual]].view(np.int64))
    assert np.all(t[1:] > 0)


def test_remd_ensemble_gpu_matches_oracle(capi):
    """Three ala10 replicas on one GPU coupled by the Server.c sweep: exchange log and final states equal the same
    ensemble run on the CPU oracle."""
    from oracle.oracle_py import Oracle
    from sdmd_b200.remd import RemdEnsemble, reduced_temperature
    pre, _, _ = load_case("ala10")
    temps = [reduced_temperature(t) for t in (236.5, 250, 263.5)]

    def build(kind):
        reps, views = [], []
        for k in range(3):
            hv = pre.copy(); hv.timestep = 1e12
            g = np.random.default_rng(50 + k).standard_normal(150000)
            r = Oracle(hv) if kind == "cpu" else capi.Replica(hv, device=0)
            r.rng_feed(g)
            if kind == "cpu":
                r.set_celllist(pre.extra["celllist"])
            else:
                r.upload_celllist(pre.extra["celllist"])
            r.predict_all() if kind == "cpu" else (r.cell_build(), r.predict_all())
            r.cbt_build()
            reps.append(r); views.append(hv)
        return RemdEnsemble(reps, temps, 3.0, 2024), views

    cpu, cpu_views = build("cpu")
    gpu, gpu_views = build("gpu")
    for _ in range(10):
        ec = cpu.step()
        eg = gpu.step()
        assert ec == eg   # CalPotenE bit-equal on every replica at every exchange
    for k in range(3):
        assert cpu.log[k] == gpu.log[k]
        gpu.reps[k].download(gpu_views[k])
        assert_fields_equal(gpu_views[k], cpu_views[k], STATE_FIELDS, what=f"replica {k}: ")
    assert any(rec[2] != k + 1 for k in range(3) for rec in gpu.log[k])
