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


@pytest.mark.parametrize("case", ["ala10", "gel10", "ala10x8", "ala10x8_wallp"])
def test_cell_build_matches_reference(capi, case):
    pre, post, fin = load_case(case)
    hv, r = _replica(capi, pre)
    assert np.array_equal(r.cellnum, pre.extra["cellnum"]) and bits_equal(r.cellsize, pre.extra["cellsize"])
    r.pbc_full()
    r.cell_build()
    assert np.array_equal(r.celllist(), pre.extra["celllist"])
    r.download(hv)
    assert_fields_equal(hv, pre, ("cell_idx", "x", "y", "z"))


@pytest.mark.parametrize("case", ["ala10", "gel10", "ala10x8", "ala10x8_wallp"])
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


@pytest.mark.parametrize("case", ["ala10", "ala10_hb", "gel10", "ala10x8", "ala10x8_wallp", "ala10x4_wallc", "ala10x8_obst", "abeta8", "ala10x4_tunnel", "ala10_flowv", "ala10_flowf", "ala10_charge"])
@pytest.mark.parametrize("loop_mode", [1, 4, 2, 5, 3], ids=["one_cta", "cluster", "sweep_persistent", "sweep_per_launch", "sweep_serial"])
def test_device_event_loop_matches_reference(capi, case, loop_mode):
    """loop_mode 4 forces the mid-size driver (one thread-block cluster) and 2 / 3 force the large-system drivers (grid-wide sweep overlapped with / followed by one event block per
    iteration) onto the protein cases too, so every loop driver is pinned to the reference trace with bonds, HB and
    HB-neighbour events."""
    pre, post, fin = load_case(case)
    hv, r = _replica(capi, pre)
    r.set_loop_mode(loop_mode)
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
    assert st["warnings"] == int(fin.extra["stats"][9])   # warningsum
    assert r.potential_energy() == fin.extra["epot"][0]
    if "charge" in case:   # charge sources: ChargeTime / ChargeEvent, centre-of-mass term of ThermostatEvent (-COM)
        assert r.kinetic_energy() == fin.extra["ekin"][0] and (tr["type"] == 8).sum() > 0
    if "flow" in case:   # drift velocity / constant force terms of ThermostatEvent and CalKinetE
        assert r.kinetic_energy() == fin.extra["ekin"][0]
        assert_fields_equal(hv, fin, ("flow_time_rec",))
    if "wall" in case or "obst" in case or "tunnel" in case:   # smooth walls / obstruction walls / spherical obstacles
        assert st["wall"] == int(fin.extra["stats"][11]) and st["wall"] > 0


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
    assert np.float64(r.potential_energy()).tobytes() == np.float64(o.potential_energy()).tobytes()
    assert np.float64(r.kinetic_energy()).tobytes() == np.float64(o.kinetic_energy()).tobytes()


@pytest.mark.parametrize("n_axis,iters", [(16, 3000), (40, 1500)])
def test_walled_gel_box_against_oracle(capi, n_axis, iters):
    """GEL box between two smooth parallel plates (outermost lattice planes inside the wall's potential step): the wall
    predictor inside the full-system pass (K2 incl. the duplicate fix-point), the one-block loop (N = 4096) and the
    sweep loop (N = 64000), PBC in x and z only — CUDA vs the CPU oracle, bit for bit."""
    from oracle.oracle_py import Oracle
    from sdmd_b200.synth import gel_box
    hv0, rng = gel_box(n_axis, jitter=0.3, wall_pad=1.2)
    gauss = rng.standard_normal(hv0.natom + 20 * iters)
    ho = hv0.copy(); o = Oracle(ho); o.rng_feed(gauss)
    hg = hv0.copy(); r = capi.Replica(hg, device=0); r.rng_feed(gauss)
    o.pbc_full(); o.cell_build(); o.predict_all(); o.cbt_build()
    r.pbc_full(); r.cell_build(); r.predict_all(); r.cbt_build()
    assert np.array_equal(r.celllist(), o.celllist())
    r.download(hg)
    assert_fields_equal(hg, ho, SLOT_FIELDS + ("cell_idx",))
    assert (ho.ev_type[1:] == 5).sum() > 0          # wall events are scheduled from the start
    assert r.stats()["pair_predictions"] == o.stats()["pair_predictions"]
    so, to = o.run(max_iter=iters, trace_cap=iters)
    sg, tg = r.run(max_iter=iters, trace_cap=iters)
    assert so["error"] == 0 and sg["error"] == 0
    assert_trace_equal(tg, to)
    r.download(hg)
    assert_fields_equal(hg, ho, STATE_FIELDS)
    assert np.array_equal(r.celllist(), o.celllist())
    for k in ("frame", "count_recal", "collision", "thermostat", "pcc", "wall", "pair_predictions", "gaussians_used"):
        assert sg[k] == so[k], k
    if n_axis == 16:
        assert so["wall"] > 0   # (at N = 64000 the first 1500 events are all thermostat / cell crossings)


def _tie_box():
    """8^3 GEL lattice (no jitter) with planted exact ties: a centre atom C and its two lattice neighbours along one axis
    approach it with mirrored velocities (all values dyadic, so C's two candidate times are the same bits).  The
    reference keeps the candidate it visits first."""
    from sdmd_b200.synth import gel_box
    hv, rng = gel_box(8, jitter=0.0)
    n = 8
    planted = []
    for axis, (cx, cy, cz) in ((0, (3, 2, 2)), (1, (5, 3, 6)), (2, (2, 6, 4)), (0, (6, 6, 1)), (1, (1, 4, 3)), (2, (4, 1, 5)),
                               (0, (1, 1, 6)), (2, (6, 3, 3))):
        idx = lambda x, y, z: 1 + x + n * y + n * n * z
        c = idx(cx, cy, cz)
        d = [0, 0, 0]; d[axis] = 1
        a = idx(cx - d[0], cy - d[1], cz - d[2]); b = idx(cx + d[0], cy + d[1], cz + d[2])
        base = [0.25, 0.125, -0.375]
        for atom, sgn in ((c, 0), (a, +1), (b, -1)):
            v = list(base); v[axis] += sgn * 2.0
            hv.vx[atom], hv.vy[atom], hv.vz[atom] = v
        planted.append((c, a, b))
    return hv, rng, planted


def test_exact_ties_keep_reference_visiting_order(capi):
    """Exact ties between two candidates of one atom must go to the one the reference visits first (strict '<', first
    wins, TimePrediction.c:1142) in the full-system pass, its resolve / re-evaluation rounds and the event loop."""
    from oracle.oracle_py import Oracle
    hv0, rng, planted = _tie_box()
    gauss = rng.standard_normal(hv0.natom + 40000)
    ho = hv0.copy(); o = Oracle(ho); o.rng_feed(gauss)
    hg = hv0.copy(); r = capi.Replica(hg, device=0); r.rng_feed(gauss)
    o.pbc_full(); o.cell_build(); o.predict_all(); o.cbt_build()
    r.pbc_full(); r.cell_build(); r.predict_all(); r.cbt_build()
    r.download(hg)
    assert_fields_equal(hg, ho, SLOT_FIELDS + ("cell_idx",))
    won = [c for c, a, b in planted if ho.ev_partner[c] in (a, b)]
    assert len(won) >= 4, won          # the planted pairs really are the earliest events of their centre atoms
    assert r.stats()["pair_predictions"] == o.stats()["pair_predictions"]
    so, to = o.run(max_iter=1500, trace_cap=1500)
    sg, tg = r.run(max_iter=1500, trace_cap=1500)
    assert_trace_equal(tg, to)
    r.download(hg)
    assert_fields_equal(hg, ho, STATE_FIELDS)


@pytest.mark.parametrize("variant", ["cell_by_cell_kernels", "two_types", "three_step_table", "ties_cell_by_cell"])
def test_round0_paths_against_oracle(capi, variant, monkeypatch):
    """Round 0 of the full-system pass has several code paths for bond-free boxes; each one must give the oracle's bits:
    the row-batched filter + cp.async evaluation of sdmd_k2_rows.cu with the step table in registers (default, every other GEL
    test), its generic table walk (two atom types / a table of three steps: DevP::uni_table = -1), and the cell-by-cell filter +
    ordered evaluation of sdmd_cuda.cu (SDMD_K2_ROWS=0)."""
    from oracle.oracle_py import Oracle
    from sdmd_b200.hostview import NT, STEP_TABLE_DTYPE
    from sdmd_b200.synth import GEL_TYPE, gel_box
    if variant in ("cell_by_cell_kernels", "ties_cell_by_cell"):
        monkeypatch.setenv("SDMD_K2_ROWS", "0")
    if variant == "ties_cell_by_cell":
        hv0, rng, _ = _tie_box()
    else:
        hv0, rng = gel_box(32, a=6.5, jitter=0.4, seed=77)
    n = hv0.natom
    if variant == "two_types":
        other = 31
        t = np.zeros(3, dtype=STEP_TABLE_DTYPE)
        t[0] = hv0.tables[0]
        for k, (core, d, e) in enumerate([(3.6, 6.0, -0.3), (3.8, 6.2, 0.25)], start=1):
            t[k]["dmin2"] = core * core; t[k]["nstep"] = 1; t[k]["d2"][0] = d * d; t[k]["e"][0] = e; t[k]["acc"][0] = e
        hv0.tables = t
        hv0.type[1::3] = other
        hv0.coll_index[other * NT + other] = 1
        hv0.coll_index[other * NT + GEL_TYPE] = 2
        hv0.coll_index[GEL_TYPE * NT + other] = 2
    if variant == "three_step_table":
        t = hv0.tables.copy()
        t[0]["nstep"] = 3
        for k, (d, e) in enumerate([(5.0, 0.4), (6.0, -0.2), (7.0, 0.1)]):
            t[0]["d2"][k] = d * d; t[0]["e"][k] = e
        t[0]["acc"][0] = 0.4; t[0]["acc"][1] = 0.2; t[0]["acc"][2] = 0.3
        hv0.tables = t
    gauss = rng.standard_normal(n + 20000)
    ho = hv0.copy(); o = Oracle(ho); o.rng_feed(gauss)
    hg = hv0.copy(); r = capi.Replica(hg, device=0); r.rng_feed(gauss)
    o.pbc_full(); o.cell_build(); o.predict_all(); o.cbt_build()
    r.pbc_full(); r.cell_build(); r.predict_all(); r.cbt_build()
    r.download(hg)
    assert_fields_equal(hg, ho, SLOT_FIELDS + ("cell_idx",), what=variant + " after Predict(all): ")
    assert r.stats()["pair_predictions"] == o.stats()["pair_predictions"]
    assert r.stats()["warnings"] == o.stats()["warnings"]
    so, to = o.run(max_iter=800, trace_cap=800)
    sg, tg = r.run(max_iter=800, trace_cap=800)
    assert_trace_equal(tg, to)
    r.close(); o.close()


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


@pytest.mark.parametrize("n_axis,a,jitter,step_e", [
    (100, 8.0, 0.3, 0.2),      # THE bench box (bench.py pass1m / loop: gel_box(100, a=8.0, jitter=0.3), default seed)
    (100, 8.0, 0.3, -0.2),     # same box, attractive well (e = -0.2)
    (80, 5.0, 0.3, 0.2),       # C5's second density at 512 000 atoms: ~6 atoms per cell, ~160 candidates, ~18 in the scan sphere
    (100, 5.0, 0.3, -0.2),     # ... and at 10^6 atoms with the well
    (72, 4.4, 0.15, 0.2),      # denser still (373 248 atoms, ~28 in the scan sphere): survivor lists past their capacity at scale
], ids=["bench_1m_a8", "bench_1m_a8_well", "dense_512k_a5", "dense_1m_a5_well", "overflow_373k_a4.4"])
def test_full_size_pass_elementwise_against_oracle(capi, n_axis, a, jitter, step_e):
    """The configurations the bench numbers are quoted on, element by element against the CPU oracle on identical inputs:
    PBC + link list (reference byte layout), every event slot after Predict(all) (time / potential bits, partner, type,
    sub-type, partnerCounter), the whole tournament tree, and the pair-evaluation count (TimePrediction.c:105)."""
    from oracle.oracle_py import Oracle
    from sdmd_b200.synth import gel_box
    hv0, rng = gel_box(n_axis, a=a, jitter=jitter, step_e=step_e)
    n = hv0.natom
    gauss = rng.standard_normal(n)
    ho = hv0.copy(); o = Oracle(ho); o.rng_feed(gauss)
    hg = hv0.copy(); r = capi.Replica(hg, device=0); r.rng_feed(gauss)
    o.pbc_full(); o.cell_build(); o.predict_all(); o.cbt_build()
    r.pbc_full(); r.cell_build(); r.predict_all(); r.cbt_build()
    assert np.array_equal(r.celllist(), o.celllist())
    r.download(hg)
    assert_fields_equal(hg, ho, SLOT_FIELDS + ("cell_idx", "x", "y", "z"), what="after Predict(all): ")
    gn, gl = r.cbt()
    on, ol = o.cbt()
    assert np.array_equal(gn[1:], on[1:]) and np.array_equal(gl[1:], ol[1:])
    sg, so = r.stats(), o.stats()
    assert sg["pair_predictions"] == so["pair_predictions"] and so["pair_predictions"] > 20 * n
    assert sg["gaussians_used"] == so["gaussians_used"] == n
    assert sg["warnings"] == so["warnings"]
    if n_axis == 100 and a == 8.0 and step_e == 0.2:
        assert so["pair_predictions"] == 41027204      # the count bench.py's two arms report for this box
    r.close(); o.close()


def test_full_size_pass_against_the_reference_itself(capi):
    """10^6 atoms against the UNMODIFIED reference, not its restatement: oracle/_ref/ref_bench_gel runs PBC + LinkList +
    Predict(all) + CreateCBT of the -DGEL build on a jittered 100^3 lattice (its (N+1)^2 connectionMap answered by a zero window,
    oracle/ref_bench.c) and dumps what it started from; the CUDA path runs the same pass on that state and must reproduce every
    event slot the reference wrote — time and potential bits, partner, type, sub-type, partnerCounter."""
    from tests import ref_bench
    if not ref_bench.available():
        pytest.skip("oracle/_ref/ref_bench_gel not built")
    hv, gauss, slots, _ = ref_bench.run(100)
    n = hv.natom
    assert n == 1000000 and len(slots) == n
    r = capi.Replica(hv, device=0)
    r.rng_feed(gauss)
    r.pbc_full(); r.cell_build(); r.predict_all(); r.cbt_build()
    r.download(hv)
    assert np.array_equal(hv.ev_time[1:].view(np.int64), slots["time"].view(np.int64))
    assert np.array_equal(hv.ev_potential[1:].view(np.int64), slots["potential"].view(np.int64))
    assert np.array_equal(hv.ev_partner[1:], slots["partner"])
    assert np.array_equal(hv.ev_type[1:].astype(np.int32), slots["type"]) and np.array_equal(hv.ev_sub[1:].astype(np.int32), slots["sub"])
    assert np.array_equal(hv.ev_pcounter[1:], slots["pcounter"])
    assert (slots["partner"] > 0).sum() > 100000 and r.stats()["pair_predictions"] > 40 * n
    r.close()


def test_remd_exchange_with_fused_energy(capi):
    """sdmd_remd_exchange_energy: CalPotenE computed on the device and written straight into the exchange record; with one rank
    the sweep leaves (num, T) alone and the energy that went in equals sdmd_cuda_potential_energy bit for bit.  An unseeded
    server stream is an error, not garbage."""
    pre, _, _ = load_case("ala10")
    hv, r = _replica(capi, pre)
    r.cell_build(); r.predict_all(); r.cbt_build()
    r.run(max_iter=2000)
    with pytest.raises(capi.SdmdError):
        r.remd_exchange_energy(1, hv.target_temperature)
    r.remd_seed_server(99)
    num, T, e = r.remd_exchange_energy(1, hv.target_temperature)
    assert (num, T) == (1, hv.target_temperature)
    assert np.float64(e).tobytes() == np.float64(r.potential_energy()).tobytes()
    assert r.remd_exchange(e, 1, hv.target_temperature) == (1, hv.target_temperature)
    r.close()


@pytest.mark.parametrize("threads", [256, 128])
def test_remd_ensemble_gpu_matches_oracle(capi, threads, monkeypatch):
    """Three ala10 replicas on one GPU coupled by the Server.c sweep: exchange log and final states equal the same
    ensemble run on the CPU oracle.  Both block widths of the many-replica kernel (128 threads is what more than two
    replicas per SM get) are forced."""
    monkeypatch.setenv("SDMD_MULTI_THREADS", str(threads))
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


def _oracle_vs_gpu(capi, hv0, gauss, iters, check_pairs=True):
    from oracle.oracle_py import Oracle
    ho = hv0.copy(); o = Oracle(ho); o.rng_feed(gauss)
    hg = hv0.copy(); r = capi.Replica(hg, device=0); r.rng_feed(gauss)
    o.pbc_full(); o.cell_build(); o.predict_all(); o.cbt_build()
    r.pbc_full(); r.cell_build(); r.predict_all(); r.cbt_build()
    assert np.array_equal(r.celllist(), o.celllist())
    r.download(hg)
    assert_fields_equal(hg, ho, SLOT_FIELDS + ("cell_idx",), what="after Predict(all): ")
    assert np.array_equal(r.cbt()[0][1:], o.cbt()[0][1:])
    if check_pairs:
        assert r.stats()["pair_predictions"] == o.stats()["pair_predictions"]
    so, to = o.run(max_iter=iters, trace_cap=iters)
    sg, tg = r.run(max_iter=iters, trace_cap=iters)
    assert so["stop_reason"] == sg["stop_reason"] == 2, (so, sg)
    assert_trace_equal(tg, to)
    r.download(hg)
    assert_fields_equal(hg, ho, STATE_FIELDS, what="final state: ")
    assert np.array_equal(r.celllist(), o.celllist())
    assert np.array_equal(r.cbt()[0][1:], o.cbt()[0][1:])
    for k in ("frame", "count_recal", "collision", "thermostat", "pcc", "pair_predictions", "gaussians_used"):
        assert sg[k] == so[k], k
    # CalPotenE / CalKinetE: ordered sums, bit-equal (large systems take the parallel addend-list path)
    assert np.float64(r.potential_energy()).tobytes() == np.float64(o.potential_energy()).tobytes()
    assert np.float64(r.kinetic_energy()).tobytes() == np.float64(o.kinetic_energy()).tobytes()
    r.close()
    return so


@pytest.mark.parametrize("n_axis,a,iters", [
    (3, 9.0, 3000),      # 27 atoms, 3x3x3 cells: every neighbour cell is a periodic image (cellnum == 3 wrap rules)
    (4, 6.0, 3000),      # box 24 A -> cellnum clamps to 3 with cells larger than the cut-off
    (12, 5.0, 4000),     # dense: ~6 atoms per cell, ~160 candidates per atom, many collisions and PCC events
    (26, 5.0, 1500),     # dense and N > 8192: 128-entry cache rows overflow -> serial re-prediction fallback
])
def test_gel_edge_shapes_against_oracle(capi, n_axis, a, iters):
    from sdmd_b200.synth import gel_box
    hv0, rng = gel_box(n_axis, a=a, jitter=0.35, temperature=1.5)
    gauss = rng.standard_normal(hv0.natom + 6 * iters + 1000)
    so = _oracle_vs_gpu(capi, hv0, gauss, iters)
    if n_axis == 12:
        assert so["collision"] > 50


def test_square_well_and_no_thermostat(capi):
    """Attractive well (step e = -0.2, private copy of the GEL row) with the thermostat switched off: only
    collision / capture / escape and cell-crossing events; no random stream is consumed."""
    from sdmd_b200.synth import gel_box
    hv0, rng = gel_box(10, a=5.5, jitter=0.3, temperature=1.0, step_e=-0.2)
    hv0.thermostat_on = 0
    so = _oracle_vs_gpu(capi, hv0, np.zeros(0), 4000)
    assert so["thermostat"] == 0 and so["gaussians_used"] == 0 and so["collision"] > 500


def test_run_stops_at_time_boundaries(capi):
    """t_until / timestep semantics of the loop (SGThread.c:75,132): same stop points as the oracle."""
    from oracle.oracle_py import Oracle
    pre, _, _ = load_case("ala10")
    ho = pre.copy(); ho.timestep = 3.0
    hg = ho.copy()
    o = Oracle(ho); o.rng_feed(pre.extra["gauss"])
    r = capi.Replica(hg, device=0); r.rng_feed(pre.extra["gauss"])
    o.set_celllist(pre.extra["celllist"]); o.predict_all(); o.cbt_build()
    r.cell_build(); r.predict_all(); r.cbt_build()
    for t_until in (0.5, 1.0, 2.5, 10.0):
        so, _ = o.run(t_until=t_until)
        sg, _ = r.run(t_until=t_until)
        for k in ("frame", "iterations", "stop_reason"):
            assert so[k] == sg[k], (t_until, k, so, sg)
        assert np.float64(so["currenttime"]).tobytes() == np.float64(sg["currenttime"]).tobytes()
    assert sg["stop_reason"] == 0   # currenttime > timestep ended the run


def test_predict_arbitrary_renew_list(capi):
    """Predict(renewList) with a longer list (what FixAssignRenewList / CGResidueRenew build): reset all, then in order."""
    from oracle.oracle_py import Oracle
    pre, _, _ = load_case("ala10")
    ho, hg = pre.copy(), pre.copy()
    o = Oracle(ho); o.rng_feed(pre.extra["gauss"])
    r = capi.Replica(hg, device=0); r.rng_feed(pre.extra["gauss"])
    o.set_celllist(pre.extra["celllist"]); o.predict_all(); o.cbt_build()
    r.cell_build(); r.predict_all(); r.cbt_build()
    atoms = [7, 3, 30, 31, 12, 55]
    o.predict(atoms); o.cbt_update(atoms)
    r.predict(atoms); r.cbt_update(atoms)
    r.download(hg)
    assert_fields_equal(hg, ho, SLOT_FIELDS)
    assert np.array_equal(r.cbt()[0][1:], o.cbt()[0][1:])


@pytest.mark.parametrize("case", ["ala10", "gel10"])
def test_prepare_one_call_equals_stepwise(capi, case):
    """sdmd_cuda_prepare (overlapped upload) == upload_coords + rng + pbc_full + cell_build + predict_all + cbt_build."""
    pre, post, _ = load_case(case)
    ha, hb = pre.copy(), pre.copy()
    a = capi.Replica(ha, device=0); b = capi.Replica(hb, device=0)
    g = pre.extra["gauss"]
    a.upload_coords(); a.rng_refeed(g); a.pbc_full(); a.cell_build(); a.predict_all(); a.cbt_build()
    for _ in range(2):                       # twice: the second call replaces live state
        b.prepare(None, g)
    a.download(ha); b.download(hb)
    assert_fields_equal(hb, ha, tuple(SLOT_FIELDS) + ("x", "y", "z", "vx", "vy", "vz"))
    assert_fields_equal(hb, post, SLOT_FIELDS)
    assert np.array_equal(a.celllist(), b.celllist())
    assert np.array_equal(a.cbt()[0][1:], b.cbt()[0][1:])
    sa, sb = a.run(max_iter=300, trace_cap=300), b.run(max_iter=300, trace_cap=300)
    assert sa[1].tobytes() == sb[1].tobytes()
    # ... and Prepare() on the resident state (sdmd_cuda_prepare_resident) == the four stage calls, from where the run stopped
    pa0, pb0 = a.stats()["pair_predictions"], b.stats()["pair_predictions"]
    a.pbc_full(); a.cell_build(); a.predict_all(); a.cbt_build()
    b.prepare_resident()
    a.download(ha); b.download(hb)
    assert_fields_equal(hb, ha, tuple(SLOT_FIELDS) + ("x", "y", "z", "vx", "vy", "vz", "cell_idx"))
    assert np.array_equal(a.celllist(), b.celllist())
    assert np.array_equal(a.cbt()[0][1:], b.cbt()[0][1:])
    assert a.stats()["pair_predictions"] - pa0 == b.stats()["pair_predictions"] - pb0 > 0


def _live_case(name, **kw):
    """Run the unmodified reference harness now (short work directory: the reference keeps file names in small fixed
    buffers) and load its three snapshots."""
    import shutil
    import tempfile
    from pathlib import Path
    from oracle.make_golden import run_case
    work = Path(tempfile.mkdtemp(prefix="s", dir="/tmp"))
    try:
        run_case(name, work, **kw)
        return load_case(name, src=work, ext="snap")
    finally:
        shutil.rmtree(work, ignore_errors=True)


def test_multichain_protein_past_one_cta_limit_live(capi):
    """SURVEY C3-like case generated on the spot by the UNMODIFIED reference (oracle/_ref/ref_harness travels with the
    repository): 80 ala10 chains, N = 4720, with bonds, several proteins, HB candidates.  Cell list, Predict(all) slots,
    CBT, energies and a 20 000-iteration trace must equal the reference bit for bit on the default driver (one persistent
    thread block up to 8192 atoms) and on both large-system sweep drivers."""
    from oracle.make_golden import REFDIR as ref_dir
    if not (ref_dir / "ref_harness").exists():
        pytest.skip("oracle/_ref/ref_harness not built")
    n_it = 20000
    pre, post, fin = _live_case("ala10x80", harness="ref_harness", coord="ala10x80.gro", max_iter=n_it, n_gauss=120000, extra=[], gel_n=-80)
    assert pre.natom == 4720
    want = ref_trace(fin)
    assert len(want) == n_it
    for mode in (0, 1, 2, 3):
        hv, r = _replica(capi, pre)
        r.set_loop_mode(mode)
        r.pbc_full(); r.cell_build()
        assert np.array_equal(r.celllist(), pre.extra["celllist"])
        r.predict_all(); r.cbt_build()
        r.download(hv)
        assert_fields_equal(hv, post, SLOT_FIELDS, what="after Predict(all): ")
        assert np.array_equal(r.cbt()[0][1:], post.extra["cbt_node"][1:])
        assert r.potential_energy() == post.extra["epot"][0]
        st, tr = r.run(max_iter=n_it, trace_cap=n_it)
        assert st["error"] == 0 and st["iterations"] == n_it
        assert_trace_equal(tr, want)
        r.download(hv)
        assert_fields_equal(hv, fin, STATE_FIELDS, what="final state: ")
        assert np.array_equal(r.celllist(), fin.extra["celllist"])
        assert np.array_equal(r.cbt()[0][1:], fin.extra["cbt_node"][1:])
        assert r.potential_energy() == fin.extra["epot"][0]
        r.close()


def test_abeta64_box_live(capi):
    """BASELINE config 3 generated on the spot by the UNMODIFIED reference: 64 x Abeta16-22 (KLVFFAE, 16 atom
    types, CONSTRn tables, charged and aromatic side chains) on a 4x4x4 grid in a 100 A box, N = 4608.  Cell list,
    Predict(all) slots, CBT, energies and a 20 000-iteration trace bit for bit, on all three loop drivers."""
    from oracle.make_golden import REFDIR as ref_dir
    if not (ref_dir / "ref_harness").exists():
        pytest.skip("oracle/_ref/ref_harness not built")
    n_it = 20000
    pre, post, fin = _live_case("abeta64", harness="ref_harness", coord="abeta64.gro", max_iter=n_it, n_gauss=120000, extra=[], gel_n=-8)
    assert pre.natom == 64 * 72
    want = ref_trace(fin)
    assert len(want) == n_it
    for mode in (0, 1, 2, 3):
        hv, r = _replica(capi, pre)
        r.set_loop_mode(mode)
        r.pbc_full(); r.cell_build()
        assert np.array_equal(r.celllist(), pre.extra["celllist"])
        r.predict_all(); r.cbt_build()
        r.download(hv)
        assert_fields_equal(hv, post, SLOT_FIELDS, what="after Predict(all): ")
        assert np.array_equal(r.cbt()[0][1:], post.extra["cbt_node"][1:])
        assert r.potential_energy() == post.extra["epot"][0]
        st, tr = r.run(max_iter=n_it, trace_cap=n_it)
        assert st["error"] == 0 and st["iterations"] == n_it
        assert_trace_equal(tr, want)
        r.download(hv)
        assert_fields_equal(hv, fin, STATE_FIELDS, what="final state: ")
        assert np.array_equal(r.celllist(), fin.extra["celllist"])
        assert r.potential_energy() == fin.extra["epot"][0]
        r.close()


def test_gel_8000_live_reference(capi):
    """20^3 GEL lattice (the reference's own CreateGELCoordinate layout) run by the UNMODIFIED -DGEL reference on the spot:
    N = 8000 takes the overlapped-sweep driver; full-system pass and a 10 000-iteration trace bit for bit."""
    from oracle.make_golden import REFDIR as ref_dir
    if not (ref_dir / "ref_harness_gel").exists():
        pytest.skip("oracle/_ref/ref_harness_gel not built")
    n_it = 10000
    pre, post, fin = _live_case("gel20", harness="ref_harness_gel", coord="GEL20.gro", max_iter=n_it, n_gauss=60000, extra=[], gel_n=20)
    assert pre.natom == 8000
    hv, r = _replica(capi, pre)
    r.pbc_full(); r.cell_build()
    assert np.array_equal(r.celllist(), pre.extra["celllist"])
    r.predict_all(); r.cbt_build()
    r.download(hv)
    assert_fields_equal(hv, post, SLOT_FIELDS, what="after Predict(all): ")
    want = ref_trace(fin)
    st, tr = r.run(max_iter=n_it, trace_cap=n_it)
    assert st["error"] == 0 and st["iterations"] == n_it
    assert_trace_equal(tr, want)
    r.download(hv)
    assert_fields_equal(hv, fin, STATE_FIELDS, what="final state: ")
    assert np.array_equal(r.celllist(), fin.extra["celllist"])
    assert r.potential_energy() == fin.extra["epot"][0]


def test_gel_42875_live_reference(capi):
    """The largest box the UNMODIFIED reference can hold: its (N+1)^2 connectionMap is sized in `int` arithmetic
    (Initialization.c:976), so N <= 46339; 35^3 = 42875 GEL atoms (7.4 GB of lazily zeroed map), run on the spot.  Cell list,
    full-system pass (several thousand duplicate copies to resolve) and a 5000-iteration trace bit for bit on the cluster
    driver (AUTO at this size) and on the overlapped-sweep driver."""
    from oracle.make_golden import REFDIR as ref_dir
    if not (ref_dir / "ref_harness_gel").exists():
        pytest.skip("oracle/_ref/ref_harness_gel not built")
    n_it = 5000
    pre, post, fin = _live_case("gel35", harness="ref_harness_gel", coord="GEL35.gro", max_iter=n_it, n_gauss=80000, extra=[], gel_n=35)
    assert pre.natom == 42875
    want = ref_trace(fin)
    for mode in (0, 2):
        hv, r = _replica(capi, pre)
        r.set_loop_mode(mode)
        r.pbc_full(); r.cell_build()
        assert np.array_equal(r.celllist(), pre.extra["celllist"])
        r.predict_all(); r.cbt_build()
        r.download(hv)
        assert_fields_equal(hv, post, SLOT_FIELDS, what="after Predict(all): ")
        assert np.array_equal(r.cbt()[0][1:], post.extra["cbt_node"][1:])
        st, tr = r.run(max_iter=n_it, trace_cap=n_it)
        assert st["error"] == 0 and st["iterations"] == n_it
        assert_trace_equal(tr, want)
        r.download(hv)
        assert_fields_equal(hv, fin, STATE_FIELDS, what="final state: ")
        assert np.array_equal(r.celllist(), fin.extra["celllist"])
        assert r.potential_energy() == fin.extra["epot"][0]
        r.close()


def test_dense_gel_live_reference(capi):
    """Denser than SURVEY C5's 5 A variant: 14^3 GEL lattice at 4.5 A (about eight atoms per cell, > 200 candidates and ~33
    atoms inside the scan sphere per atom, so the filter's 32-entry rows overflow and the evaluation falls back to the
    cell scan), reference run on the spot; full-system pass and 5000 iterations bit for bit."""
    from oracle.make_golden import REFDIR as ref_dir
    if not (ref_dir / "ref_harness_gel").exists():
        pytest.skip("oracle/_ref/ref_harness_gel not built")
    n_it = 5000
    pre, post, fin = _live_case("gel14d", harness="ref_harness_gel", coord="GEL14.gro", max_iter=n_it, n_gauss=40000, extra=[], gel_n=14,
                                gel_a_nm=0.45)
    assert pre.natom == 2744
    for mode in (0, 2):
        hv, r = _replica(capi, pre)
        r.set_loop_mode(mode)
        r.pbc_full(); r.cell_build(); r.predict_all(); r.cbt_build()
        r.download(hv)
        assert_fields_equal(hv, post, SLOT_FIELDS, what="after Predict(all): ")
        st, tr = r.run(max_iter=n_it, trace_cap=n_it)
        assert st["error"] == 0
        assert_trace_equal(tr, ref_trace(fin))
        r.download(hv)
        assert_fields_equal(hv, fin, STATE_FIELDS, what="final state: ")
        r.close()


def test_warningsum_live_reference(capi):
    """`warningsum`: 8^3 GEL atoms 3.95 A apart (inside the 4.0 A hard core) make FindPair warn for every nearest-neighbour
    evaluation (ToolFunctions.c:119-139); `stats.warnings` of the CUDA path equals the reference's counter after Predict(all)
    plus 2000 loop iterations, on the one-block and the overlapped-sweep driver."""
    from oracle.make_golden import REFDIR as ref_dir
    if not (ref_dir / "ref_harness_gel").exists():
        pytest.skip("oracle/_ref/ref_harness_gel not built")
    n_it = 2000
    pre, post, fin = _live_case("gel8w", harness="ref_harness_gel", coord="GEL8.gro", max_iter=n_it, n_gauss=20000, extra=[], gel_n=8,
                                gel_a_nm=0.395)
    for mode in (0, 2):
        hv, r = _replica(capi, pre)
        r.set_loop_mode(mode)
        r.pbc_full(); r.cell_build(); r.predict_all(); r.cbt_build()
        r.download(hv)
        assert_fields_equal(hv, post, SLOT_FIELDS, what="after Predict(all): ")
        assert r.stats()["warnings"] == 6 * 512
        assert r.potential_energy() == post.extra["epot"][0]   # the harness calls CalPotenE here too: 3 * 512 more warnings
        st, tr = r.run(max_iter=n_it, trace_cap=n_it)
        assert st["error"] == 0
        assert_trace_equal(tr, ref_trace(fin))
        assert st["warnings"] == int(fin.extra["stats"][9]) > 9 * 512
        r.close()


def test_ala10_300k_iterations_live_reference(capi):
    """300 000 loop iterations of ala10 (HB formation and breaking included) against a trace the reference produces now."""
    from oracle.make_golden import REFDIR as ref_dir
    if not (ref_dir / "ref_harness").exists():
        pytest.skip("oracle/_ref/ref_harness not built")
    n_it = 300000
    pre, post, fin = _live_case("ala10_live", harness="ref_harness", coord="ala10.gro", max_iter=n_it, n_gauss=600000, extra=[], gel_n=0)
    hv, r = _replica(capi, pre)
    r.pbc_full(); r.cell_build(); r.predict_all(); r.cbt_build()
    want = ref_trace(fin)
    st, tr = r.run(max_iter=n_it, trace_cap=n_it)
    assert st["error"] == 0 and st["iterations"] == n_it and st["hb"] > 0
    assert_trace_equal(tr, want)
    r.download(hv)
    assert_fields_equal(hv, fin, STATE_FIELDS, what="final state: ")
    assert r.potential_energy() == fin.extra["epot"][0]


def test_error_paths_fail_loudly(capi):
    """Misuse comes back as an error code + message through the C-ABI, never as a silent no-op."""
    pre, _, _ = load_case("ala10")
    with pytest.raises(capi.SdmdError):
        capi.Replica(pre.copy(), device=4096)                 # no such device
    bad = pre.copy()
    bad.box = np.array([1e7, 150.0, 150.0])                   # > 65535 cells along x: rejected, and nothing leaks
    with pytest.raises(capi.SdmdError):
        capi.Replica(bad, device=0)
    hv = pre.copy()
    r = capi.Replica(hv, device=0)
    with pytest.raises(capi.SdmdError):
        r.predict_all()                                       # needs a fresh cell_build
    r.cell_build()
    with pytest.raises(capi.SdmdError):
        r.predict_all()                                       # thermostat on, no gaussians fed
    with pytest.raises(capi.SdmdError):
        r.rng_feed(np.zeros(1 << 22))                         # more than the ring holds
    with pytest.raises(capi.SdmdError):
        r.set_loop_mode(7)
    r.rng_feed(pre.extra["gauss"])
    r.predict_all(); r.cbt_build()
    st, _ = r.run(max_iter=10)
    assert st["iterations"] == 10 and st["error"] == 0        # and the context is still usable afterwards
