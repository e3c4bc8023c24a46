"""GPU: the reference's own driver with the hot path on the GPU (sdmd_b200/host/_build/sDMD_cuda) must write the
same trajectory and connection map, byte for byte, as the unmodified reference binary on the same input."""
import os
import shutil
import subprocess
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]
REF = ROOT / "oracle" / "_ref"
# the fixed-seed build (-DDEBUG_RANDOM, like oracle/_ref/sDMD_ref): byte-for-byte comparisons need the same seed on both sides;
# the production binary sDMD_cuda seeds from time(NULL) like upstream
DROPIN = ROOT / "sdmd_b200" / "host" / "_build" / "sDMD_cuda_fixedseed"

pytestmark = pytest.mark.gpu

# every drop-in run of this module also cross-checks the glue's sparse connection rows against the dense connectionMap scan
os.environ["SDMD_GLUE_CHECK_CM"] = "1"


def _run(binary, work: Path, time_units):
    work.mkdir(parents=True)
    shutil.copytree(REF / "input", work / "input")
    p = work / "input" / "parameter.txt"
    lines = p.read_text().splitlines()
    lines[1] = "Time               %g" % time_units
    p.write_text("\n".join(lines) + "\n")
    (work / "out").mkdir()
    # every optional per-dump output the reference offers: potential / kinetic energy, temperature, hydrogen-bond tallies
    r = subprocess.run([str(binary), "-f", "ala10.gro", "-i", str(work / "input"), "-o", str(work / "out"), "-pot", "-kin", "-tem", "-HBn"],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    return r.stdout.replace("\r", "\n")


def _counter(out, key):
    for line in out.splitlines():
        if line.startswith(key):
            return int(line.split("=")[1])
    raise AssertionError(key)


def test_dropin_driver_writes_identical_outputs(tmp_path):
    if not (REF / "sDMD_ref").exists() or not DROPIN.exists():
        pytest.skip("needs oracle/_ref/sDMD_ref and the drop-in driver (built where /root/reference exists)")
    ref_out = _run(REF / "sDMD_ref", tmp_path / "ref", 300)
    gpu_out = _run(DROPIN, tmp_path / "gpu", 300)
    for name in ("out_trj.gro", "out_cnt.txt"):
        a = (tmp_path / "ref" / "out" / name).read_bytes()
        b = (tmp_path / "gpu" / "out" / name).read_bytes()
        assert len(a) > 1000 and a == b, name
    # the per-dump outputs (-pot -kin -tem -HBn): same set of files, same bytes; the HB file carries HBnumformed and alphaHBformed
    def _outputs(d):      # (savedData.dat holds a time stamp-dependent seed, timeEvaluation.dat wall-clock times, the log the command line)
        return sorted(p.name for p in d.iterdir() if p.is_file() and not p.name.startswith(("savedData", "timeEvaluation")) and "log" not in p.name.lower())
    ref_files, gpu_files = _outputs(tmp_path / "ref" / "out"), _outputs(tmp_path / "gpu" / "out")
    assert ref_files == gpu_files and len(ref_files) >= 6, (ref_files, gpu_files)
    for name in ref_files:
        if name == "SysInfo.dat":      # raw structs with heap pointers in them (PropertyStr, ConstraintStr: DataSave.c:225-256)
            continue
        assert (tmp_path / "ref" / "out" / name).read_bytes() == (tmp_path / "gpu" / "out" / name).read_bytes(), name
    hb = [ln.split() for ln in (tmp_path / "gpu" / "out" / next(n for n in gpu_files if "HB" in n)).read_text().splitlines() if ln and ln[0].isdigit()]
    assert max(int(r[1]) for r in hb) > 0, "no hydrogen bond was ever formed in the run: the HB tallies are not exercised"
    for key in ("ReCalculate times", "collision times", "bond times", "HB times", "HB Neighbor times", "thermostat times",
                "PBC&CC times"):
        assert _counter(ref_out, key) == _counter(gpu_out, key), key
    assert _counter(gpu_out, "bond times") > 100000


def _set_param(work, line_no, text):
    p = work / "input" / "parameter.txt"
    lines = p.read_text().splitlines()
    lines[line_no] = text
    p.write_text("\n".join(lines) + "\n")


@pytest.mark.parametrize("first_leg,second_leg", [("ref", "gpu"), ("gpu", "gpu"), ("gpu", "ref")],
                         ids=["ref_then_dropin", "dropin_then_dropin", "dropin_then_ref"])
def test_dropin_continue_run_matches_reference(tmp_path, first_leg, second_leg):
    """`continue` runs (parameter.txt line 1, the annealing schedule of demo/Folding/fold.bash): 100 time units new,
    then continued to 200 from savedData.dat.  Every combination of who writes the checkpoint and who resumes from it —
    stock -> drop-in, drop-in -> drop-in (a savedData.dat the GPU driver wrote, read back by the GPU driver) and
    drop-in -> stock — must end with the files the all-stock pair of runs writes."""
    if not (REF / "sDMD_ref").exists() or not DROPIN.exists():
        pytest.skip("needs oracle/_ref/sDMD_ref and the drop-in driver")
    binaries = {"ref": REF / "sDMD_ref", "gpu": DROPIN}
    outs = {}
    for tag, legs in (("want", ("ref", "ref")), ("got", (first_leg, second_leg))):
        work = tmp_path / tag
        _run(binaries[legs[0]], work, 100)
        _set_param(work, 0, "Continue?          continue")
        _set_param(work, 1, "Time               200")
        r = subprocess.run([str(binaries[legs[1]]), "-f", "ala10.gro", "-i", str(work / "input"), "-o", str(work / "out")],
                           capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
        outs[tag] = r.stdout.replace("\r", "\n")
    # (savedData.dat itself holds raw heap pointers of the writing process — two stock runs already differ in it — so what is
    # compared is everything the resumed run derives from it)
    for name in ("out_trj.gro", "out_cnt.txt"):
        a = (tmp_path / "want" / "out" / name).read_bytes()
        b = (tmp_path / "got" / "out" / name).read_bytes()
        assert len(a) > 1000 and a == b, name
    for key in ("collision times", "bond times", "thermostat times", "PBC&CC times"):
        assert _counter(outs["want"], key) == _counter(outs["got"], key), key


@pytest.mark.parametrize("wall_type", ["parallel", "cylinder", "obstacles", "tunnel"])
def test_dropin_walled_run_matches_reference(tmp_path, wall_type):
    """BASELINE config 4 (confinement): WallExist smooth with WallType parallel / cylinder, or flat obstruction walls plus
    hard spherical obstacles, or a tunnel between two holed obstruction walls, chains of ala10 inside — the drop-in driver must write the same trajectory and counters
    (wall events included) as the stock binary."""
    if not (REF / "sDMD_ref").exists() or not DROPIN.exists():
        pytest.skip("needs oracle/_ref/sDMD_ref and the drop-in driver")
    from oracle.make_golden import WALLS, prepare_input, write_multi_ala10
    spec = WALLS[{"parallel": "ala10x8_wallp", "cylinder": "ala10x4_wallc", "obstacles": "ala10x8_obst", "tunnel": "ala10x4_tunnel"}[wall_type]]
    _, copies, spacing, box = spec[:4]
    outs = {}
    for tag, binary in (("ref", REF / "sDMD_ref"), ("gpu", DROPIN)):
        work = tmp_path / tag
        work.mkdir()
        inp = prepare_input(work, time_units=60, wall_type=spec[0], obst=spec[4] if len(spec) > 4 else None,
                            sphobst=spec[5] if len(spec) > 5 else None, tunnel=spec[6] if len(spec) > 6 else None)
        write_multi_ala10(inp / "walled.gro", inp / "ala10.gro", copies_per_axis=copies, spacing_nm=spacing, box_nm=box)
        r = subprocess.run([str(binary), "-f", "walled.gro", "-i", str(inp), "-o", str(work / "out")],
                           capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
        outs[tag] = r.stdout.replace("\r", "\n")
    for name in ("out_trj.gro", "out_cnt.txt"):
        a = (tmp_path / "ref" / "out" / name).read_bytes()
        b = (tmp_path / "gpu" / "out" / name).read_bytes()
        assert len(a) > 1000 and a == b, name
    # the same SET of output files (WallInfo.dat / ObstInfo.dat of Prepare(), SGThread.c:37-38, are read by the analysis tools:
    # analysis/SystemInformation.c:373-383) with the same bytes; the log carries wall-clock stamps and is left out
    names = lambda tag: sorted(p.name for p in (tmp_path / tag / "out").iterdir())
    assert names("ref") == names("gpu")
    geometry = [n for n in names("ref") if n in ("WallInfo.dat", "ObstInfo.dat")]   # (SysInfo.dat holds raw heap pointers)
    assert geometry
    for name in geometry:
        assert (tmp_path / "ref" / "out" / name).read_bytes() == (tmp_path / "gpu" / "out" / name).read_bytes(), name
    for key in ("collision times", "bond times", "thermostat times", "PBC&CC times", "Wall times"):
        assert _counter(outs["ref"], key) == _counter(outs["gpu"], key), key
    assert _counter(outs["gpu"], "Wall times") > 0


@pytest.mark.parametrize("mark", ["compress", "pulse", "expand"])
def test_dropin_moving_wall_matches_reference(tmp_path, mark):
    """"WallDyn compress | pulse | expand" (moving smooth wall: DoWallDyn Events.c:502-540 once per time unit, WallTime's retry
    TimePrediction.c:675-697 when an atom is found outside): the drop-in driver writes the same trajectory — box edges in every
    frame included — and the same counters as the stock binary."""
    if not (REF / "sDMD_ref").exists() or not DROPIN.exists():
        pytest.skip("needs oracle/_ref/sDMD_ref and the drop-in driver")
    from oracle.make_golden import WALLS, prepare_input, write_multi_ala10
    spec = WALLS["ala10x8_wallp"]
    outs = {}
    for tag, binary in (("ref", REF / "sDMD_ref"), ("gpu", DROPIN)):
        work = tmp_path / tag
        work.mkdir()
        inp = prepare_input(work, time_units=80, wall_type=spec[0])
        p = inp / "parameter.txt"
        lines = p.read_text().splitlines()
        lines[11] = "WallDyn            " + mark
        p.write_text("\n".join(lines) + "\n")
        write_multi_ala10(inp / "walled.gro", inp / "ala10.gro", copies_per_axis=spec[1], spacing_nm=spec[2], box_nm=spec[3])
        r = subprocess.run([str(binary), "-f", "walled.gro", "-i", str(inp), "-o", str(work / "out"), "-Wsz", "5", "-Wrt", "20"],
                           capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
        outs[tag] = r.stdout.replace("\r", "\n")
    for name in ("out_trj.gro", "out_cnt.txt"):
        a = (tmp_path / "ref" / "out" / name).read_bytes()
        b = (tmp_path / "gpu" / "out" / name).read_bytes()
        assert len(a) > 1000 and a == b, name
    boxes = {ln for ln in (tmp_path / "gpu" / "out" / "out_trj.gro").read_text().splitlines() if len(ln.split()) == 3 and ln.startswith("   ")}
    assert len(boxes) > (3 if mark != "expand" else 0), boxes          # the box really moved ("expand" from the original size ends after one step)
    for key in ("collision times", "bond times", "thermostat times", "PBC&CC times", "Wall times"):
        assert _counter(outs["ref"], key) == _counter(outs["gpu"], key), key


@pytest.mark.parametrize("flow", ["velocity 0.02 0.0 0.01", "force 5.0 0.0 0.0", "charge 1 : 100.0 75.0 75.0 0.05 0.2"],
                         ids=["velocity", "force", "charge"])
def test_dropin_flow_run_matches_reference(tmp_path, flow):
    """"Flow velocity ..." / "Flow force ..." / "Flow charge ..." (drift terms in ThermostatEvent and in the temperature CalKinetE
    reports; charge sources with their own predictor and event): the
    drop-in driver writes the same trajectory, energies and counters as the stock binary."""
    if not (REF / "sDMD_ref").exists() or not DROPIN.exists():
        pytest.skip("needs oracle/_ref/sDMD_ref and the drop-in driver")
    from oracle.make_golden import prepare_input
    outs = {}
    for tag, binary in (("ref", REF / "sDMD_ref"), ("gpu", DROPIN)):
        work = tmp_path / tag
        work.mkdir()
        inp = prepare_input(work, time_units=150, flow=flow)
        extra = ["-COM"] if flow.startswith("charge") else []      # charge sources: also the centre-of-mass thermostat term
        r = subprocess.run([str(binary), "-f", "ala10.gro", "-i", str(inp), "-o", str(work / "out"), *extra],
                           capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
        outs[tag] = r.stdout.replace("\r", "\n")
    for name in ("out_trj.gro", "out_cnt.txt"):
        a = (tmp_path / "ref" / "out" / name).read_bytes()
        b = (tmp_path / "gpu" / "out" / name).read_bytes()
        assert len(a) > 1000 and a == b, name
    for key in ("collision times", "bond times", "thermostat times", "PBC&CC times"):
        assert _counter(outs["ref"], key) == _counter(outs["gpu"], key), key


def test_dropin_remd_with_reference_server(tmp_path):
    """demo/REMD-style run with the reference's own sServer: three replica processes of the drop-in driver (sharing
    the GPU) must log the same temperature walk (out_REMD<id>.txt) and trajectories as three stock sDMD processes."""
    if not (REF / "sServer").exists() or not DROPIN.exists():
        pytest.skip("needs oracle/_ref and the drop-in driver")
    from tests.remd_ref import run_reference_remd
    ref_logs, _ = run_reference_remd(tmp_path / "r", 3, 5.0, 40.0, 99)
    gpu_logs, _ = run_reference_remd(tmp_path / "g", 3, 5.0, 40.0, 99, binary=DROPIN)
    assert ref_logs == gpu_logs
    assert len(ref_logs[0].splitlines()) >= 7
    for k in range(3):
        a = (tmp_path / "r" / f"ref{k}" / "out").glob("out_trj*")
        b = (tmp_path / "g" / f"ref{k}" / "out").glob("out_trj*")
        assert next(a).read_bytes() == next(b).read_bytes()


def _device_count():
    from sdmd_b200 import capi
    return int(capi.lib().sdmd_cuda_device_count())


@pytest.mark.parametrize("nrep", [2, 4, 8])
def test_remd_launcher_nccl_matches_reference_server(tmp_path, nrep):
    """The C launcher (sREMD_cuda, role of REMDExec.c) with its default exchange — one NCCL communicator over the replicas, one
    GPU each, CalPotenE + all-gather + replicated Server.c sweep inside sdmd_remd_exchange_energy — against the stock binaries
    coupled through the reference's sServer with the same server seed: every replica's temperature walk (out_REMD<id>.txt) and
    trajectory must be identical.  Needs one GPU per replica (run with `gpurun --gpus N`); skipped on smaller boxes."""
    launcher = DROPIN.parent / "sREMD_cuda"
    if not (REF / "sServer").exists() or not DROPIN.exists() or not launcher.exists():
        pytest.skip("needs oracle/_ref and the drop-in driver + launcher")
    if _device_count() < nrep:
        pytest.skip(f"needs {nrep} GPUs")
    from tests.remd_ref import TEMPS_K, run_reference_remd
    rate, tu, seed = 5.0, 40.0, 99
    ref_logs, _ = run_reference_remd(tmp_path / "r", nrep, rate, tu, seed)
    work = tmp_path / "g"
    work.mkdir()
    shutil.copytree(REF / "input", work / "input")
    lines = (REF / "demo_REMD" / "parameter.txt").read_text().splitlines()
    lines[1] = "Time               %g" % tu
    (work / "input" / "parameter.txt").write_text("\n".join(lines) + "\n")
    (work / "out").mkdir()
    (work / "REMDConfig.txt").write_text("#REPLICA EXCHANGE MD CONFIGURATION FILE\n"
                                         f"ReplicaNum              = {nrep}\n"
                                         "Temperatures            = " + " ".join("%g" % t for t in TEMPS_K[:nrep]) + "\n"
                                         "SocketPort              = 20202\n"
                                         f"ExchangeRate            = {int(rate)}\n")
    r = subprocess.run([str(launcher), "-f", "REMDConfig.txt", "-args", "-i", str(work / "input"), "-f", "ala10.gro", "-o", str(work / "out"), "-end",
                        "-seed", str(seed), "-exe", str(DROPIN)], cwd=work, capture_output=True, text=True, timeout=240)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:] + "".join((work / f"outputData_{k}.txt").read_text()[-800:] for k in range(nrep))
    for k in range(nrep):
        got = list((work / "out").glob(f"out_REMD*{k}.txt"))
        assert len(got) == 1, sorted(p.name for p in (work / "out").iterdir())
        assert got[0].read_text() == ref_logs[k], k
        a = next((tmp_path / "r" / f"ref{k}" / "out").glob("out_trj*"))
        b = next((work / "out").glob(f"out_trj*{k}.gro"))
        assert a.read_bytes() == b.read_bytes(), k
    assert any(len(set(ln.split()[2] for ln in log.splitlines())) > 1 for log in ref_logs)      # somebody actually swapped
