"""GPU: the reference's own driver with the hot path on the GPU (sdmd_b200/host/_build/sDMD_cuda) must write the
same trajectory and connection map, byte for byte, as the unmodified reference binary on the same input."""
import shutil
import subprocess
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]
REF = ROOT / "oracle" / "_ref"
DROPIN = ROOT / "sdmd_b200" / "host" / "_build" / "sDMD_cuda"

pytestmark = pytest.mark.gpu


def _run(binary, work: Path, time_units):
    work.mkdir(parents=True)
    shutil.copytree(REF / "input", work / "input")
    p = work / "input" / "parameter.txt"
    lines = p.read_text().splitlines()
    lines[1] = "Time               %g" % time_units
    p.write_text("\n".join(lines) + "\n")
    (work / "out").mkdir()
    r = subprocess.run([str(binary), "-f", "ala10.gro", "-i", str(work / "input"), "-o", str(work / "out")],
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
    for key in ("ReCalculate times", "collision times", "bond times", "HB times", "HB Neighbor times", "thermostat times",
                "PBC&CC times"):
        assert _counter(ref_out, key) == _counter(gpu_out, key), key
    assert _counter(gpu_out, "bond times") > 100000
