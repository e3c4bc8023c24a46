"""profiles/ncu_traffic.json from an `ncu --set full` capture (no GPU needed): dram__bytes_read.sum + dram__bytes_write.sum per launch
of every kernel in the report, averaged over its launches.  bench.py's `roofline.traffic` reads the entries of the dominant
kernels from that file, so the figure always comes from a capture of the kernels as they are now:

    gpurun -- 'ncu --set full --clock-control none -k regex:k_predict_all -s 8 -c 4 -o gpurun_out/r2_k2 python scripts/k2_breakdown.py'
    python scripts/ncu_traffic.py gpurun_out/r2_k2.ncu-rep 1000000 "python scripts/k2_breakdown.py (1M-atom GEL box)"
"""
import csv, io, json, subprocess, sys
from pathlib import Path

rep, natom, what = sys.argv[1], int(sys.argv[2]), sys.argv[3]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
col = {h: i for i, h in enumerate(hdr)}
SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
TSCALE = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6, "nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3, "second": 1e6}


def val(r, name, scale):
    i = col[name]
    return float(r[i].replace(",", "")) * scale[units[i]]


acc = {}
for r in rows[2:]:
    name = r[col["Kernel Name"]].split("(")[0].replace("void ", "").strip()
    b = val(r, "dram__bytes_read.sum", SCALE) + val(r, "dram__bytes_write.sum", SCALE)
    t = val(r, "gpu__time_duration.sum", TSCALE)
    a = acc.setdefault(name, {"dram_bytes": 0.0, "time_us": 0.0, "launches": 0})
    a["dram_bytes"] += b; a["time_us"] += t; a["launches"] += 1
for a in acc.values():
    a["dram_bytes"] /= a["launches"]; a["time_us"] /= a["launches"]
out = {"natom": natom, "source": f"{Path(rep).name}: ncu --set full --clock-control none, {what}; dram__bytes_read.sum + dram__bytes_write.sum per launch "
                                 "(scripts/ncu_traffic.py)", "kernels": acc}
dst = Path(__file__).resolve().parents[1] / "profiles" / "ncu_traffic.json"
dst.write_text(json.dumps(out, indent=1) + "\n")
print(json.dumps(out, indent=1))
