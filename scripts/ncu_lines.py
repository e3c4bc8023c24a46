"""Aggregate `ncu --page source --print-source cuda,sass --csv` output by CUDA source line.
usage: ncu -i X.ncu-rep --page source --print-source cuda,sass --csv | python scripts/ncu_lines.py [topN]"""
import csv, sys
top = int(sys.argv[1]) if len(sys.argv) > 1 else 50
rows = list(csv.reader(sys.stdin))
cur_file = ""
hdr = None
agg = {}
tot = 0
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur_file = r[1].split("/")[-1]; continue
    if r[0] == "Function Name": continue
    if r[0] == "Line No": hdr = r; si = r.index("# Samples"); ii = r.index("Instructions Executed"); bi = r.index("stall_barrier"); continue
    if hdr is None or r[0] == "": continue
    try: s = int(r[si]) - int(r[bi] or 0); ins = int(r[ii])
    except Exception: continue
    key = (cur_file, int(r[0]), r[1].strip()[:100])
    a = agg.setdefault(key, [0, 0]); a[0] += s; a[1] += ins
    tot += s
print("total non-barrier samples", tot)
for (f, ln, src), (s, ins) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{s:7d} {100*s/max(tot,1):5.1f}%  inst {ins:9d}  {f}:{ln}  {src}")
