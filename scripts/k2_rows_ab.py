"""A/B of the two round-0 paths on the 1M-atom box: SDMD_K2_ROWS=0 (cell-by-cell filter + ordered evaluation, sdmd_cuda.cu)
against the default (row-batched filter + cp.async evaluation, sdmd_k2_rows.cu).  Each leg in a child process."""
import os, subprocess, sys
for rows in ("0", "1"):
    env = dict(os.environ, SDMD_K2_ROWS=rows)
    print(f"=== SDMD_K2_ROWS={rows}", flush=True)
    r = subprocess.run([sys.executable, "scripts/k2_breakdown.py", *sys.argv[1:]], env=env, capture_output=True, text=True, timeout=600)
    print(r.stdout + r.stderr[-2000:], flush=True)
