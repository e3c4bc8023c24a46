"""Summarise an .ncu-rep here (no GPU needed): headline metrics per kernel + the hottest SASS regions.
usage: python scripts/ncu_summary.py gpurun_out/x.ncu-rep [--sass N]"""
import csv, subprocess, sys, io
rep = sys.argv[1]
WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__thread_inst_executed_per_inst_executed.ratio', 'smsp__inst_executed.sum', 'launch__registers_per_thread',
        'launch__grid_size', 'launch__block_size', 'launch__shared_mem_per_block_dynamic', 'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_registers',
        'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'sm__cycles_active.avg', 'sm__cycles_elapsed.max', 'launch__waves_per_multiprocessor']
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[0]
for r in rows[2:]:
    print('###', r[hdr.index('Kernel Name')][:120])
    for i, h in enumerate(hdr):
        if h in WANT or ('issue_stalled' in h and h.endswith('per_issue_active.ratio') and float(r[i] or 0) > 0.3):
            print(f'- {h} = {r[i]} {rows[1][i]}')
if '--sass' in sys.argv:
    n = int(sys.argv[sys.argv.index('--sass') + 1])
    src = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(src)))
    hdr = rows[1]; data = rows[2:]
    ia = hdr.index('Instructions Executed'); isamp = hdr.index('# Samples'); isrc = hdr.index('Source'); ith = hdr.index('Avg. Threads Executed')
    tot = sum(int(r[ia]) for r in data); ts = sum(int(r[isamp]) for r in data)
    print(f'total warp instructions {tot}, samples {ts}, SASS lines {len(data)}')
    for s in range(0, len(data), n):
        seg = data[s:s + n]
        ni = sum(int(r[ia]) for r in seg); sm = sum(int(r[isamp]) for r in seg)
        if ni > tot * 0.01 or sm > ts * 0.01:
            print(f'{s:6d}-{s+n:6d} inst {ni/1e6:8.2f}M ({100*ni/tot:5.1f}%) samples {sm:6d} ({100*sm/ts:5.1f}%)  {seg[0][isrc].strip()[:50]}')
