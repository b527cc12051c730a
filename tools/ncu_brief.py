"""Brief per-kernel table from an .ncu-rep (raw page): python tools/ncu_brief.py file.ncu-rep [extra metric substrings...]"""
import csv
import subprocess
import sys

BASE = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'sm__pipe_tc_cycles_active.avg.pct_of_peak_sustained_elapsed',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__t_sector_hit_rate.pct', 'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread',
        'smsp__cycles_active.avg', 'lts__t_bytes.sum']
raw = subprocess.run(['ncu', '-i', sys.argv[1], '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
extra = sys.argv[2:]
for r in rows[2:]:
    print('==', r[hdr.index('Kernel Name')][:110], 'grid', r[hdr.index('Grid Size')])
    for k in BASE:
        if k in hdr:
            print(f'   {k:80s} {r[hdr.index(k)]:>16s} {units[hdr.index(k)]}')
    for i, h in enumerate(hdr):
        if any(e in h for e in extra):
            try:
                v = float(r[i].replace(',', ''))
            except ValueError:
                continue
            if 'issue_stalled' in h and v < 0.2:
                continue
            print(f'   {h[:80]:80s} {r[i]:>16s} {units[i]}')
