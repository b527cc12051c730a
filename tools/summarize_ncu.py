"""Turn ncu outputs brought back in gpurun_out/ into small tracked summaries under profiles/.

  python tools/summarize_ncu.py launches gpurun_out/x_launches.csv profiles/rNN_launches.md
  python tools/summarize_ncu.py full     gpurun_out/x.ncu-rep     profiles/rNN_full.md
"""
import collections
import csv
import subprocess
import sys

KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread',
        'launch__occupancy_limit_shared_mem', 'smsp__inst_executed.sum']


def launches(src, dst):
    lines = [ln for ln in open(src) if not ln.startswith('==')]
    agg = collections.OrderedDict()
    for row in csv.DictReader(lines):
        if row.get('Metric Name') != 'gpu__time_duration.sum':
            continue
        k = (row['Kernel Name'].split('(')[0], row['Grid Size'], row['Block Size'])
        agg.setdefault(k, []).append(float(row['Metric Value'].replace(',', '')))
    tot = sum(sum(v) for v in agg.values())
    with open(dst, 'w') as f:
        f.write(f'# ncu launch list ({src}): gpu__time_duration.sum, --clock-control none\n\n')
        f.write('Per-launch times are cold-cache and serialised: compare SHARES, not absolutes.\n\n')
        f.write('| share | launches | avg us | kernel | grid | block |\n|---|---|---|---|---|---|\n')
        for (name, grid, block), v in agg.items():
            f.write(f'| {sum(v) / tot * 100:.1f}% | {len(v)} | {sum(v) / len(v) / 1e3:.1f} | `{name}` | {grid} | {block} |\n')
        f.write(f'\ntotal {tot / 1e6:.3f} ms over {sum(len(v) for v in agg.values())} launches\n')


def full(src, dst):
    raw = subprocess.run(['ncu', '-i', src, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {k: hdr.index(k) for k in KEYS if k in hdr}
    ki, gi = hdr.index('Kernel Name'), hdr.index('Grid Size')
    with open(dst, 'w') as f:
        f.write(f'# ncu --set full ({src}), --clock-control none\n\n')
        f.write('| kernel | grid | ' + ' | '.join(k.split('.')[0] for k in idx) + ' |\n')
        f.write('|---|---|' + '---|' * len(idx) + '\n')
        f.write('| (unit) | | ' + ' | '.join(units[i] for i in idx.values()) + ' |\n')
        for r in rows[2:]:
            name = r[ki].split('(')[0].replace('void ', '').replace('scv::', '')
            f.write(f'| `{name}` | {r[gi]} | ' + ' | '.join(r[i] for i in idx.values()) + ' |\n')


if __name__ == '__main__':
    {'launches': launches, 'full': full}[sys.argv[1]](sys.argv[2], sys.argv[3])
