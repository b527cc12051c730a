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




def layers(src, dst, names=None):
    """by-name metric capture of one device batch (long CSV: one row per kernel and metric) -> per-layer table +
    DRAM bytes per chip (profiles/rNN_conv_traffic.json)."""
    import json
    lines = [ln for ln in open(src) if ln.startswith('"')]
    per = collections.OrderedDict()
    for row in csv.DictReader(lines):
        k = int(row['ID'])
        per.setdefault(k, {'kernel': row['Kernel Name'].split('(')[0].replace('void ', '').replace('scv::', ''), 'grid': row['Grid Size']})
        per[k][row['Metric Name']] = float(row['Metric Value'].replace(',', ''))
    short = {'gpu__time_duration.sum': 'us', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active': 'tensor%',
             'sm__pipe_tc_cycles_active.avg.pct_of_peak_sustained_elapsed': 'tc%', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed': 'dram%',
             'lts__throughput.avg.pct_of_peak_sustained_elapsed': 'lts%', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed': 'l1tex%',
             'sm__throughput.avg.pct_of_peak_sustained_elapsed': 'sm%', 'dram__bytes_read.sum': 'DRAM rd MB', 'dram__bytes_write.sum': 'DRAM wr MB',
             'launch__registers_per_thread': 'regs'}
    rows = [v for v in per.values()]
    conv = [r for r in rows if 'conv_' in r['kernel']]
    nchips = int(sys.argv[4]) if len(sys.argv) > 4 else 63
    with open(dst, 'w') as f:
        f.write(f'# ncu by-name metrics of one device batch ({nchips} chips) ({src}), --clock-control none\n\n')
        f.write('| # | kernel | grid | ' + ' | '.join(short.values()) + ' |\n|---|---|---|' + '---|' * len(short) + '\n')
        for i, r in enumerate(rows):
            vals = []
            for k, s in short.items():
                v = r.get(k, 0.0)
                vals.append(f'{v / 1e3:.1f}' if s == 'us' else (f'{v / 1e6:.1f}' if 'MB' in s else f'{v:.1f}'))
            f.write(f'| {i} | `{r["kernel"]}` | {r["grid"]} | ' + ' | '.join(vals) + ' |\n')
        tot_us = sum(r['gpu__time_duration.sum'] for r in conv) / 1e3
        tot_b = sum(r['dram__bytes_read.sum'] + r['dram__bytes_write.sum'] for r in conv)
        f.write(f'\n{len(conv)} conv launches: {tot_us:.0f} us, {tot_b / 1e9:.2f} GB of DRAM traffic for {nchips} chips = '
                f'{tot_b / nchips / 1e6:.1f} MB per chip (algorithmic FLOPs 67.41 GFLOP per chip -> '
                f'{67.41e9 * nchips / (tot_us * 1e-6) / 1e12:.0f} TFLOP/s under ncu clocks).\n')
    with open(dst.replace('_all_layers_metrics.md', '_conv_traffic.json'), 'w') as f:
        json.dump({'dram_bytes_per_chip': tot_b / nchips, 'conv_launches': len(conv), 'chips': nchips, 'source': src}, f)


if __name__ == '__main__':
    {'launches': launches, 'full': full, 'layers': layers}[sys.argv[1]](sys.argv[2], sys.argv[3])
