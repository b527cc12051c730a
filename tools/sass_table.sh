#!/bin/bash
# SASS evidence per kernel family: counts of the tcgen05 / TMA mnemonics in libscv.so (cuobjdump -sass).
LIB=satellite_computervision_b200/libscv.so
cuobjdump -sass $LIB > /tmp/scv.sass
python - <<'P'
import re, collections
fam = collections.OrderedDict()
cur = None
keys = ['UTCHMMA', 'UTCBAR', 'LDTM', 'STTM', 'UTMALDG', 'UTMASTG', 'UBLKCP', 'SYNCS', 'UTCATOMSWS']
for line in open('/tmp/scv.sass'):
    m = re.search(r'Function : (\S+)', line)
    if m:
        name = m.group(1)
        f = re.search(r'(conv_fused2_kernel|conv_rows_kernel|conv_slabw_kernel|conv_slab2_kernel|conv_slab_kernel|conv_ptile_kernel|conv_umma_kernel|extract_u16x6_kernel|extract_kernel|stitch_kernel_vec|stitch_kernel_scalar|tile_stats_kernel|head_tiles_kernel)', name)
        cur = f.group(1) if f else 'other'
        fam.setdefault(cur, collections.Counter())['__instances'] += 1
        continue
    if cur is None:
        continue
    for k in keys:
        if re.search(r'\b' + k, line):
            fam[cur][k] += 1
print('| kernel family | instantiations | ' + ' | '.join(keys) + ' |')
print('|---|---|' + '---|' * len(keys))
for f, c in fam.items():
    print(f'| `{f}` | {c["__instances"]} | ' + ' | '.join(str(c[k]) for k in keys) + ' |')
P
