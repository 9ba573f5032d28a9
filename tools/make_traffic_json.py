"""ncu launch list of tools/fault_hunt.py (--passes 1 --max-batches 1) -> profiles/r02_traffic.json, the file bench.py's
roofline.traffic / roofline.hbm_classes read: per-launch DRAM bytes of the kernel classes next to their algorithmic bytes.

    python tools/make_traffic_json.py launches.csv rows_in_batch out.json [committed csv name]
"""
import collections
import csv
import json
import sys

path, M, out = sys.argv[1], int(sys.argv[2]), sys.argv[3]
name = sys.argv[4] if len(sys.argv) > 4 else path
rows = [r for r in csv.reader(open(path, errors='replace')) if len(r) > 5]
hdr = next(i for i, r in enumerate(rows) if 'Kernel Name' in r)
h = rows[hdr]
ki, mi, vi, ui, ii = h.index('Kernel Name'), h.index('Metric Name'), h.index('Metric Value'), h.index('Metric Unit'), h.index('ID')
per = collections.OrderedDict()
for r in rows[hdr + 1:]:
    try:
        v = float(r[vi].replace(',', ''))
    except ValueError:
        continue
    unit = r[ui].strip()
    if r[mi] == 'gpu__time_duration.sum':
        v *= {'ns': 1e-3, 'us': 1.0, 'ms': 1e3, 's': 1e6}.get(unit, 1e-3)
    else:
        v *= {'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}.get(unit, 1.0)
    per.setdefault(r[ii], {'name': r[ki]})[r[mi]] = v


def cls(pred):
    sel = [d for d in per.values() if pred(d['name'])]
    n = len(sel)
    if not n:
        return None
    rd = sum(d.get('dram__bytes_read.sum', 0.0) for d in sel); wr = sum(d.get('dram__bytes_write.sum', 0.0) for d in sel)
    us = sum(d.get('gpu__time_duration.sum', 0.0) for d in sel)
    return {'kernel': sel[0]['name'].split('(')[0], 'launches': n, 'dram_bytes_per_launch': (rd + wr) / n,
            'dram_gbps_under_ncu': (rd + wr) / (us * 1e-6) / 1e9, 'us_per_launch': us / n}


gemm = cls(lambda k: 'gemm_tc_kernel<256' in k and ', 100,' not in k)
# 8 GEMMs per layer: A [M,K] + out [M,N] in bf16 + W [N,K]: (1024+4096) + (4096+1024) + (1024+3072) + (1024+1024) + (1024+2048->1024 GLU)
# + (1024+1024) + (1024+4096) + (4096+1024) elements per row
per_row = 2 * ((1024 + 4096) * 2 * 2 + (1024 + 3072) + (1024 + 1024) * 2 + (1024 + 1024))
w_bytes = 2 * (4 * 1024 * 4096 + 3072 * 1024 + 1024 * 1024 * 2 + 2048 * 1024)
alg_gemm = (M * per_row + w_bytes) / 8.0
res = {'note': f'ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none over one batch of the '
               f'bench shard (M = {M} token rows, 19 layers): profiles/{name}; per-launch DRAM bytes (read + write); times under ncu are '
               'cold-cache and serialised',
       'rows_per_batch': M,
       'gemm_tc_kernel': {**{k: v for k, v in gemm.items() if k != 'kernel'}, 'algorithmic_bytes_per_launch': alg_gemm,
                          'ratio': gemm['dram_bytes_per_launch'] / alg_gemm,
                          'algorithmic': 'per launch, averaged over the 8 GEMMs of a layer: bf16 A [M,K] + bf16 out [M,N] + bf16 W [N,K]'},
       'hbm_classes': {}}
for key, pred, alg, what in (
        ('add_layernorm', lambda k: 'add_layernorm_kernel<__nv_bfloat16, 0>' in k or 'add_layernorm_kernel<__nv_bfloat16, false>' in k, M * 12288,
         'fp32 x read + written (8 KB/row), bf16 d read (2 KB), bf16 LN output written (2 KB)'),
        ('dwconv_ring', lambda k: 'dwconv_ring' in k or 'dwconv_mma' in k, M * 4096, 'bf16 GLU output read (2 KB/row) + bf16 output written (2 KB/row)'),
        ('attention', lambda k: 'attention_tc' in k, M * 8192,
         'bf16 qkv read once (6 KB/row) + bf16 output written (2 KB/row); K/V re-reads by the query tiles of a clip hit L2')):
    c = cls(pred)
    if c:
        c['algorithmic_bytes_per_launch'] = alg
        c['ratio'] = c['dram_bytes_per_launch'] / alg
        c['algorithmic'] = what
        res['hbm_classes'][key] = c
json.dump(res, open(out, 'w'), indent=1)
print(json.dumps({k: (v if not isinstance(v, dict) else {kk: vv for kk, vv in v.items() if kk in ('launches', 'ratio', 'dram_gbps_under_ncu', 'us_per_launch')})
                  for k, v in res.items() if k != 'note'}, indent=1)[:1500])
