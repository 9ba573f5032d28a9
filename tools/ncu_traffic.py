"""ncu launch list (gpu__time_duration.sum, dram__bytes_read.sum, dram__bytes_write.sum per launch, --csv) ->
per-kernel summary: launches, total / mean time, mean DRAM bytes per launch, achieved DRAM GB/s; written as JSON
(profiles/r02_traffic.json is what bench.py's roofline.traffic reads) and printed as a table.

    python tools/ncu_traffic.py launches.csv [out.json] [rows=65496]
"""
import collections
import csv
import json
import sys

rows = [r for r in csv.reader(open(sys.argv[1], errors='replace')) if len(r) > 5]
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
        v *= {'ns': 1e-3, 'us': 1.0, 'ms': 1e3, 's': 1e6}.get(unit, 1e-3)              # -> us
    else:
        v *= {'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}.get(unit, 1.0)     # -> bytes
    name = r[ki].split('(')[0].replace('void ', '').replace('(anonymous namespace)::', '')
    per.setdefault(r[ii], {'name': name})[r[mi]] = v
agg = collections.defaultdict(lambda: {'launches': 0, 'us': 0.0, 'rd': 0.0, 'wr': 0.0})
for d in per.values():
    a = agg[d['name']]
    a['launches'] += 1
    a['us'] += d.get('gpu__time_duration.sum', 0.0)
    a['rd'] += d.get('dram__bytes_read.sum', 0.0)
    a['wr'] += d.get('dram__bytes_write.sum', 0.0)
tot = sum(a['us'] for a in agg.values())
out = {}
print(f'total {tot / 1e3:.2f} ms over {sum(a["launches"] for a in agg.values())} launches')
for k, a in sorted(agg.items(), key=lambda kv: -kv[1]['us']):
    n = a['launches']
    byts = (a['rd'] + a['wr']) / n
    gbs = (a['rd'] + a['wr']) / (a['us'] * 1e-6) / 1e9 if a['us'] > 0 else 0.0
    out[k[:80]] = {'launches': n, 'ms_total': a['us'] / 1e3, 'share': a['us'] / tot, 'us_per_launch': a['us'] / n,
                   'dram_bytes_per_launch': byts, 'dram_read_per_launch': a['rd'] / n, 'dram_write_per_launch': a['wr'] / n,
                   'dram_gbps': gbs}
    print(f'{a["us"] / 1e3:9.3f} ms {100 * a["us"] / tot:5.1f}%  x{n:<5d} {byts / 1e6:9.2f} MB/launch {gbs:7.0f} GB/s  {k[:70]}')
if len(sys.argv) > 2:
    json.dump(out, open(sys.argv[2], 'w'), indent=1)
