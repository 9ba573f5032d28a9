"""Aggregate an ncu `--metrics gpu__time_duration.sum --csv` launch list by kernel name."""
import csv, sys, collections
rows = [r for r in csv.reader(open(sys.argv[1], errors='replace')) if len(r) > 5]
hdr = next(i for i, r in enumerate(rows) if 'Kernel Name' in r)
h = rows[hdr]; ki, vi, ui = h.index('Kernel Name'), h.index('Metric Value'), h.index('Metric Unit')
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows[hdr + 1:]:
    try: v = float(r[vi].replace(',', ''))
    except ValueError: continue
    scale = {'ns': 1e-3, 'us': 1.0, 'ms': 1e3, 's': 1e6}.get(r[ui].strip(), 1e-3)
    name = r[ki].split('(')[0].replace('void ', '').replace('(anonymous namespace)::', '')[:60]
    agg[name][0] += 1; agg[name][1] += v * scale
tot = sum(v[1] for v in agg.values())
print(f'total {tot/1e3:.2f} ms over {sum(v[0] for v in agg.values())} launches')
for k, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f'{us/1e3:10.3f} ms  {100*us/tot:5.1f}%  x{n:<6d} {k}')
