"""Top stall locations of an .ncu-rep (SASS level): python tools/ncu_hot.py rep [N]"""
import csv, io, subprocess, sys
out = subprocess.run(['ncu', '-i', sys.argv[1], '--page', 'source', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hi = next(i for i, r in enumerate(rows) if r and r[0] == 'Address')
h = rows[hi]
ci, si, ei = h.index('Warp Stall Sampling (All Samples)'), h.index('Source'), h.index('Instructions Executed')
stall_cols = [(i, c) for i, c in enumerate(h) if c.startswith('stall_') and 'Not Issued' not in c]
data = []
for k, r in enumerate(rows[hi + 1:]):
    try: v = float(r[ci])
    except (ValueError, IndexError): continue
    top = max(stall_cols, key=lambda ic: float(r[ic[0]] or 0))
    data.append((v, k, r[si].strip()[:90], r[ei], top[1]))
tot = sum(d[0] for d in data)
n = int(sys.argv[2]) if len(sys.argv) > 2 else 25
print(f'total samples {tot:.0f}, {len(data)} instructions')
for v, k, s, e, t in sorted(data, reverse=True)[:n]:
    print(f'{100*v/tot:5.1f}%  #{k:<5d} exec {e:>9}  {t:<22} {s}')
