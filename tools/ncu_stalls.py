"""Developer tool: per-instruction stall samples of one kernel in an .ncu-rep (source page), in program order.

    python tools/ncu_stalls.py gpurun_out/x.ncu-rep [min_samples] [first_addr_suffix last_addr_suffix]
"""
import csv, io, subprocess, sys
path = sys.argv[1]
mins = int(sys.argv[2]) if len(sys.argv) > 2 else 60
raw = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
h, v = rows[0], rows[2]
want = ['gpu__time_duration.sum', 'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum', 'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'launch__registers_per_thread', 'dram__bytes_read.sum', 'dram__bytes_write.sum']
for n, x in zip(h, v):
    if n in want: print(f'{n}: {x}')
src = subprocess.run(['ncu', '-i', path, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr = rows[1]; ix = {n: i for i, n in enumerate(hdr)}
data = rows[2:]
stalls = [n for n in hdr if n.startswith('stall_') and 'Not Issued' not in n]
if len(sys.argv) > 4:
    a = [i for i, r in enumerate(data) if r[ix['Address']].endswith(sys.argv[3])][0]
    b = [i for i, r in enumerate(data) if r[ix['Address']].endswith(sys.argv[4])][0]
    data = data[a:b + 1]
tot = sum(int(r[ix['# Samples']] or 0) for r in data)
print('instructions', len(data), 'samples', tot)
agg = {n: sum(int(r[ix[n]] or 0) for r in data) for n in stalls}
print(sorted(((k, v) for k, v in agg.items() if v), key=lambda kv: -kv[1]))
cum = 0
for r in data:
    s = int(r[ix['# Samples']] or 0); cum += s
    if s >= mins:
        st = {n: int(r[ix[n]] or 0) for n in stalls}; big = sorted(st.items(), key=lambda kv: -kv[1])[:2]
        print(f"{r[ix['Address']][-5:]} {s:5d} cum={100 * cum / max(tot, 1):5.1f}% ex={r[ix['Instructions Executed']]:>8} {r[ix['Source']][:72]:72s} {big}")
