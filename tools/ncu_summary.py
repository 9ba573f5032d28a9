"""Compact per-kernel summary of an .ncu-rep (run where ncu is installed; no GPU needed).

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep > profiles/rNN_<what>.md
"""
import csv
import io
import subprocess
import sys

KEYS = [
    ('gpu__time_duration.sum', 'duration'),
    ('dram__bytes_read.sum', 'dram read'),
    ('dram__bytes_write.sum', 'dram write'),
    ('dram__throughput.avg.pct_of_peak_sustained_elapsed', 'dram %peak'),
    ('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'dram %peak'),
    ('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed', 'tensor pipe %'),
    ('sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active', 'hmma %'),
    ('sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm %'),
    ('smsp__issue_active.avg.pct_of_peak_sustained_active', 'issue active %'),
    ('sm__warps_active.avg.pct_of_peak_sustained_active', 'warps active %'),
    ('launch__registers_per_thread', 'regs'),
    ('launch__grid_size', 'grid'),
    ('launch__block_size', 'block'),
    ('lts__throughput.avg.pct_of_peak_sustained_elapsed', 'L2 %'),
    ('l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'L1 %'),
]


def main(path):
    out = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    print(f'# ncu summary of `{path}`\n')
    print('| kernel | ' + ' | '.join(dict.fromkeys(n for _, n in KEYS)) + ' |')
    print('|---|' + '---|' * len(dict.fromkeys(n for _, n in KEYS)))
    for r in rows[2:]:
        rec = dict(zip(hdr, r))
        unit = dict(zip(hdr, units))
        name = rec.get('Kernel Name', '?')
        name = name.split('(')[0].replace('<unnamed>::', '').replace('void ', '')[:70]
        cells = {}
        for k, n in KEYS:
            if k in rec and rec[k] != '' and n not in cells:
                v = rec[k]
                try:
                    f = float(v.replace(',', ''))
                    v = f'{f:.4g}'
                except ValueError:
                    pass
                u = unit.get(k, '')
                cells[n] = f'{v} {u}'.strip() if u not in ('%', '') else v
        print(f'| `{name}` | ' + ' | '.join(cells.get(n, '') for n in dict.fromkeys(n for _, n in KEYS)) + ' |')


if __name__ == '__main__':
    main(sys.argv[1])
