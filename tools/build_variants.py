#!/usr/bin/env python
"""Developer A/B builds of libb200tok.so: every variant relinks the current objects with one translation unit
rebuilt from another source file and/or with extra -D flags.  Select one at run time with B2T_LIB_PATH.

    python tools/build_variants.py name:unit.cu[:source path][:-DFLAG=1,...] ...
"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from audiotoken_b200 import build as B  # noqa: E402


def main():
    B.build()
    out_dir = os.path.join(B.LIBDIR, 'variants')
    os.makedirs(out_dir, exist_ok=True)
    for spec in sys.argv[1:]:
        parts = spec.split(':')
        name, unit = parts[0], parts[1]
        src = parts[2] if len(parts) > 2 and parts[2] else os.path.join(B.CSRC, unit)
        defs = parts[3].split(',') if len(parts) > 3 and parts[3] else []
        obj = os.path.join(out_dir, f'{name}_{unit[:-3]}.o')
        cmd = [B.NVCC] + B.FLAGS + ['-I', B.CSRC] + defs + ['-c', src, '-o', obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            print(f'{name}: nvcc failed\n{r.stdout}\n{r.stderr}')
            continue
        objs = [os.path.join(B.OBJDIR, s[:-3] + '.o') for s in B.sources() if s != unit] + [obj]
        lib = os.path.join(out_dir, f'libb200tok_{name}.so')
        r = subprocess.run([B.NVCC, '-shared', '-gencode', 'arch=compute_100a,code=sm_100a', '-o', lib] + objs,
                           capture_output=True, text=True)
        print(name, 'ok' if r.returncode == 0 else r.stderr, lib)


if __name__ == '__main__':
    main()
