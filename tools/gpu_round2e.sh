#!/bin/bash
mkdir -p gpurun_out/r2
O=gpurun_out/r2
timeout 300 python -m pytest tests/test_gpu_hubert.py -m gpu -q -p no:cacheprovider --timeout 280 > $O/tests_e.log 2>&1; echo "tests exit=$?"; tail -3 $O/tests_e.log
timeout 600 python bench.py --workload hs --steps 3 --warmup 3 --no-cpu-baseline > $O/bench_hs2.json 2> $O/bench_hs2.err; echo "hs exit=$?"; tail -3 $O/bench_hs2.err
timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/bench_default4.json 2> $O/bench_default4.err; echo "bench exit=$?"
python - <<'PY'
import json
for f in ('bench_hs2','bench_default4'):
    try:
        d=json.loads([l for l in open(f'gpurun_out/r2/{f}.json') if l.startswith('{')][-1])
        print(f,'value',round(d['value']),'e2e',round(d['e2e']['value']),'ms',round(d['ms_per_step'],1),'launches',d['gpu_launches'])
        if 'files_e2e' in d: print('   files_e2e',{k:v for k,v in d['files_e2e'].items() if k!='what'})
    except Exception as e: print(f,'no line',e)
PY
