#!/bin/bash
# round 2, session 2: full GPU suite + c3 bench with the P-in-TMEM attention as default
mkdir -p gpurun_out/r2
O=gpurun_out/r2
timeout 900 python -m pytest tests -m gpu -q -x -p no:cacheprovider --timeout 600 > $O/tests_f.log 2>&1; echo "tests exit=$?"; tail -4 $O/tests_f.log
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $O/bench_f.json 2> $O/bench_f.err; echo "bench exit=$?"; tail -2 $O/bench_f.err
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r2/bench_f.json') if l.startswith('{')][-1])
print('value',round(d['value']),'e2e',round(d['e2e']['value']),'ms',round(d['ms_per_step'],1),'launches',d['gpu_launches'], d['clocks'])
print(d['breakdown_ms_per_step'])
PY
