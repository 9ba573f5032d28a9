#!/bin/bash
mkdir -p gpurun_out/r2
O=gpurun_out/r2
timeout 900 python -m pytest tests/test_gpu_hubert.py tests/test_gpu_api.py tests/test_gpu_kernels.py -m gpu -q -p no:cacheprovider --timeout 600 > $O/tests_c.log 2>&1; echo "tests exit=$?"; tail -4 $O/tests_c.log
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $O/bench_default2.json 2> $O/bench_default2.err; echo "bench exit=$?"; tail -3 $O/bench_default2.err
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r2/bench_default2.json') if l.startswith('{')][-1])
print('value',round(d['value']),'e2e',round(d['e2e']['value']), 'breakdown', {k:round(v,1) for k,v in d['breakdown_ms_per_step'].items()})
print('acoustic_c4',round(d['acoustic_c4']['value']), round(d['acoustic_c4']['e2e']['value']))
print('files_e2e',{k:v for k,v in d.get('files_e2e',{}).items() if k!='what'})
print('traffic', d['roofline'].get('traffic'))
PY
