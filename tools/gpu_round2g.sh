#!/bin/bash
# GEMM change check: kernel tests + soak + c3 bench
mkdir -p gpurun_out/r2
O=gpurun_out/r2
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_pipeline.py tests/test_gpu_soak.py -m gpu -q -x -p no:cacheprovider --timeout 500 > $O/tests_g.log 2>&1; echo "tests exit=$?"; tail -3 $O/tests_g.log
timeout 300 python tools/kernel_soak.py --kernel gemm --iters 300 2>&1 | tail -6
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extras > $O/bench_g.json 2> $O/bench_g.err; echo "bench exit=$?"; tail -2 $O/bench_g.err
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r2/bench_g.json') if l.startswith('{')][-1])
print('value',round(d['value']),'e2e',round(d['e2e']['value']),'ms',round(d['ms_per_step'],1),'launches',d['gpu_launches'], d['clocks'])
print(d['breakdown_ms_per_step']); print(d['roofline']['achieved'], d['roofline']['frac'])
PY
