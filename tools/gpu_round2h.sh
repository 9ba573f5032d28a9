#!/bin/bash
# A/B of the FFN residual epilogue inside one box: c3 bench, no extras
mkdir -p gpurun_out/r2
O=gpurun_out/r2
timeout 400 python -m pytest tests/test_gpu_pipeline.py tests/test_gpu_autocast_parity.py -m gpu -q -x -p no:cacheprovider --timeout 380 > $O/tests_h.log 2>&1; echo "tests exit=$?"; tail -3 $O/tests_h.log
for opt in ffn_resid_epilogue=0 ffn_resid_epilogue=1 ffn_resid_epilogue=0 ffn_resid_epilogue=1; do
B2T_OPTS=$opt timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extras > $O/bench_h.json 2> $O/bench_h.err; echo "$opt bench exit=$?"
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r2/bench_h.json') if l.startswith('{')][-1])
print('  value',round(d['value']),'e2e',round(d['e2e']['value']),'ms',round(d['ms_per_step'],1), d['clocks']['sm_mhz'], {k:round(v,1) for k,v in d['breakdown_ms_per_step'].items()})
PY
done
