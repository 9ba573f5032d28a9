#!/bin/bash
# 8-GPU verification of the final code: multi-device tests, then the driver's launch line for c3 and c4
mkdir -p gpurun_out/r2
O=gpurun_out/r2
nvidia-smi --query-gpu=index,name,driver_version,clocks.max.sm --format=csv > $O/smi8b.csv 2>&1
timeout 500 python -m pytest tests/test_gpu_soak.py -m gpu -q -x -p no:cacheprovider --timeout 450 -k "second_device or two_rank or trap" > $O/tests_multi.log 2>&1; echo "multi tests exit=$?"; tail -3 $O/tests_multi.log
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 8 --steps 10 --warmup 3 > $O/bench_c3_n8b.json 2> $O/bench_c3_n8b.err; echo "c3 n8 exit=$?"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus 8 --workload c4 --steps 5 --warmup 3 > $O/bench_c4_n8b.json 2> $O/bench_c4_n8b.err; echo "c4 n8 exit=$?"
for f in $O/bench_c3_n8b.json $O/bench_c4_n8b.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads([l for l in open(sys.argv[1]) if l.startswith('{')][-1])
    print(sys.argv[1],'value',round(d['value']),'e2e',round(d['e2e']['value']),'ms',round(d['ms_per_step'],1),'n',d['n_gpus'],d['clocks'])
except Exception as e: print('  no line',e); print(open(sys.argv[1].replace('.json','.err')).read()[-1500:])
PY
done
