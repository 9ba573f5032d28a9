#!/bin/bash
# 8-GPU check of the final code: the driver's launch line for c3 (short), then c4
mkdir -p gpurun_out/r2
O=gpurun_out/r2
nvidia-smi --query-gpu=index,name,driver_version,clocks.max.sm --format=csv > $O/smi8c.csv 2>&1
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 8 --steps 6 --warmup 3 > $O/bench_c3_n8c.json 2> $O/bench_c3_n8c.err; echo "c3 n8 exit=$?"
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus 8 --workload c4 --steps 4 --warmup 3 > $O/bench_c4_n8c.json 2> $O/bench_c4_n8c.err; echo "c4 n8 exit=$?"
for f in $O/bench_c3_n8c.json $O/bench_c4_n8c.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads([l for l in open(sys.argv[1]) if l.startswith('{')][-1])
    print(sys.argv[1],'value',round(d['value']),'e2e',round(d['e2e']['value']),'ms',round(d['ms_per_step'],1),'n',d['n_gpus'],d['clocks'])
except Exception as e: print('  no line',e); print(open(sys.argv[1].replace('.json','.err')).read()[-1500:])
PY
done
