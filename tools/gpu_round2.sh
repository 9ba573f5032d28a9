#!/bin/bash
# Round-2 verification on a 2-GPU box: the whole GPU suite (soak tests included) on GPU 0+1, then benches.
mkdir -p gpurun_out/r2
O=gpurun_out/r2
timeout 1500 python -m pytest tests -m gpu -x -q -p no:cacheprovider --timeout 900 > $O/pytest_gpu.log 2>&1; echo "pytest exit=$?"; tail -5 $O/pytest_gpu.log
( CUDA_VISIBLE_DEVICES=0 timeout 300 python bench.py --steps 5 --warmup 3 > $O/bench_c3_n1.json 2> $O/bench_c3_n1.err; echo "bench c3 exit=$?" ) &
( CUDA_VISIBLE_DEVICES=1 timeout 300 python bench.py --workload c4 --steps 5 --warmup 3 > $O/bench_c4_n1.json 2> $O/bench_c4_n1.err; echo "bench c4 exit=$?" ) &
wait
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 10 --warmup 3 > $O/bench_c3_n2.json 2> $O/bench_c3_n2.err; echo "bench n2 exit=$?"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 2 --workload c4 --steps 10 --warmup 3 > $O/bench_c4_n2.json 2> $O/bench_c4_n2.err; echo "bench c4 n2 exit=$?"
for f in $O/bench_*.json; do echo $f; python - "$f" <<'PY'
import json,sys
try:
    d=json.loads([l for l in open(sys.argv[1]) if l.startswith('{')][-1])
    print('  value',round(d['value']),'e2e',round(d['e2e']['value']),'ms',round(d['ms_per_step'],1),'n',d['n_gpus'],'breakdown',{k:round(v,1) for k,v in d.get('breakdown_ms_per_step',{}).items()})
except Exception as e: print('  no line',e)
PY
done
