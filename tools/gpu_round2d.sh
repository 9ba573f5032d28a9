#!/bin/bash
mkdir -p gpurun_out/r2
O=gpurun_out/r2
timeout 300 python -m pytest tests/test_gpu_api.py tests/test_gpu_hubert.py -m gpu -q -p no:cacheprovider --timeout 280 > $O/tests_d.log 2>&1; echo "tests exit=$?"; tail -3 $O/tests_d.log
timeout 600 python bench.py --workload hs --steps 3 --warmup 3 --no-cpu-baseline > $O/bench_hs.json 2> $O/bench_hs.err; echo "hs exit=$?"; tail -3 $O/bench_hs.err
timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/bench_default3.json 2> $O/bench_default3.err; echo "bench exit=$?"
python - <<'PY'
import json
for f in ('bench_hs','bench_default3'):
    try:
        d=json.loads([l for l in open(f'gpurun_out/r2/{f}.json') if l.startswith('{')][-1])
        print(f,'value',round(d['value']),'e2e',round(d['e2e']['value']),'ms',round(d['ms_per_step'],1),'launches',d['gpu_launches'])
        if 'files_e2e' in d: print('   files_e2e',{k:v for k,v in d['files_e2e'].items() if k!='what'})
    except Exception as e: print(f,'no line',e)
PY
# launch list of one mHuBERT batch
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 600 --csv --log-file $O/launches_hubert.csv python - <<'PY' > $O/ncu_hubert.log 2>&1
import sys; sys.path.insert(0,'.')
import numpy as np, torch, bench
from audiotoken_b200 import packing
from audiotoken_b200.hubert import HubertEncoder, feat_lengths, plan_hubert
dev=torch.device('cuda:0')
enc=HubertEncoder(device='cuda:0', precision='bf16')
lengths=bench.shard_lengths(0,'c3')
rows=np.minimum(np.array([packing.length_tokens(int(n),16000,50) for n in lengths]), int(feat_lengths(480000)))
idx=packing.bucket_by_rows(rows.tolist(),32768)[3]
ln=lengths[idx]; offs=np.zeros(len(idx),dtype=np.int64); offs[1:]=np.cumsum(ln)[:-1]
wave=(0.1*torch.randn(int(ln.sum()),device=dev)).clamp_(-1,1)
plan=plan_hubert(ln,offs,480000,rows[idx])
for _ in range(2): enc.encode_plan(wave,plan,True)
torch.cuda.synchronize(); print('rows',plan.total_rows,'clips',plan.n_clips)
PY
tail -2 $O/ncu_hubert.log
python tools/ncu_traffic.py $O/launches_hubert.csv $O/traffic_hubert.json 2>&1 | head -24
