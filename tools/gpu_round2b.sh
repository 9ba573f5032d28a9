#!/bin/bash
mkdir -p gpurun_out/r2
O=gpurun_out/r2
timeout 1200 python -m pytest tests/test_gpu_autocast_parity.py tests/test_gpu_api.py -m gpu -q -s -p no:cacheprovider --timeout 900 > $O/parity1.log 2>&1; echo "parity exit=$?"
grep -v "^$" $O/parity1.log | tail -45
timeout 600 python bench.py --steps 3 --warmup 3 > $O/bench_default.json 2> $O/bench_default.err; echo "bench exit=$?"; tail -3 $O/bench_default.err
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r2/bench_default.json') if l.startswith('{')][-1])
print('value',round(d['value']),'e2e',round(d['e2e']['value']))
print('acoustic_c4',{k:(round(v) if isinstance(v,float) else v) for k,v in d.get('acoustic_c4',{}).items() if k in('value','ms_per_step')}, round(d.get('acoustic_c4',{}).get('e2e',{}).get('value',0)))
print('files_e2e',{k:v for k,v in d.get('files_e2e',{}).items() if k!='what'})
PY
# launch list with DRAM bytes: one production batch (49 clips, 65 496 rows, 19 layers)
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 1200 --csv --log-file $O/launches_c3_batch3.csv python tools/fault_hunt.py --passes 1 --max-batches 4 > $O/ncu_launches.log 2>&1; echo "ncu exit=$?"; tail -2 $O/ncu_launches.log
python tools/ncu_traffic.py $O/launches_c3_batch3.csv $O/traffic_all.json | head -30
