#!/bin/bash
# Final 1-GPU capture of round 2: full GPU suite, smoke, the driver's default bench line, the reference arm, the ncu launch list
mkdir -p gpurun_out/r2
O=gpurun_out/r2
nvidia-smi --query-gpu=index,name,driver_version,clocks.max.sm --format=csv > $O/final_smi.csv 2>&1
timeout 1200 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout 900 > $O/final_tests.log 2>&1; echo "tests exit=$?"; tail -3 $O/final_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/final_smoke.log 2>&1; echo "smoke exit=$?"; tail -4 $O/final_smoke.log
timeout 900 python bench.py > $O/bench_final.json 2> $O/bench_final.err; echo "bench exit=$?"; tail -2 $O/bench_final.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_final_reference.json 2> $O/bench_final_reference.err; echo "reference arm exit=$?"
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r2/bench_final.json') if l.startswith('{')][-1])
print('value',round(d['value']),'e2e',round(d['e2e']['value']),'ms',round(d['ms_per_step'],1),'launches',d['gpu_launches'], d['clocks'])
print(d['breakdown_ms_per_step']); print('roofline', d['roofline']['achieved'], d['roofline']['frac'], 'cpu', d.get('cpu_baseline',{}).get('value'))
print('acoustic', round(d['acoustic_c4']['value']), round(d['acoustic_c4']['e2e']['value']), 'files', round(d['files_e2e']['value']))
r=json.loads([l for l in open('gpurun_out/r2/bench_final_reference.json') if l.startswith('{')][-1]); print('reference arm', r['value'], r['cpu_baseline']['cores'])
PY
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 2500 --csv --log-file $O/launches_final.csv python tools/fault_hunt.py --passes 1 --max-batches 1 > $O/ncu_final.log 2>&1; echo "ncu exit=$?"; tail -2 $O/ncu_final.log
python tools/ncu_traffic.py $O/launches_final.csv $O/traffic_all_final.json | head -16
