for m in 2 6; do
  B2T_OPTS=dwconv_ring=$m timeout 300 python bench.py --no-extras --no-cpu-baseline --steps 3 --warmup 3 2>/dev/null | python -c "
import json,sys
d=json.loads([l for l in sys.stdin if l.startswith('{')][-1])
print('ring=$m value',round(d['value']),'e2e',round(d['e2e']['value']),'ms',round(d['ms_per_step'],1), {k:round(v,1) for k,v in d['breakdown_ms_per_step'].items()}, d['clocks']['sm_mhz'])"
done
timeout 300 python bench.py --workload c4 --no-cpu-baseline --steps 3 --warmup 3 2>/dev/null | python -c "
import json,sys
d=json.loads([l for l in sys.stdin if l.startswith('{')][-1])
print('c4 value',round(d['value']),'e2e',round(d['e2e']['value']),'ms',round(d['ms_per_step'],1), round(d['e2e']['ms_per_step'],1))"
