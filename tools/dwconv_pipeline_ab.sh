timeout 100 python tools/dwconv_ab.py --modes 2,7,8 2>&1 | grep "us per launch"
for m in 2 8 7; do
  B2T_OPTS=dwconv_ring=$m timeout 300 python bench.py --no-extras --no-cpu-baseline --steps 3 --warmup 3 2>/dev/null | python -c "
import json,sys
d=json.loads([l for l in sys.stdin if l.startswith('{')][-1])
print('ring=$m value',round(d['value']),'e2e',round(d['e2e']['value']),'ms',round(d['ms_per_step'],1), {k:round(v,1) for k,v in d['breakdown_ms_per_step'].items()}, d['clocks']['sm_mhz'])"
done
