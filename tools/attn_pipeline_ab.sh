#!/bin/bash
# attention options inside the power-capped pipeline (same box, back to back)
for o in "attn_poly_exp=1" "attn_poly_exp=0" "attn_two_pass=6" "attn_poly_exp=1"; do
  B2T_OPTS=$o timeout 300 python bench.py --no-extras --no-cpu-baseline --steps 3 --warmup 3 2>/dev/null | python -c "
import json,sys
d=json.loads([l for l in sys.stdin if l.startswith('{')][-1])
print('$o value',round(d['value']),'e2e',round(d['e2e']['value']),'ms',round(d['ms_per_step'],1), {k:round(v,1) for k,v in d['breakdown_ms_per_step'].items()}, d['clocks']['sm_mhz'])"
done
