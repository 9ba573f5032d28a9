#!/bin/bash
# developer A/B: attention sweep over variant builds of attention_tc.cu (tools/build_variants.py)
V=audiotoken_b200/lib/variants
for v in ${VARIANTS:-base}; do
  lp=""; [ $v != base ] && lp=$V/libb200tok_$v.so
  B2T_LIB_PATH=$lp timeout 120 python tools/attn_sweep.py 2>&1 | grep "p-in-tmem" | grep -v " vs " | sed "s/^/[$v] /"
done
