#!/bin/bash
V=audiotoken_b200/lib/variants
for v in base nomax noresc oneacc all3; do
  lp=""; [ $v != base ] && lp=$V/libb200tok_$v.so
  B2T_LIB_PATH=$lp timeout 120 python tools/attn_sweep.py 2>&1 | grep "single-pass:\|two-pass:" | grep "24x1500\|64x500" | sed "s/^/[$v] /"
done
