#!/bin/bash
# Attention kernel with parts switched off (results are garbage, only the time matters):
#   1 = no exp2 (MUFU), 2 = no TMEM loads of S, 4 = no TMEM stores of P, 8 = no P.V MMAs, 16 = no Q.K^T MMAs,
#   32 = no output epilogue (scale, pack, stores), 64 = no TMA loads
for dbg in ${@:-0 1 32 62 63 126 127}; do
  echo "== CLIPEBC_ATTN_DBG=$dbg"; CLIPEBC_ATTN_DBG=$dbg python profiles/attn_bench.py 64 | grep "197+32 impl=3"
done
