#!/bin/bash
# Which resource bounds the CTA-pair mainloop? Experiment knobs (results are garbage, only the time matters):
#   2 = no epilogue math/stores, 6 = no TMEM reads either, 8 = no weight loads, 16 = no activation loads
export CLIPEBC_GEMM_MC=1
for dbg in 0 6 14 22 30; do
  export CLIPEBC_GEMM_DBG=$dbg; echo "== DBG=$dbg"
  python profiles/gemm_one_bench.py 2 12608 2304 768 2 256
  python profiles/gemm_one_bench.py 2 113664 2304 768 2 256
  python profiles/gemm_one_bench.py 2 57600 768 6912 2 256
done
