#!/bin/bash
# A/B of the cluster-of-4 weight-multicast variant (CLIPEBC_GEMM_MC=2) against the plain CTA-pair kernel (=1),
# hot-path GEMM shapes at 64 windows, same box, same run.   usage: bash profiles/gemm_mc_ab.sh
for mc in 1 2; do
  export CLIPEBC_GEMM_MC=$mc; echo "== MC=$mc"
  for bn in 256 192; do
    python profiles/gemm_one_bench.py 2 12608 2304 768 2 $bn
    python profiles/gemm_one_bench.py 2 12608 768 768 4 $bn
    python profiles/gemm_one_bench.py 2 12608 3072 768 3 $bn
    python profiles/gemm_one_bench.py 2 12608 768 3072 4 $bn
    python profiles/gemm_one_bench.py 2 57600 768 6912 2 $bn
  done
  python profiles/gemm_one_bench.py 2 57600 512 2304 1 256
done
