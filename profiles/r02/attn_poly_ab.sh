#!/bin/bash
# A/B of the FMA-pipe share of the softmax exponentials (library built with make EXTRA=-DCLIPEBC_ATTN_POLY_AB).
for p in 0 4 3 2 0 4; do echo "== CLIPEBC_ATTN_POLY=$p"; CLIPEBC_ATTN_POLY=$p python profiles/attn_bench.py 64; done
for p in 4 2; do echo "== tests CLIPEBC_ATTN_POLY=$p"; CLIPEBC_ATTN_POLY=$p python -m pytest tests/test_kernels_gpu.py -q -k attention 2>&1 | tail -3; done
