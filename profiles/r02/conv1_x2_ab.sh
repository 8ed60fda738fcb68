#!/bin/bash
# A/B of the conv1 gather (run at the commit that still had the CLIPEBC_C1_OLD / CLIPEBC_C1_RS test knobs): old cell-per-thread kernel vs the column-walking x2 kernel with rs = 1..4
run() { timeout 200 python bench.py --workload windows64 --steps 20 --warmup 5 --no-cpu-baseline --no-gpu-library-baseline 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l)
        ks=d.get('breakdown') or d.get('kernels') or {}
        print('$1', d['value'], d['ms_per_step'], json.dumps(ks)[:0])
        import re
        s=json.dumps(d)
        m=re.search(r'conv1_interp[^}]*}', s)
        print('   ', m.group(0)[:200] if m else 'no conv1 entry')
"; }
CLIPEBC_C1_OLD=1 run old
for rs in 1 2 3 4; do CLIPEBC_C1_RS=$rs run rs$rs; done
CLIPEBC_C1_OLD=1 run old
CLIPEBC_C1_RS=2 run rs2
