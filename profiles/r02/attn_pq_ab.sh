#!/bin/bash
# A/B of the attention generations (run at the commit that still had the CLIPEBC_ATTN_OLD test knob):
# two chains x one softmax group (12 warps) vs two chains x two groups (20 warps)
run() { timeout 200 python bench.py --workload $2 --steps 20 --warmup 5 --no-cpu-baseline --no-gpu-library-baseline 2>/dev/null | python -c "
import sys,json,re
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l)
        s=json.dumps(d)
        m=re.search(r'\"attention\": \{[^}]*}', s)
        print('$1 $2', round(d['value'],1), d['unit'], round(d['ms_per_step'],4), 'ms/step;', m.group(0)[:120] if m else '')
"; }
for i in 1 2; do
CLIPEBC_ATTN_OLD=1 run old windows64
run new windows64
done
CLIPEBC_ATTN_OLD=1 run old sliding
run new sliding
