#!/bin/bash
# usage: tools/bench_short.sh [extra bench args]  -- prints value, ms/step and per-kernel ms from bench.py's JSON line
python bench.py --no-cpu-baseline "$@" 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); k=d['roofline']['kernel_ms_per_step']
        print('value %.0f ms/step %.3f e2e %.0f | fwd %.3f bwd %.3f | other %s' % (d['value'], d['ms_per_step'], d['e2e']['value'], k.get('adm_multislice_fwd',0), k.get('adm_multislice_bwd',0), {a[4:]:round(b,3) for a,b in k.items() if 'multislice' not in a}))
    else: print(l[-300:], end='')
"
