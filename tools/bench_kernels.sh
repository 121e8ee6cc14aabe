#!/bin/bash
# prints per-kernel ms for a bench run; usage: tools/bench_kernels.sh [extra bench args]
python bench.py --steps 3 --warmup 3 --no-cpu-baseline "$@" 2>&1 | tail -1 | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: print(l[:800]); continue
    print('value %.1f e2e %.1f nmse %.3e full_scan %s'%(d['value'], d['e2e']['value'], d['check']['nmse_mean'], d.get('full_scan')), ' '.join('%s=%.3f'%(k,v['avg_launch_ms']) for k,v in d['kernels'].items()))
"
