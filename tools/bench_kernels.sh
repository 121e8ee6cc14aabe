#!/bin/bash
# prints per-kernel ms for a device-resident bench run; usage: tools/bench_kernels.sh [extra bench args]
python bench.py --steps 3 --warmup 3 --kernels-only "$@" 2>&1 | tail -1 | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: print(l[:800]); continue
    print('value %.1f nmse %.3e parity %s'%(d['value'], d['check']['nmse_mean'], (d.get('parity') or {}).get('max_rel_theta')), ' '.join('%s=%.3f'%(k,v['avg_launch_ms']) for k,v in d['kernels'].items()))
"
