#!/bin/bash
# per-kernel table of one AM step under a given MRGCN_TAB mask: tools/bench_kernels.sh <mask> [extra bench args]
m=$1; shift
MRGCN_TAB=$m python bench.py --steps 8 --warmup 3 --no-cpu-baseline "$@" 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('MRGCN_TAB=$m ms/step %.3f e2e %.2f' % (d['ms_per_step'], d['e2e']['ms_per_step']))
for k,v in sorted(d['kernels'].items(), key=lambda kv:-kv[1]['ms_per_step'])[:11]: print('   %-22s %d %.3f'%(k,v['launches_per_step'],v['ms_per_step']))
"
