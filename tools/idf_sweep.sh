#!/bin/bash
# Sweep of the tile geometry of the fused identity backward (csrc/ident_bwd.cu) on the AM shape; prints ms of the kernels.
mkdir -p gpurun_out
for cfg in "${@}"; do
  IFS=, read tj s mcap <<< "$cfg"
  MRGCN_IDF_TJ=$tj MRGCN_IDF_S=$s MRGCN_IDF_MCAP=$mcap timeout 300 python bench.py --no-cpu-baseline --steps 10 --warmup 3 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); k=d['kernels']
print('$cfg', 'step %.3f' % d['ms_per_step'], ' '.join('%s %.3f' % (n, k[n]['ms_per_step']) for n in ('ident_bwd_fused','comp_chunk_reduce','comp_reduce') if n in k))
"
done
