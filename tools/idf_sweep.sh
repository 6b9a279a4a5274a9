#!/bin/bash
# Sweep of the tile geometry of the fused identity backward (csrc/ident_bwd.cu) on the AM shape; every argument is a
# comma-separated list of NAME=VALUE environment settings (MRGCN_IDF_TJ / _SV / _SM / _MCAP); prints ms of the kernels.
# Stops at the first configuration that fails.
mkdir -p gpurun_out
for cfg in "${@}"; do
  env $(echo "$cfg" | tr ',' ' ') timeout 200 python bench.py --no-cpu-baseline --steps 10 --warmup 3 2>gpurun_out/sweep_last.err | grep -v "^mrgcn: mbar" > gpurun_out/sweep_last.json
  python - "$cfg" <<'PY' || { echo "$cfg FAILED"; tail -c 300 gpurun_out/sweep_last.err; exit 1; }
import json, sys
d = json.loads(open('gpurun_out/sweep_last.json').read().strip().splitlines()[-1]); k = d['kernels']
print(sys.argv[1], 'step %.3f' % d['ms_per_step'], ' '.join('%s %.3f' % (n, k[n]['ms_per_step']) for n in ('ident_bwd_fused', 'comp_chunk_reduce', 'comp_reduce') if n in k))
PY
done
