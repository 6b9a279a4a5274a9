#!/bin/bash
# step time of every bench shape under each table-kernel mask (MRGCN_TAB): which backward kernels win where
for shape in "$@"; do
  for m in 1 3 5 7; do
    MRGCN_TAB=$m python bench.py --shape $shape --steps 10 --warmup 3 --no-cpu-baseline > /tmp/ms.json 2> /tmp/ms.err || { tail -3 /tmp/ms.err; continue; }
    python - $shape $m <<'PY'
import json, sys
d = json.loads(open("/tmp/ms.json").read().strip().splitlines()[-1])
top = sorted(d["kernels"].items(), key=lambda kv: -kv[1]["ms_per_step"])[:6]
print(sys.argv[1], "mask", sys.argv[2], "ms/step %.3f" % d["ms_per_step"], " ".join("%s=%.2f" % (k, v["ms_per_step"]) for k, v in top))
PY
  done
done
