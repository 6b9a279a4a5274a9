timeout -k 5 400 python -m pytest tests -m gpu -q 2>&1 | grep -E "passed|failed|FAILED|max err|timed out|Error" | head -20
timeout -k 5 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r01_l.json 2> gpurun_out/bench_r01_l.err
python -c "
import json; d=json.load(open('gpurun_out/bench_r01_l.json')); print(d['ms_per_step'], d['value'], d['e2e']['ms_per_step'])
for k,v in sorted(d['kernels'].items(), key=lambda kv:-kv[1]['ms_per_step'])[:10]: print('%-20s %7.3f ms %s'%(k, v['ms_per_step'], v['gbps']))
"; tail -2 gpurun_out/bench_r01_l.err
