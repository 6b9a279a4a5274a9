for n in 8 2; do
timeout -k 5 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 10 --warmup 3 > gpurun_out/bench_r01_n$n.json 2>gpurun_out/bench_r01_n$n.err
python -c "
import json; d=json.loads(open('gpurun_out/bench_r01_n$n.json').read().strip().splitlines()[-1]); print('N=$n', d['ms_per_step'], d['value'], d['e2e']['ms_per_step']); tot=0
for k,v in sorted(d['kernels'].items(), key=lambda kv:-kv[1]['ms_per_step'])[:12]: print('%-20s %7.3f ms'%(k, v['ms_per_step']))
print('sum kernels', sum(v['ms_per_step'] for v in d['kernels'].values()))
"; tail -2 gpurun_out/bench_r01_n$n.err; done
