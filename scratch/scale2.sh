n=2
t0=$(date +%s)
timeout -k 5 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2952$n bench.py --gpus $n --steps 10 --warmup 3 > gpurun_out/bench_r01_n$n.json 2>gpurun_out/bench_r01_n$n.err; echo "exit=$? wall=$(( $(date +%s) - t0 ))s"
python -c "
import json; d=json.loads(open('gpurun_out/bench_r01_n$n.json').read().strip().splitlines()[-1]); print('N=$n', d['ms_per_step'], d['value'], d['e2e']['ms_per_step'])"
tail -3 gpurun_out/bench_r01_n$n.err
