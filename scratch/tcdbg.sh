for pc in 2 3; do echo "pieces=$pc"; MRGCN_FEAT_TC_PIECES=$pc timeout -k 5 200 python -m pytest tests -m gpu -q 2>&1 | grep -E "passed|failed|max err|timed out|Error" | head -6; done
for d in 0 15; do MRGCN_TC_DEBUG=$d timeout -k 5 150 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | grep -E "timed out|^\{" | head -5 | python -c "
import json,sys
for line in sys.stdin:
    if line.startswith('{'):
        d=json.loads(line); print('dbg=$d', d['ms_per_step'], 'feat_msg_fwd %.3f  bwd_x_msg %.3f ident_msg_fwd %.3f ident_bwd_c %.3f'%(d['kernels']['feat_msg_fwd']['ms_per_step'], d['kernels']['feat_bwd_x_msg']['ms_per_step'], d['kernels']['ident_msg_fwd']['ms_per_step'], d['kernels']['ident_bwd_c']['ms_per_step']))
    else: print(line.strip())"; done
