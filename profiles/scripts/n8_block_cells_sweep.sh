#!/bin/bash
N=8
for opt in "--option block_cells=10" "--option block_cells=6" "--option block_cells=13"; do
echo "== N=$N $opt"
timeout 100 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29671 bench.py --gpus $N --steps 800 --warmup 200 --equil 600 --no_cpu_baseline --no_e2e $opt 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); r=d['roofline']
        print('steps/s %.1f ms/step %.4f pair_ms %.4f share %.3f rebuilds %s nv %s threads %s'%(d['value'],d['ms_per_step'],r['kernel_ms'],r['kernel_share_of_step'],d['rebuilds'],d.get('pair_nv'),d.get('pair_threads')))
    elif 'rror' in l: print(l.strip())
"
done
