#!/bin/bash
N=$1
for opt in "--option comm_group=1" "--option comm_group=0" "--option overlap_halo=1"; do
echo "== N=$N $opt"
timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29671 bench.py --gpus $N --steps 1000 --warmup 200 --no_cpu_baseline --no_e2e $opt 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); r=d['roofline']
        print('steps/s %.1f ms/step %.4f pair_ms %.4f share %.3f rebuilds %s T %.4f events %s'%(d['value'],d['ms_per_step'],r['kernel_ms'],r['kernel_share_of_step'],d['rebuilds'],d['temperature'],d['reaction_events']))
    elif 'rror' in l: print(l.strip())
"
done
