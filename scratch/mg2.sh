#!/bin/bash
timeout 100 python -m pytest tests/test_gpu_reactions.py tests/test_gpu_parity.py -x -q 2>&1 | tail -n 3
MGPU_NSIDE=18 timeout 100 python -u -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29655 tests/mgpu_worker.py 2>&1 | grep -E "MGPU_OK|Error|error|assert" | tail -n 6
for opt in "--option overlap_halo=1" "--option overlap_halo=0"; do
echo "== $opt"
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29671 bench.py --gpus 2 --steps 1000 --warmup 200 --no_cpu_baseline --no_e2e $opt 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); r=d['roofline']
        print('steps/s %.1f ms/step %.4f pair_ms %.4f share %.3f rebuilds %s T %.4f events %s'%(d['value'],d['ms_per_step'],r['kernel_ms'],r['kernel_share_of_step'],d['rebuilds'],d['temperature'],d['reaction_events']))
    elif 'rror' in l: print(l.strip())
"
done
