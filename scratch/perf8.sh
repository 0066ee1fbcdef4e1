#!/bin/bash
timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_reactions.py tests/test_gpu_surface.py -x -q 2>&1 | tail -n 3
timeout 120 python bench.py --steps 1000 --warmup 200 --no_cpu_baseline --no_e2e 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); r=d['roofline']
        print('steps/s %.1f ms/step %.4f pair_ms %.4f share %.3f rebuilds %s T %.4f'%(d['value'],d['ms_per_step'],r['kernel_ms'],r['kernel_share_of_step'],d['rebuilds'],d['temperature']))
    elif 'rror' in l: print(l.strip())
"
