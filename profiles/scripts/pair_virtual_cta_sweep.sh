#!/bin/bash
timeout 200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_reactions.py -x -q 2>&1 | tail -n 3
for opt in "" "--option pair_nv=1" "--option pair_nv=2" "--option pair_nv=3" "--option pair_nv=4" "--option pair_nv=5" "--option pair_nv=3 --option pair_ni=2"; do
  echo "== $opt"
  timeout 120 python bench.py --steps 400 --warmup 100 --equil 500 --no_cpu_baseline --no_e2e $opt 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); r=d['roofline']
        print('steps/s %.1f ms/step %.4f pair_ms %.4f share %.3f threads %s grid %s nv %s smem %s rebuilds %s T %.4f'%(d['value'],d['ms_per_step'],r['kernel_ms'],r['kernel_share_of_step'],d.get('pair_threads'),d.get('pair_grid'),d.get('pair_nv'),d.get('pair_smem'),d['rebuilds'],d['temperature']))
    elif 'rror' in l: print(l.strip())
"
done
