#!/bin/bash
# quick N=1 perf matrix: bucket timers, then pair_warps variants
for opt in "--option timers=1" "--option pair_warps=6" "--option pair_warps=5" ""; do
  echo "== $opt"
  timeout 120 python bench.py --steps 600 --warmup 100 --equil 600 --no_cpu_baseline --no_e2e $opt 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); r=d['roofline']
        print('steps/s %.1f ms/step %.4f pair_ms %.4f share %.3f threads %s grid %s home_max %s rebuilds %s buckets %s'%(d['value'],d['ms_per_step'],r['kernel_ms'],r['kernel_share_of_step'],d.get('pair_threads'),d.get('pair_grid'),d.get('home_max'),d['rebuilds'],d.get('buckets_s')))
    elif 'rror' in l: print(l.strip())
"
done
