#!/bin/bash
for opt in "" "--option sync_chunk=3" "--option sync_chunk=8" "--option sync_chunk=12" "--option sync_chunk=24"; do
  echo "== $opt"
  timeout 120 python bench.py --steps 1000 --warmup 200 --no_cpu_baseline --no_e2e $opt 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); r=d['roofline']
        print('steps/s %.1f ms/step %.4f pair_ms %.4f launches %s rebuilds %s'%(d['value'],d['ms_per_step'],r['kernel_ms'],d['gpu_launches'],d['rebuilds']))
    elif 'rror' in l: print(l.strip())
"
done
