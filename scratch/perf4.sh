#!/bin/bash
for opt in "--option tables_in_smem=0" "--option block_cells=7" "--option block_cells=5" "--option block_cells=4"; do
  echo "== $opt"
  timeout 120 python bench.py --steps 400 --warmup 100 --equil 500 --no_cpu_baseline --no_e2e $opt 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); r=d['roofline']
        print('steps/s %.1f ms/step %.4f pair_ms %.4f share %.3f threads %s grid %s home_max %s tile_max %s rebuilds %s T %.4f'%(d['value'],d['ms_per_step'],r['kernel_ms'],r['kernel_share_of_step'],d.get('pair_threads'),d.get('pair_grid'),d.get('home_max'),d.get('tile_max'),d['rebuilds'],d['temperature']))
    elif 'rror' in l: print(l.strip())
"
done
