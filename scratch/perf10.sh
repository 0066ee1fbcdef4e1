#!/bin/bash
CLB_BUILD_KERNEL=2 timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_reactions.py tests/test_gpu_edge.py -x -q 2>&1 | tail -n 4
for bk in 1 2; do
echo "== build_kernel=$bk"
CLB_BUILD_KERNEL=$bk CLB_TRACE=1 timeout 120 python bench.py --steps 600 --warmup 100 --equil 500 --no_cpu_baseline --no_e2e 2>&1 | python -c "
import sys,json,re
b=[]
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); r=d['roofline']
        print('steps/s %.1f ms/step %.4f pair_ms %.4f rebuilds %s T %.5f bonds %s'%(d['value'],d['ms_per_step'],r['kernel_ms'],d['rebuilds'],d['temperature'],d['new_bonds_total']))
    elif 'clb rebuild' in l:
        m=re.search(r'build=([0-9.]+)ms',l)
        if m: b.append(float(m.group(1)))
    elif 'rror' in l: print(l.strip())
print('build ms (median of %d): %.3f'%(len(b), sorted(b)[len(b)//2] if b else -1))
"
done
