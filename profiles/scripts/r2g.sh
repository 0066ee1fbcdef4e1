#!/bin/bash
mkdir -p gpurun_out
python -c 'import __graft_entry__ as g; g.build(); g.smoke()' > gpurun_out/r2g_smoke.log 2>&1; echo "smoke rc=$?"
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_edge.py tests/test_gpu_reactions.py -x -q > gpurun_out/r2g_pytest.log 2>&1; echo "pytest rc=$?"; tail -n 4 gpurun_out/r2g_pytest.log
timeout 600 python scratch/sweep_target.py c2 > gpurun_out/r2g_target_c2.log 2>&1; echo "sweep rc=$?"
timeout 600 python scratch/sweep_target.py c3 6 > gpurun_out/r2g_target_c3.log 2>&1; echo "sweep c3 rc=$?"
timeout 600 python bench.py --steps 20 --warmup 5 --no_cpu_baseline > gpurun_out/r2g_bench_s20.json 2> gpurun_out/r2g_bench_s20.err; echo "bench rc=$?"
