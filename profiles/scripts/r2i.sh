#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_edge.py tests/test_gpu_reactions.py -x -q > gpurun_out/r2i_pytest.log 2>&1; echo "pytest rc=$?"; tail -n 4 gpurun_out/r2i_pytest.log
timeout 600 python scratch/ab_build.py c2 > gpurun_out/r2i_ab_c2.log 2>&1; echo "ab rc=$?"; cat gpurun_out/r2i_ab_c2.log | tail -8
timeout 600 python scratch/ab_build.py c5 10 > gpurun_out/r2i_ab_c5.log 2>&1; echo "ab c5 rc=$?"; cat gpurun_out/r2i_ab_c5.log | tail -8
CLB_TRACE=1 timeout 300 python bench.py --steps 20 --warmup 5 --no_cpu_baseline > gpurun_out/r2i_bench_s20.json 2> gpurun_out/r2i_bench_s20.err; echo "bench rc=$?"
