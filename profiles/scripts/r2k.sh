#!/bin/bash
mkdir -p gpurun_out
timeout -k 5 200 python scratch/ab_build.py c2 > gpurun_out/r2k_ab_c2.log 2>&1; echo "ab rc=$?"; tail -n 8 gpurun_out/r2k_ab_c2.log
timeout -k 5 400 python -m pytest tests/test_gpu_parity.py tests/test_gpu_edge.py tests/test_gpu_reactions.py -x -q > gpurun_out/r2k_pytest.log 2>&1; echo "pytest rc=$?"; tail -n 4 gpurun_out/r2k_pytest.log
timeout -k 5 400 python scratch/sweep_pipe.py c2 > gpurun_out/r2k_pipe_c2.log 2>&1; echo "pipe sweep rc=$?"
timeout -k 5 300 python scratch/ab_build.py c5 10 > gpurun_out/r2k_ab_c5.log 2>&1; echo "ab c5 rc=$?"; tail -n 8 gpurun_out/r2k_ab_c5.log
CLB_TRACE=1 timeout -k 5 300 python bench.py --steps 20 --warmup 5 --no_cpu_baseline > gpurun_out/r2k_bench_s20.json 2> gpurun_out/r2k_bench_s20.err; echo "bench rc=$?"
