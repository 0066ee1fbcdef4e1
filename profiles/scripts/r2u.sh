#!/bin/bash
mkdir -p gpurun_out
timeout -k 5 90 python scratch/bisect.py 2 1 0 16 > gpurun_out/r2u_bisect.log 2>&1; echo "bisect rc=$?"; tail -n 3 gpurun_out/r2u_bisect.log
timeout -k 5 200 python scratch/ab_build.py c2 > gpurun_out/r2u_ab_c2.log 2>&1; echo "ab rc=$?"; tail -n 7 gpurun_out/r2u_ab_c2.log
timeout -k 5 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_edge.py tests/test_gpu_reactions.py -x -q > gpurun_out/r2u_pytest.log 2>&1; echo "pytest rc=$?"; tail -n 3 gpurun_out/r2u_pytest.log
CLB_TRACE=1 timeout -k 5 300 python bench.py --steps 20 --warmup 5 --no_cpu_baseline > gpurun_out/r2u_bench_s20.json 2> gpurun_out/r2u_bench_s20.err; echo "bench rc=$?"; grep "clb rebuild" gpurun_out/r2u_bench_s20.err | tail -2
