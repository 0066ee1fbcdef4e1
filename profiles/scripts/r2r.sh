#!/bin/bash
mkdir -p gpurun_out
timeout -k 5 90 python scratch/bisect.py 2 1 0 16 > gpurun_out/r2r_bisect.log 2>&1; echo "bisect rc=$?"; tail -n 3 gpurun_out/r2r_bisect.log
timeout -k 5 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_edge.py tests/test_gpu_reactions.py -x -q > gpurun_out/r2r_pytest.log 2>&1; echo "pytest rc=$?"; tail -n 3 gpurun_out/r2r_pytest.log
timeout -k 5 200 python scratch/ab_build.py c2 > gpurun_out/r2r_ab_c2.log 2>&1; echo "ab rc=$?"; tail -n 7 gpurun_out/r2r_ab_c2.log
CLB_TRACE=1 timeout -k 5 300 python bench.py --steps 400 --warmup 100 --no_cpu_baseline > gpurun_out/r2r_bench_s400.json 2> gpurun_out/r2r_bench_s400.err; echo "bench rc=$?"
