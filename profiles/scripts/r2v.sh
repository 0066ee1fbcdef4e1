#!/bin/bash
mkdir -p gpurun_out
timeout -k 5 200 python -m pytest tests/test_gpu_parity.py -x -q -k "cell_pair or trajectory or reproducible" > gpurun_out/r2v_pytest1.log 2>&1; echo "pytest1 rc=$?"; tail -n 6 gpurun_out/r2v_pytest1.log
timeout -k 5 300 python bench.py --steps 400 --warmup 100 --no_cpu_baseline --no_e2e > gpurun_out/r2v_bench_s400.json 2> gpurun_out/r2v_bench_s400.err; echo "bench rc=$?"
timeout -k 5 300 python bench.py --steps 400 --warmup 100 --no_cpu_baseline --no_e2e --no_parity --option resort_criterion=1 > gpurun_out/r2v_bench_s400_c1.json 2> gpurun_out/r2v_bench_s400_c1.err; echo "bench c1 rc=$?"
timeout -k 5 600 python -m pytest tests -m gpu -x -q > gpurun_out/r2v_pytest_all.log 2>&1; echo "pytest all rc=$?"; tail -n 4 gpurun_out/r2v_pytest_all.log
