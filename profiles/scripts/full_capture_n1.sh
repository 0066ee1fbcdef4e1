#!/bin/bash
timeout 400 python -m pytest tests -m gpu -x -q > gpurun_out/r1g_pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/r1g_pytest_gpu.log; tail -n 3 gpurun_out/r1g_pytest_gpu.log
timeout 60 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -n 2
timeout 400 python bench.py > gpurun_out/r1g_bench_n1.json 2> gpurun_out/r1g_bench_n1.err; echo "bench exit $?"; cut -c1-400 gpurun_out/r1g_bench_n1.json
timeout 300 python bench.py --impl reference > gpurun_out/r1g_bench_reference.json 2> gpurun_out/r1g_bench_reference.err; echo "ref exit $?"; cut -c1-300 gpurun_out/r1g_bench_reference.json
CMD="python bench.py --steps 40 --warmup 20 --equil 40 --no_cpu_baseline --no_e2e"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r1g_launches.csv $CMD > gpurun_out/r1g_ncu_launch.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"k_pair_forces_tab2|k_build_lists|k_bonded|k_integrate" -s 12 -c 8 -f -o gpurun_out/prof_r1g $CMD > gpurun_out/r1g_ncu_full.log 2>&1
ls gpurun_out | grep r1g
