#!/bin/bash
mkdir -p gpurun_out
python -c 'import __graft_entry__ as g; g.build(); g.smoke()' > gpurun_out/r3c_smoke.log 2>&1; echo "smoke rc=$?"; tail -n 1 gpurun_out/r3c_smoke.log
timeout -k 5 600 python -m pytest tests -m gpu -x -q > gpurun_out/r3c_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -n 4 gpurun_out/r3c_pytest_gpu.log
timeout -k 5 300 python bench.py --steps 20 --warmup 5 > gpurun_out/r3c_n1_default.json 2> gpurun_out/r3c_n1_default.err; echo "bench default rc=$?"
timeout -k 5 200 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r3c_ref.json 2> gpurun_out/r3c_ref.err; echo "ref rc=$?"
timeout -k 5 400 python bench.py > gpurun_out/r3c_n1_noflags.json 2> gpurun_out/r3c_n1_noflags.err; echo "bench noflags rc=$?"
timeout -k 5 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r3c_launches.csv python bench.py --steps 20 --warmup 5 --no_cpu_baseline --no_e2e --no_parity > gpurun_out/r3c_ncu.log 2>&1; echo "ncu rc=$?"
