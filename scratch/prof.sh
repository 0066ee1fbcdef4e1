#!/bin/bash
set -x
CMD="python bench.py --steps 40 --warmup 20 --equil 40 --no_cpu_baseline --no_e2e"
timeout 120 $CMD > gpurun_out/r1f_plain.log 2>&1 || exit 1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r1f_launches.csv $CMD > gpurun_out/r1f_ncu_launch.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"k_pair_forces_tab2|k_build_lists|k_bonded" -s 9 -c 6 -f -o gpurun_out/prof_r1f $CMD > gpurun_out/r1f_ncu_full.log 2>&1
ls -la gpurun_out/
