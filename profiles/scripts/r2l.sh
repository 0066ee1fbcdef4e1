#!/bin/bash
mkdir -p gpurun_out
CMD="python bench.py --steps 20 --warmup 5 --no_cpu_baseline --no_e2e --no_parity"
timeout -k 5 200 $CMD > gpurun_out/r2l_bench.json 2> gpurun_out/r2l_bench.err; echo "bench rc=$?"
timeout -k 5 500 ncu --set full --clock-control none --import-source on -k regex:"k_build_lists2|k_pair_forces_tab3" -s 2 -c 4 -f -o gpurun_out/r2l_prof $CMD > gpurun_out/r2l_ncu.log 2>&1; echo "ncu rc=$?"
ncu -i gpurun_out/r2l_prof.ncu-rep --page raw --csv > gpurun_out/r2l_raw.csv 2>/dev/null
ncu -i gpurun_out/r2l_prof.ncu-rep --page source --csv --print-source sass > gpurun_out/r2l_sass.csv 2>/dev/null
ls -la gpurun_out | grep r2l
