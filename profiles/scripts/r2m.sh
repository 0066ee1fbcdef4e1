#!/bin/bash
mkdir -p gpurun_out
CMD="python bench.py --steps 20 --warmup 5 --no_cpu_baseline --no_e2e --no_parity"
timeout -k 5 500 ncu --set full --clock-control none --import-source on -k regex:"k_build_lists2" -s 1 -c 1 -f -o gpurun_out/r2m_prof $CMD > gpurun_out/r2m_ncu.log 2>&1; echo "ncu rc=$?"
ncu -i gpurun_out/r2m_prof.ncu-rep --page raw --csv > gpurun_out/r2m_raw.csv 2>/dev/null
ncu -i gpurun_out/r2m_prof.ncu-rep --page source --csv --print-source sass > gpurun_out/r2m_sass.csv 2>/dev/null
timeout -k 5 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r2m_launches.csv $CMD > gpurun_out/r2m_ncu2.log 2>&1; echo "ncu launches rc=$?"
ls -la gpurun_out | grep r2m
