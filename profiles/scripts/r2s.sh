#!/bin/bash
mkdir -p gpurun_out
for w in c3 c4; do
timeout -k 5 400 python bench.py --workload $w --steps 200 --warmup 50 --no_parity --no_cpu_baseline > gpurun_out/r2s_${w}_full.json 2> gpurun_out/r2s_${w}_full.err; echo "$w full rc=$?"
timeout -k 5 400 python bench.py --workload $w --scale 6 --steps 200 --warmup 50 --cpu_seconds 10 > gpurun_out/r2s_${w}_s6.json 2> gpurun_out/r2s_${w}_s6.err; echo "$w scale6 rc=$?"
done
timeout -k 5 300 python bench.py --workload c5 --scale 10 --steps 200 --warmup 50 --cpu_seconds 10 > gpurun_out/r2s_c5_s10.json 2> gpurun_out/r2s_c5_s10.err; echo "c5 scale10 rc=$?"
CMD="python bench.py --workload c5 --scale 10 --steps 20 --warmup 5 --no_cpu_baseline --no_e2e --no_parity"
timeout -k 5 400 ncu --set full --clock-control none --import-source on -k regex:"k_pair_forces_tab3" -s 3 -c 1 -f -o gpurun_out/r2s_c5_pair $CMD > gpurun_out/r2s_ncu.log 2>&1; echo "ncu rc=$?"
ncu -i gpurun_out/r2s_c5_pair.ncu-rep --page raw --csv > gpurun_out/r2s_c5_pair_raw.csv 2>/dev/null
