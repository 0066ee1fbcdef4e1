#!/bin/bash
mkdir -p gpurun_out
for S in 13 14; do
timeout -k 5 200 python bench.py --workload c5 --scale $S --steps 100 --warmup 20 --no_cpu_baseline --no_e2e --no_parity > gpurun_out/r3e_c5_n1_s$S.json 2> gpurun_out/r3e_c5_n1_s$S.err; echo "c5 n1 s$S rc=$?"
done
