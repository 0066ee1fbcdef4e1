#!/bin/bash
mkdir -p gpurun_out
timeout -k 5 400 python bench.py --steps 400 --warmup 100 --no_cpu_baseline --no_e2e > gpurun_out/r2p_n1_s400.json 2> gpurun_out/r2p_n1_s400.err; echo "bench n1 s400 rc=$?"
timeout -k 5 400 python bench.py --steps 20 --warmup 5 > gpurun_out/r2p_n1_default.json 2> gpurun_out/r2p_n1_default.err; echo "bench n1 default rc=$?"
timeout -k 5 300 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r2p_ref.json 2> gpurun_out/r2p_ref.err; echo "ref rc=$?"
