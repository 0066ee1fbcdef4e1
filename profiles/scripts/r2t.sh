#!/bin/bash
mkdir -p gpurun_out
timeout -k 5 300 python scratch/sweep_c5.py c5 10 > gpurun_out/r2t_sweep_c5.log 2>&1; echo "c5 rc=$?"
timeout -k 5 300 python scratch/sweep_c5.py c3 6 > gpurun_out/r2t_sweep_c3.log 2>&1; echo "c3 rc=$?"
