#!/bin/bash
mkdir -p gpurun_out
NSIDE=50 timeout -k 5 200 python scratch/sweep_target.py c2 > gpurun_out/r2y_target_n50.log 2>&1; echo "rc=$?"
NSIDE=60 timeout -k 5 200 python scratch/sweep_target.py c2 > gpurun_out/r2y_target_n60.log 2>&1; echo "rc=$?"
