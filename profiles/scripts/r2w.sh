#!/bin/bash
mkdir -p gpurun_out
timeout -k 5 120 python scratch/dbg_crit.py > gpurun_out/r2w_dbg.log 2>&1; echo "dbg rc=$?"; grep "crit" gpurun_out/r2w_dbg.log | awk 'NR%3==1' | head -30
