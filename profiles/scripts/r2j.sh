#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/r2j_bisect.log; : > $L
for combo in "1 0 0" "2 0 0" "1 1 0" "1 0 1" "2 1 1"; do
  echo "== combo (build perm pipe) $combo" >> $L
  timeout -k 5 90 python scratch/bisect.py $combo 16 >> $L 2>&1; rc=$?
  echo "rc=$rc" >> $L
  if [ $rc -ne 0 ]; then
    echo "== sanitizer for $combo" >> $L
    timeout -k 5 240 compute-sanitizer --tool memcheck --print-limit 8 python scratch/bisect.py $combo 10 2>&1 | grep -v "^\s*$" | head -80 >> $L
    timeout -k 5 240 compute-sanitizer --tool synccheck --print-limit 8 python scratch/bisect.py $combo 10 2>&1 | grep -v "^\s*$" | head -60 >> $L
  fi
done
tail -n 60 $L
