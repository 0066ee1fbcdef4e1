mkdir -p gpurun_out
timeout 40 python scratch/r4a.py > gpurun_out/r4a_stdout.log 2>&1
echo "rc=$?" >> gpurun_out/r4a_stdout.log
tail -5 gpurun_out/r4a_restrict.log
