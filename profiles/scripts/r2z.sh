#!/bin/bash
mkdir -p gpurun_out
timeout -k 5 300 python -m pytest tests/test_gpu_multi.py -x -q > gpurun_out/r2z_pytest_multi.log 2>&1; echo "pytest multi rc=$?"; tail -n 3 gpurun_out/r2z_pytest_multi.log
N=4
run() { timeout -k 5 $1 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $2 bench.py --gpus $N "${@:4}" > gpurun_out/$3.json 2> gpurun_out/$3.err; echo "$3 rc=$?"; }
run 200 29704 r2z_n4_ov1 --steps 1000 --warmup 200 --no_cpu_baseline --no_e2e --no_parity
run 200 29705 r2z_n4_ov0 --steps 1000 --warmup 200 --no_cpu_baseline --no_e2e --no_parity --option overlap_halo=0
run 300 29706 r2z_n4_default --steps 20 --warmup 5
