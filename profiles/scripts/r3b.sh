#!/bin/bash
mkdir -p gpurun_out
N=8
run() { timeout -k 5 $1 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $2 bench.py --gpus $N "${@:4}" > gpurun_out/$3.json 2> gpurun_out/$3.err; rc=$?; echo "$3 rc=$rc"; return $rc; }
run 150 29721 r3b_n8_c5_s8 --workload c5 --scale 8 --steps 100 --warmup 20 --no_cpu_baseline --no_e2e || exit 0
run 420 29722 r3b_n8_c5_full --workload c5 --steps 100 --warmup 20 --no_cpu_baseline --no_e2e --no_parity
