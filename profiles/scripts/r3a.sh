#!/bin/bash
mkdir -p gpurun_out
N=8
run() { timeout -k 5 $1 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $2 bench.py --gpus $N "${@:4}" > gpurun_out/$3.json 2> gpurun_out/$3.err; echo "$3 rc=$?"; }
run 240 29711 r3a_n8_default --steps 20 --warmup 5
run 150 29712 r3a_n8_s1000 --steps 1000 --warmup 200 --no_cpu_baseline --no_e2e --no_parity
run 400 29713 r3a_n8_c5 --workload c5 --steps 100 --warmup 20 --no_cpu_baseline --no_e2e --no_parity
