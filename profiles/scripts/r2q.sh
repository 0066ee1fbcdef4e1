#!/bin/bash
mkdir -p gpurun_out
N=8
run() { timeout -k 5 $1 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $2 bench.py --gpus $N "${@:4}" > gpurun_out/$3.json 2> gpurun_out/$3.err; echo "$3 rc=$?"; }
run 200 29701 r2q_n8_peer1 --steps 1000 --warmup 200 --no_cpu_baseline --no_e2e --no_parity --option comm_peer=1
run 200 29702 r2q_n8_peer0 --steps 1000 --warmup 200 --no_cpu_baseline --no_e2e --no_parity --option comm_peer=0
run 300 29703 r2q_n8_default --steps 20 --warmup 5
N=4
run 200 29704 r2q_n4_peer1 --steps 1000 --warmup 200 --no_cpu_baseline --no_e2e --no_parity --option comm_peer=1
run 200 29705 r2q_n4_peer0 --steps 1000 --warmup 200 --no_cpu_baseline --no_e2e --no_parity --option comm_peer=0
