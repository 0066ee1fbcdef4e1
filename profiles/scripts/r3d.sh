#!/bin/bash
mkdir -p gpurun_out
N=$1; S=$2
timeout -k 5 $3 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2973$N bench.py --gpus $N --workload c5 --scale $S --steps 100 --warmup 20 --no_cpu_baseline --no_e2e --no_parity > gpurun_out/r3d_c5_n${N}_s${S}.json 2> gpurun_out/r3d_c5_n${N}_s${S}.err; echo "c5 n$N s$S rc=$?"
