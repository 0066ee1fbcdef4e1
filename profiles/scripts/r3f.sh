#!/bin/bash
mkdir -p gpurun_out
N=$1
timeout -k 5 $2 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2974$N bench.py --gpus $N --workload c3 --steps 100 --warmup 20 --no_cpu_baseline --no_e2e --no_parity > gpurun_out/r3f_c3_n${N}.json 2> gpurun_out/r3f_c3_n${N}.err; echo "c3 n$N rc=$?"
