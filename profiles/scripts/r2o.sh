#!/bin/bash
mkdir -p gpurun_out
timeout -k 5 300 python -m pytest tests/test_gpu_multi.py -x -q > gpurun_out/r2o_pytest_multi.log 2>&1; echo "pytest multi rc=$?"; tail -n 3 gpurun_out/r2o_pytest_multi.log
for peer in 1 0; do
timeout -k 5 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2967$peer bench.py --gpus 2 --steps 400 --warmup 100 --no_cpu_baseline --option comm_peer=$peer > gpurun_out/r2o_n2_peer$peer.json 2> gpurun_out/r2o_n2_peer$peer.err; echo "bench n2 peer=$peer rc=$?"
done
timeout -k 5 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29672 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2o_n2_default.json 2> gpurun_out/r2o_n2_default.err; echo "bench n2 default rc=$?"
timeout -k 5 400 python bench.py --steps 20 --warmup 5 > gpurun_out/r2o_n1_default.json 2> gpurun_out/r2o_n1_default.err; echo "bench n1 default rc=$?"
