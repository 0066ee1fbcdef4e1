#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r2f_gpu.txt
timeout 600 python -m pytest tests/test_gpu_multi.py -x -q > gpurun_out/r2f_pytest_multi.log 2>&1; echo "pytest multi rc=$?"; tail -n 5 gpurun_out/r2f_pytest_multi.log
for peer in 1 0; do
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29671 bench.py --gpus 2 --steps 400 --warmup 100 --no_cpu_baseline --option comm_peer=$peer > gpurun_out/r2f_bench_n2_peer$peer.json 2> gpurun_out/r2f_bench_n2_peer$peer.err; echo "bench n2 peer=$peer rc=$?"
done
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29672 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2f_bench_n2_default.json 2> gpurun_out/r2f_bench_n2_default.err; echo "bench n2 default rc=$?"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29673 bench.py --impl reference --gpus 2 --steps 3 --warmup 1 > gpurun_out/r2f_bench_n2_ref.json 2> gpurun_out/r2f_bench_n2_ref.err; echo "ref n2 rc=$?"
