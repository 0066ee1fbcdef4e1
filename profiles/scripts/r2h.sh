#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_edge.py tests/test_gpu_reactions.py -x -q > gpurun_out/r2h_pytest.log 2>&1; echo "pytest rc=$?"; tail -n 4 gpurun_out/r2h_pytest.log
CLB_TRACE=1 timeout 300 python bench.py --steps 20 --warmup 5 --no_cpu_baseline --no_e2e > gpurun_out/r2h_bench_s20.json 2> gpurun_out/r2h_bench_s20.err; echo "bench rc=$?"
for peer in 0 1; do
CLB_TRACE=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2967$peer bench.py --gpus 2 --steps 60 --warmup 10 --no_cpu_baseline --no_e2e --no_parity --option comm_peer=$peer > gpurun_out/r2h_n2_peer$peer.json 2> gpurun_out/r2h_n2_peer$peer.err; echo "bench n2 peer=$peer rc=$?"
done
MGPU_NSIDE=48 CLB_TRACE=1 timeout 300 python -u -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29655 tests/mgpu_worker.py > gpurun_out/r2h_mgpu48.log 2>&1; echo "mgpu48 rc=$?"
