#!/bin/bash
mkdir -p gpurun_out
for peer in 1 0; do
CLB_TRACE=1 timeout -k 5 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2967$peer bench.py --gpus 2 --steps 60 --warmup 10 --no_cpu_baseline --no_e2e --no_parity --option comm_peer=$peer > gpurun_out/r2n_n2_peer$peer.json 2> gpurun_out/r2n_n2_peer$peer.err; echo "bench n2 peer=$peer rc=$?"
done
MGPU_NSIDE=48 CLB_TRACE=1 timeout -k 5 200 python -u -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29655 tests/mgpu_worker.py > gpurun_out/r2n_mgpu48.log 2>&1; echo "mgpu48 rc=$?"
grep -c "clb rebuild" gpurun_out/r2n_n2_peer0.err gpurun_out/r2n_n2_peer1.err
