#!/bin/bash
mkdir -p gpurun_out
GC_N=400000 GC_NF=256 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tools/peer_gather_check.py > gpurun_out/r02_peer_gather_2gpu.json 2> gpurun_out/r02_peer_gather_2gpu.err; echo "rc $?" >> gpurun_out/r02_peer_gather_2gpu.err
tail -3 gpurun_out/r02_peer_gather_2gpu.err; cat gpurun_out/r02_peer_gather_2gpu.json | tail -2
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 --configs none --e2e-steps 0 --quick > gpurun_out/r02_bench_2gpu_c.json 2> gpurun_out/r02_bench_2gpu_c.err; echo "bench rc $?" >> gpurun_out/r02_bench_2gpu_c.err
tail -4 gpurun_out/r02_bench_2gpu_c.err
