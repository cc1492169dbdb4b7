#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r02_bench_2gpu.json 2> gpurun_out/r02_bench_2gpu.err; echo "bench rc $?" >> gpurun_out/r02_bench_2gpu.err
tail -8 gpurun_out/r02_bench_2gpu.err
nproc; free -g | head -2; nvidia-smi topo -m | head -8
