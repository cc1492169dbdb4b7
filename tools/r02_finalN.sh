#!/bin/bash
# usage: r02_finalN.sh N   -- bench.py at N GPUs as the driver launches it
N=$1
mkdir -p gpurun_out
timeout 870 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2952$N bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r02_bench_${N}gpu.json 2> gpurun_out/r02_bench_${N}gpu.err
echo "bench N=$N rc $?"
