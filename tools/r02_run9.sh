#!/bin/bash
mkdir -p gpurun_out
nproc > gpurun_out/r02_8gpu_host.txt; free -g >> gpurun_out/r02_8gpu_host.txt; nvidia-smi topo -m >> gpurun_out/r02_8gpu_host.txt 2>&1; lscpu | grep -E "NUMA|Socket|Model name" >> gpurun_out/r02_8gpu_host.txt
for n in 8 4; do
( time timeout 870 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 20 --warmup 5 > gpurun_out/r02_bench_${n}gpu.json 2> gpurun_out/r02_bench_${n}gpu.err ) 2>&1 | grep real
echo "rc $?" >> gpurun_out/r02_bench_${n}gpu.err
done
tail -5 gpurun_out/r02_bench_8gpu.err
