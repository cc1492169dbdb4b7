#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 tools/gather_sweep.py > gpurun_out/r02_gather_sweep_8gpu.txt 2> gpurun_out/r02_gather_sweep_8gpu.err; echo "rc $?"
cat gpurun_out/r02_gather_sweep_8gpu.txt; tail -3 gpurun_out/r02_gather_sweep_8gpu.err
