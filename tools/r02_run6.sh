#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r02_pytest3.txt 2>&1; echo "pytest rc $?" >> gpurun_out/r02_pytest3.txt
tail -4 gpurun_out/r02_pytest3.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 --configs c3 --e2e-steps 0 --quick > gpurun_out/r02_bench_2gpu_b.json 2> gpurun_out/r02_bench_2gpu_b.err; echo "bench rc $?" >> gpurun_out/r02_bench_2gpu_b.err
tail -4 gpurun_out/r02_bench_2gpu_b.err
