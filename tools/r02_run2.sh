#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r02_pytest2.txt 2>&1; echo "pytest rc $?" >> gpurun_out/r02_pytest2.txt
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r02_bench_a.json 2> gpurun_out/r02_bench_a.err; echo "bench rc $?" >> gpurun_out/r02_bench_a.err
timeout 300 python tools/bench_full_path.py > gpurun_out/r02_full_path2.json 2> gpurun_out/r02_full_path2.err
tail -15 gpurun_out/r02_pytest2.txt; tail -5 gpurun_out/r02_bench_a.err
