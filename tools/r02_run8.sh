#!/bin/bash
mkdir -p gpurun_out
( time timeout 900 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/r02_bench_ref_n1m.json 2> gpurun_out/r02_bench_ref_n1m.err ) 2>&1 | grep real
( time timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r02_bench_n1m.json 2> gpurun_out/r02_bench_n1m.err ) 2>&1 | grep real
tail -3 gpurun_out/r02_bench_n1m.err
