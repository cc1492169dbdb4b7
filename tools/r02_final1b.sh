#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke.txt 2>&1; tail -1 gpurun_out/r02_smoke.txt
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r02_pytest_final.txt 2>&1; tail -1 gpurun_out/r02_pytest_final.txt
timeout 900 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/r02_bench_ref_n1m.json 2> gpurun_out/r02_bench_ref_n1m.err
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r02_bench_n1m.json 2> gpurun_out/r02_bench_n1m.err; echo "bench rc $?"
