#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/d2h_probe2.py > gpurun_out/r02_d2h_probe2.json 2> gpurun_out/r02_d2h_probe2.err
for fc in 32 128 256; do RSK_K2_FCHUNK=$fc timeout 120 python tools/bench_full_path.py 2>/dev/null | head -2 | sed "s/^/fchunk=$fc /" >> gpurun_out/r02_full_fchunk.txt; done
timeout 600 python bench.py --steps 10 --warmup 3 --configs c3 --e2e-steps 0 > gpurun_out/r02_bench_b.json 2> gpurun_out/r02_bench_b.err; echo "bench rc $?" >> gpurun_out/r02_bench_b.err
cat gpurun_out/r02_d2h_probe2.json; tail -3 gpurun_out/r02_d2h_probe2.err; cat gpurun_out/r02_full_fchunk.txt; tail -3 gpurun_out/r02_bench_b.err
