#!/bin/bash
mkdir -p gpurun_out
for t in 6 10 15; do ./tools/ddio_probe $t 0.5 >> gpurun_out/r02_ddio_probe.txt 2>&1; done
cat gpurun_out/r02_ddio_probe.txt
timeout 600 python bench.py --steps 10 --warmup 3 --configs c3,full --e2e-steps 0 --quick > gpurun_out/r02_bench_c.json 2> gpurun_out/r02_bench_c.err; echo "bench rc $?" >> gpurun_out/r02_bench_c.err
tail -3 gpurun_out/r02_bench_c.err
