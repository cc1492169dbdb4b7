#!/bin/bash
mkdir -p gpurun_out
CMD2="python bench.py --workload full --steps 2 --warmup 3 --configs none --e2e-steps 0 --quick --no-cpu-baseline"
$CMD2 > gpurun_out/r02_ncu_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_eval -s 4 -c 2 -f -o gpurun_out/r02_prof_k_eval_full $CMD2 > gpurun_out/r02_ncu_c.log 2>&1
echo "ncu c rc $?"
ls -la gpurun_out/*.ncu-rep; tail -n 3 gpurun_out/r02_ncu_c.log
