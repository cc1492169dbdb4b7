#!/bin/bash
# launch list + full capture of the headline kernel (n1m) -- one ncu call
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --configs none --e2e-steps 0 --quick --no-cpu-baseline"
$CMD > gpurun_out/r02_ncu_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_bench_n1m.csv $CMD > gpurun_out/r02_ncu_a.log 2>&1
echo "ncu a rc $?"
ncu --set full --clock-control none --import-source on -k regex:k_eval -s 3 -c 2 -f -o gpurun_out/r02_prof_k_eval_n1m $CMD > gpurun_out/r02_ncu_b.log 2>&1
echo "ncu b rc $?"
ls -la gpurun_out/*.ncu-rep; tail -3 gpurun_out/r02_ncu_a.log gpurun_out/r02_ncu_b.log
