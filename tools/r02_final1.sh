#!/bin/bash
# round-2 final single-GPU measurements: both bench arms as the driver runs them, then the ncu passes of the same command
mkdir -p gpurun_out
timeout 900 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/r02_bench_ref_n1m.json 2> gpurun_out/r02_bench_ref_n1m.err
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r02_bench_n1m.json 2> gpurun_out/r02_bench_n1m.err; echo "bench rc $?"
CMD="python bench.py --steps 2 --warmup 3 --configs none --e2e-steps 0 --quick --no-cpu-baseline"
$CMD > gpurun_out/r02_ncu_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_bench_n1m.csv $CMD > gpurun_out/r02_ncu_a.log 2>&1
echo "ncu a rc $?"
ncu --set full --clock-control none --import-source on -k regex:k_eval -s 3 -c 2 -f -o gpurun_out/r02_prof_k_eval_n1m $CMD > gpurun_out/r02_ncu_b.log 2>&1
echo "ncu b rc $?"
CMD2="python bench.py --workload full --steps 2 --warmup 3 --configs none --e2e-steps 0 --quick --no-cpu-baseline"
$CMD2 > gpurun_out/r02_ncu_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_evalILi3 -s 3 -c 2 -f -o gpurun_out/r02_prof_k_eval_full $CMD2 > gpurun_out/r02_ncu_c.log 2>&1
echo "ncu c rc $?"
ls -la gpurun_out/*.ncu-rep
