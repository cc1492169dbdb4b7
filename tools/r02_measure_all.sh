#!/bin/bash
# Everything profiles/r02_* quotes, as run under gpurun in round 2 (each block was its own gpurun call; NGPU = GPUs leased).
# Outputs go to gpurun_out/ and are copied to profiles/ by hand (see profiles/r02_summary.md for what is what).
set -u
mkdir -p gpurun_out
what=${1:-help}
case "$what" in
  probe)      # NGPU=1: host description, fp64 tensor-path probe, host delivery ceilings, stored-cube path by mask density
    { nproc; lscpu | head -30; free -g; grep -i allowed /proc/self/status; cat /sys/kernel/mm/transparent_hugepage/enabled
      nvidia-smi topo -m; nvidia-smi --query-gpu=name,pcie.link.gen.current,pcie.link.width.current --format=csv; } > gpurun_out/r02_host_facts.txt 2>&1
    ./tools/dmma_probe > gpurun_out/r02_dmma_probe.txt 2>&1        # nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/dmma_probe tools/dmma_probe.cu
    python tools/d2h_probe.py > gpurun_out/r02_d2h_probe.json
    python tools/d2h_probe2.py > gpurun_out/r02_d2h_probe2.json
    for t in 6 10 15; do ./tools/ddio_probe $t 0.5; done > gpurun_out/r02_ddio_probe.txt 2>&1   # nvcc -O3 -std=c++17 -Xcompiler -march=sapphirerapids,-pthread
    python tools/bench_full_path.py > gpurun_out/r02_full_path.json
    for fc in 32 128 256; do RSK_K2_FCHUNK=$fc python tools/bench_full_path.py | head -2 | sed "s/^/fchunk=$fc /"; done > gpurun_out/r02_full_fchunk.txt
    python tools/fchunk_small.py > gpurun_out/r02_fchunk_small.txt ;;
  tests)      # NGPU=1: what the driver runs at round end
    python -c "import __graft_entry__ as g; g.smoke()" && python -m pytest tests -m gpu -q ;;
  bench1)     # NGPU=1: both bench arms as the driver launches them
    python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/r02_bench_ref_n1m.json
    python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r02_bench_n1m.json ;;
  benchN)     # NGPU=N in {2,4,8}: r02_measure_all.sh benchN N
    N=$2
    python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2952$N \
      bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r02_bench_${N}gpu.json ;;
  ncu)        # NGPU=1: launch list of one bench step, full captures of the headline and the stored-cube kernels
    CMD="python bench.py --steps 2 --warmup 3 --configs none --e2e-steps 0 --quick --no-cpu-baseline"
    $CMD > gpurun_out/r02_ncu_plain.log 2>&1 &&
    ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_bench_n1m.csv $CMD
    ncu --set full --clock-control none --import-source on -k regex:k_eval -s 3 -c 2 -f -o gpurun_out/r02_prof_k_eval_n1m $CMD
    CMD2="python bench.py --workload full --steps 2 --warmup 3 --configs none --e2e-steps 0 --quick --no-cpu-baseline"
    $CMD2 > gpurun_out/r02_ncu_plain2.log 2>&1 &&
    ncu --set full --clock-control none --import-source on -k regex:k_eval -s 4 -c 2 -f -o gpurun_out/r02_prof_k_eval_full $CMD2 ;;
  gather)     # NGPU=2 and 8: bit-equality of both assembly routes, route x streams x slabs sweep
    N=$2
    GC_N=400000 GC_NF=256 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 tools/peer_gather_check.py > gpurun_out/r02_peer_gather_${N}gpu.json
    python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 tools/gather_sweep.py > gpurun_out/r02_gather_sweep_${N}gpu.txt ;;
  *) echo "usage: $0 probe|tests|bench1|benchN N|ncu|gather N" ;;
esac
