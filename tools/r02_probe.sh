#!/bin/bash
# round-2 first GPU call: host facts, fp64 tensor-path probe, GPU test suite, host ceilings, stored-Stokes path
mkdir -p gpurun_out
{
  echo "== nproc $(nproc)"; lscpu | head -30; free -g; grep -i allowed /proc/self/status
  (numactl -H 2>/dev/null || ls /sys/devices/system/node/)
  cat /sys/kernel/mm/transparent_hugepage/enabled
  nvidia-smi topo -m; nvidia-smi --query-gpu=name,pcie.link.gen.current,pcie.link.width.current,pcie.link.gen.max --format=csv
  grep -c avx512f /proc/cpuinfo; ulimit -l
} > gpurun_out/r02_host_facts.txt 2>&1
./tools/dmma_probe > gpurun_out/r02_dmma_probe.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest1.txt 2>&1; echo "pytest rc $?" >> gpurun_out/r02_pytest1.txt
timeout 300 python tools/d2h_probe.py > gpurun_out/r02_d2h_probe.json 2> gpurun_out/r02_d2h_probe.err
timeout 300 python tools/bench_full_path.py > gpurun_out/r02_full_path.json 2> gpurun_out/r02_full_path.err
tail -5 gpurun_out/r02_pytest1.txt
