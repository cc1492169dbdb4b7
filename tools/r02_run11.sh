#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_round2_gpu.py -m gpu -q -x -k "overlap" 2>&1 | tail -5
for ov in 0 1; do
  if [ $ov = 0 ]; then export RSK_BENCH_NO_OVERLAP=1; else unset RSK_BENCH_NO_OVERLAP; fi
  for wl in c2 n1m; do
    timeout 300 python bench.py --workload $wl --steps 20 --warmup 5 --configs none --e2e-steps 0 --quick --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('overlap=$ov', '$wl', 'value %.4g ms %.4f kernel %.4f'%(d['value'],d['ms_per_step'],d['roofline']['kernel_ms']))"
  done
done
