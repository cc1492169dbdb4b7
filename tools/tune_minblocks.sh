#!/bin/bash
# Rebuild librsk.so on the GPU box with different register caps (min resident blocks) and run the bench.
cd pyradiosky_b200/csrc
for v in "1 1" "4 1" "5 1" "1 3" "1 4"; do set -- $v
  nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -shared -DRSK_MINB_SI=$1 -DRSK_MINB_BANDED=$2 -o librsk.so rsk_kernels.cu rsk_host.cpp || exit 1
  for w in c2 c4; do
  (cd ../..; python bench.py --steps 15 --warmup 3 --no-cpu-baseline --e2e-steps 2 --workload $w 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('minb_si=$1 minb_banded=$2 $w value=%.3e ms/step=%.3f k2_ms=%.3f frac=%.3f' % (d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['frac']))")
  done
done
