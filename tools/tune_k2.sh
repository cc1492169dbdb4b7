#!/bin/bash
# Sweep the K2 launch-shape overrides with the real bench (tuning aid).
for tf in 1 0; do for fc in 16 32 64 128; do for sm in 0 56000 72000 112000; do
  RSK_K2_TILEFAST=$tf RSK_K2_FCHUNK=$fc RSK_K2_SMEM=$sm python bench.py --steps 10 --warmup 3 --no-cpu-baseline --e2e-steps 2 --workload ${1:-c2} 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('tf=$tf fc=$fc smem=$sm  value=%.3e  ms/step=%.3f  k2_ms=%.3f  frac=%.3f  sync_value=%.3e' % (d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['frac'], d['sync_api']['value']))"
done; done; done
