#!/bin/bash
# How much does the per-block set-up cost? Same sweep as tune_k2.sh on an unpolarised catalogue (no trig in K2).
for tf in 1 0; do for fc in 8 16 32 64; do for sm in 0 112000; do
  RSK_BENCH_UNPOLARIZED=1 RSK_K2_TILEFAST=$tf RSK_K2_FCHUNK=$fc RSK_K2_SMEM=$sm python bench.py --steps 10 --warmup 3 --no-cpu-baseline --e2e-steps 2 --workload ${1:-c2} 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('unpol tf=$tf fc=$fc smem=$sm  value=%.3e  ms/step=%.3f  k2_ms=%.3f  frac=%.3f' % (d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['frac']))"
done; done; done
