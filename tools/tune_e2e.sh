#!/bin/bash
for sl in 32 64 128; do for th in 8 12 16; do
RSK_E2E_SLAB=$sl RSK_E2E_THREADS=$th python bench.py --steps 5 --warmup 3 --no-cpu-baseline --e2e-steps 5 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('slab=$sl threads=$th e2e=%.3e ms/step=%.1f' % (d['e2e']['value'], d['e2e']['ms_per_step']))"
done; done
