#!/bin/bash
# LST-sweep kernel: register cap (library variants built with -DRSK_SWEEP_MINB=n) x components per thread
cd "$(dirname "$0")/.."
cp pyradiosky_b200/csrc/librsk.so /tmp/librsk_keep.so
mkdir -p pyradiosky_b200/csrc/variants
for mb in 1 3 4; do
  [ -f pyradiosky_b200/csrc/variants/librsk_mb$mb.so ] || (cd pyradiosky_b200/csrc && nvcc -gencode arch=compute_100a,code=sm_100a \
    -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -Xcompiler -pthread -Xcompiler -Wno-subobject-linkage -shared -DRSK_SWEEP_MINB=$mb \
    -o variants/librsk_mb$mb.so rsk_kernels.cu rsk_host.cpp rsk_h5.cpp rsk_h5_write.cpp -lz 2>/dev/null)
done
for mb in 1 3 4; do
  cp pyradiosky_b200/csrc/variants/librsk_mb$mb.so pyradiosky_b200/csrc/librsk.so
  touch pyradiosky_b200/csrc/librsk.so
  for cpt in 1 2 4; do
    echo -n "minb=$mb cpt=$cpt "
    RSK_SWEEP_CPT=$cpt python tools/bench_sweep.py --reps 3 --cpu-sample 100 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('kernel_ms', round(d['kernel_ms'],3), 'evals/s %.3e' % d['kernel_value'], 'err_uas', d['max_alt_err_uas_vs_oracle'])"
  done
done
cp /tmp/librsk_keep.so pyradiosky_b200/csrc/librsk.so
