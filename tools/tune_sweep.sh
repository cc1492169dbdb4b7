#!/bin/bash
# LST-sweep kernel: register cap (-DRSK_SWEEP_MINB=n) x time steps per block (-DRSK_SWEEP_TIMES=n) x components per thread
cd "$(dirname "$0")/.."
cp pyradiosky_b200/csrc/librsk.so /tmp/librsk_keep.so
mkdir -p pyradiosky_b200/csrc/variants
VARIANTS="${VARIANTS:-4:16 4:32 4:64 3:32 3:64}"
for v in $VARIANTS; do
  mb=${v%%:*}; nt=${v##*:}
  [ -f pyradiosky_b200/csrc/variants/librsk_mb${mb}_t$nt.so ] || (cd pyradiosky_b200/csrc && nvcc -gencode arch=compute_100a,code=sm_100a \
    -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -Xcompiler -pthread -Xcompiler -Wno-subobject-linkage -shared -DRSK_SWEEP_MINB=$mb -DRSK_SWEEP_TIMES=$nt \
    -o variants/librsk_mb${mb}_t$nt.so rsk_kernels.cu rsk_host.cpp rsk_h5.cpp rsk_h5_write.cpp -lz 2>/dev/null)
done
for v in $VARIANTS; do
  mb=${v%%:*}; nt=${v##*:}
  cp pyradiosky_b200/csrc/variants/librsk_mb${mb}_t$nt.so pyradiosky_b200/csrc/librsk.so
  touch pyradiosky_b200/csrc/librsk.so
  for cpt in ${CPTS:-1 2}; do
    echo -n "minb=$mb times=$nt cpt=$cpt "
    RSK_SWEEP_CPT=$cpt python tools/bench_sweep.py --reps 3 --cpu-sample 100 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('kernel_ms', round(d['kernel_ms'],3), 'evals/s %.3e' % d['kernel_value'], 'err_uas', d['max_alt_err_uas_vs_oracle'])"
  done
done
cp /tmp/librsk_keep.so pyradiosky_b200/csrc/librsk.so
