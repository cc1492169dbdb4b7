#!/bin/bash
# Jones-array apparent-coherency kernel (one output plane per blockIdx.y): register cap sweep
cd "$(dirname "$0")/.."
cp pyradiosky_b200/csrc/librsk.so /tmp/librsk_keep.so
for mb in 2 3 4; do
  (cd pyradiosky_b200/csrc && nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -Xcompiler -pthread \
     -Xcompiler -Wno-subobject-linkage -shared -DRSK_MINB_PLANE=$mb -o librsk.so rsk_kernels.cu rsk_host.cpp rsk_h5.cpp rsk_h5_write.cpp -lz 2>/dev/null)
  echo "RSK_MINB_PLANE=$mb"
  python tools/prof_apparent.py --time 2>/dev/null | grep antenna
done
cp /tmp/librsk_keep.so pyradiosky_b200/csrc/librsk.so
