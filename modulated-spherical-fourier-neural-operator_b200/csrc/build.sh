#!/bin/bash
# Build libmsfno_b200.so in-tree for sm_100a (cross-compiles without a GPU).
set -e
cd "$(dirname "$0")"
OUT=../libmsfno_b200.so
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -Xptxas -v"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
mkdir -p build
pids=()
for f in fft fft2d plan gemm_ffma specconv specattn sht elementwise gemm_tc conv_tc conv1x1; do
  [ -f $f.cu ] || continue
  ( $NVCC $FLAGS -c $f.cu -o build/$f.o > build/$f.log 2>&1 || { cat build/$f.log; exit 1; } ) &
  pids+=($!)
done
for p in "${pids[@]}"; do wait $p; done
$NVCC -shared -o $OUT build/*.o -lcudart
echo "built $OUT"
