#!/bin/bash
# Build libmsfno_b200.so in-tree for sm_100a (cross-compiles without a GPU).
set -e
cd "$(dirname "$0")"
OUT=${MSFNO_OUT:-../libmsfno_b200.so}     # MSFNO_OUT / MSFNO_OBJ: a second build (e.g. -DMSFNO_TRACE) beside the product library
OBJ=${MSFNO_OBJ:-build}
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -Xptxas -v ${MSFNO_EXTRA_FLAGS:-}"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
mkdir -p $OBJ
SRCS="fft fft2d plan gemm_ffma specconv specattn sht elementwise gemm_tc conv_tc conv1x1 mlp_tc dft_tc losses peer"
pids=()
for f in $SRCS; do
  rm -f $OBJ/$f.o
  ( $NVCC $FLAGS -c $f.cu -o $OBJ/$f.o > $OBJ/$f.log 2>&1 ) &
  pids+=($!)
done
fail=0
for p in "${pids[@]}"; do wait $p || fail=1; done
for f in $SRCS; do
  if [ ! -f $OBJ/$f.o ]; then echo "=== nvcc failed on $f.cu ==="; grep -v "^ptxas info\|Function properties\|bytes stack frame\|Compiling entry\|^$" $OBJ/$f.log | head -30; fail=1; fi
done
[ $fail -eq 0 ] || { echo "BUILD FAILED"; exit 1; }
OBJS=""; for f in $SRCS; do OBJS="$OBJS $OBJ/$f.o"; done
$NVCC -shared -o $OUT $OBJS -lcudart
echo "built $OUT"
