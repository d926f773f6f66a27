#!/bin/bash
# product library + the -DMSFNO_TRACE build beside it (build_trace/, git-ignored, ships with gpurun)
set -e
cd "$(dirname "$0")/.."
mkdir -p build_trace
( MSFNO_OUT=$PWD/build_trace/libmsfno_b200_trace.so MSFNO_OBJ=$PWD/build_trace/obj MSFNO_EXTRA_FLAGS="-DMSFNO_TRACE" bash modulated-spherical-fourier-neural-operator_b200/csrc/build.sh 2>&1 | tail -1 ) &
bash modulated-spherical-fourier-neural-operator_b200/csrc/build.sh 2>&1 | tail -1
wait
