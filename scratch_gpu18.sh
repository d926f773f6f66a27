#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/profile_kernels.py tf32 > gpurun_out/plain2.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"fft2d" -s 2 -c 2 -o gpurun_out/prof_fft2d python tools/profile_kernels.py tf32 > gpurun_out/ncu2.log 2>&1
echo "ncu full exit $?" >> gpurun_out/summary.txt
tail -2 gpurun_out/ncu2.log
