#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/profile_conv.py > gpurun_out/plain.log 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"conv_tc_kernel" -s 1 -c 1 -o gpurun_out/prof_conv2 python tools/profile_conv.py > gpurun_out/ncu.log 2>&1
echo "ncu exit $?" >> gpurun_out/summary.txt; tail -2 gpurun_out/ncu.log
