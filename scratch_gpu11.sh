#!/bin/bash
mkdir -p gpurun_out
timeout 600 python bench.py --steps 1 --warmup 3 --precision tf32 --no-cpu-baseline --no-graph > gpurun_out/plain.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/launches_tf32_v2.csv python bench.py --steps 1 --warmup 3 --precision tf32 --no-cpu-baseline --no-graph > gpurun_out/ncu1.log 2>&1
echo "ncu launches exit $?" >> gpurun_out/summary.txt
timeout 300 python tools/profile_kernels.py tf32 > gpurun_out/plain2.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"rfft2d_kernel|gemm_tc_kernel" -s 4 -c 10 -o gpurun_out/prof_r1b python tools/profile_kernels.py tf32 > gpurun_out/ncu2.log 2>&1
echo "ncu full exit $?" >> gpurun_out/summary.txt
tail -3 gpurun_out/ncu2.log
cat gpurun_out/summary.txt
