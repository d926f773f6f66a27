#!/bin/bash
# Round-2 evidence on ONE B200 (run through gpurun): launch lists of both tiers, ncu --set full of the 3xTF32 GEMM, of the
# 2880-point FFT kernels and of the fp32 / tf32 sharded-transform stages.  Every profiled command first runs without ncu.
mkdir -p gpurun_out
for t in fp32 tf32; do
  python bench.py --one-tier --precision $t --no-cpu-baseline --no-multi-gpu-extras --no-graph --steps 2 --warmup 3 > /dev/null 2>&1 || exit 1
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1400 --csv --log-file gpurun_out/r02_launches_bench_${t}_v2.csv \
    python bench.py --one-tier --precision $t --no-cpu-baseline --no-multi-gpu-extras --no-graph --steps 2 --warmup 3 > gpurun_out/ncu_$t.log 2>&1
  python tools/launch_shares.py gpurun_out/r02_launches_bench_${t}_v2.csv > gpurun_out/r02_launch_shares_${t}_v2.txt 2>&1
  head -12 gpurun_out/r02_launch_shares_${t}_v2.txt
done
python tools/x3_one.py 7440 1024 1024 1 1 3 || exit 1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tc3_kernel -c 2 -o gpurun_out/r02_ncu_x3_gemm python tools/x3_one.py 7440 1024 1024 1 1 3 > gpurun_out/ncu_x3.log 2>&1
ncu -i gpurun_out/r02_ncu_x3_gemm.ncu-rep --page raw --csv > gpurun_out/r02_ncu_x3_gemm_raw.csv 2>/dev/null
python tools/time_sharded_stages.py > gpurun_out/r02_sharded_stages_n1_final.json 2>/dev/null || exit 1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"rfft2d_kernel|irfft2d_kernel|gemm_tc3_kernel" -c 8 -o gpurun_out/r02_ncu_sharded_stages python tools/time_sharded_stages.py > gpurun_out/ncu_sharded.log 2>&1
ncu -i gpurun_out/r02_ncu_sharded_stages.ncu-rep --page raw --csv > gpurun_out/r02_ncu_sharded_stages_raw.csv 2>/dev/null
ls -la gpurun_out/*.csv | tail -6
