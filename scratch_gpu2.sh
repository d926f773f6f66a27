#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_net.py -q -m gpu --tb=short > gpurun_out/test_gpu_net.log 2>&1; echo "net exit $?" >> gpurun_out/summary.txt
tail -15 gpurun_out/test_gpu_net.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_nonlinear_fp32.json 2> gpurun_out/bench_nonlinear_fp32.err; echo "bench exit $?" >> gpurun_out/summary.txt
cat gpurun_out/bench_nonlinear_fp32.json; tail -5 gpurun_out/bench_nonlinear_fp32.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; echo "ref exit $?" >> gpurun_out/summary.txt
cat gpurun_out/bench_reference.json
nproc >> gpurun_out/summary.txt
timeout 900 python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/plain.log 2>&1 &&
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches_r1.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu.log 2>&1
echo "ncu exit $?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt
