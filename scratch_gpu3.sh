#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/time_stages.py --precision tf32 > gpurun_out/stages_tf32.json 2> gpurun_out/stages_tf32.err; echo "stages exit $?" >> gpurun_out/summary.txt
cat gpurun_out/stages_tf32.json; tail -3 gpurun_out/stages_tf32.err
timeout 600 python bench.py --steps 5 --warmup 3 --precision tf32 --no-cpu-baseline > gpurun_out/bench_tf32.json 2> gpurun_out/bench_tf32.err; echo "bench exit $?" >> gpurun_out/summary.txt
cat gpurun_out/bench_tf32.json; tail -3 gpurun_out/bench_tf32.err
cat gpurun_out/summary.txt
