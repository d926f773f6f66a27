#!/bin/bash
mkdir -p gpurun_out
for f in test_gpu_sht test_gpu_spectral; do
timeout 600 python -m pytest tests/$f.py -q -m gpu --tb=short > gpurun_out/$f.log 2>&1; echo "$f exit $?" >> gpurun_out/summary.txt
tail -12 gpurun_out/$f.log
done
timeout 300 python tools/time_stages.py --precision tf32 > gpurun_out/stages_tf32.json 2> gpurun_out/stages_tf32.err; echo "stages exit $?" >> gpurun_out/summary.txt
cat gpurun_out/stages_tf32.json; tail -3 gpurun_out/stages_tf32.err
timeout 600 python bench.py --steps 10 --warmup 3 --precision tf32 --no-cpu-baseline > gpurun_out/bench_tf32.json 2> gpurun_out/bench_tf32.err; echo "bench exit $?" >> gpurun_out/summary.txt
cat gpurun_out/bench_tf32.json; tail -3 gpurun_out/bench_tf32.err
cat gpurun_out/summary.txt
