#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_tc.py -q -m gpu --tb=short > gpurun_out/test_gpu_tc.log 2>&1; echo "tc exit $?" >> gpurun_out/summary.txt
tail -25 gpurun_out/test_gpu_tc.log
timeout 300 python tools/time_stages.py --precision tf32 > gpurun_out/stages_tf32.json 2> gpurun_out/stages_tf32.err; echo "stages exit $?" >> gpurun_out/summary.txt
cat gpurun_out/stages_tf32.json; tail -3 gpurun_out/stages_tf32.err
timeout 600 python bench.py --steps 10 --warmup 3 --precision tf32 --no-cpu-baseline > gpurun_out/bench_tf32.json 2> gpurun_out/bench_tf32.err; echo "bench exit $?" >> gpurun_out/summary.txt
python -c "
import json; d=json.load(open('gpurun_out/bench_tf32.json')); print(d['ms_per_step'], d['value'], d['e2e'])"
cat gpurun_out/summary.txt
