#!/bin/bash
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -q -m gpu --tb=short > gpurun_out/test_all_gpu.log 2>&1; echo "all gpu tests exit $?" >> gpurun_out/summary.txt
tail -25 gpurun_out/test_all_gpu.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_tf32.json 2> gpurun_out/bench_tf32.err; echo "bench exit $?" >> gpurun_out/summary.txt
python -c "
import json; d=json.load(open('gpurun_out/bench_tf32.json')); print('tf32', d['ms_per_step'], d['value'], d['e2e']['value'])"
cat gpurun_out/summary.txt
