#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_fullsize.py -q -m gpu --tb=short > gpurun_out/test_fullsize.log 2>&1; echo "fullsize exit $?" >> gpurun_out/summary.txt
tail -30 gpurun_out/test_fullsize.log
timeout 600 python -m pytest tests/test_gpu_tc.py -q -m gpu --tb=line -k "conv1x1 or small_net" > gpurun_out/test_gpu_tc.log 2>&1; echo "tc exit $?" >> gpurun_out/summary.txt
tail -3 gpurun_out/test_gpu_tc.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_tf32.json 2> gpurun_out/bench_tf32.err; echo "bench exit $?" >> gpurun_out/summary.txt
python -c "
import json; d=json.load(open('gpurun_out/bench_tf32.json')); print('tf32', d['ms_per_step'], d['value'], d['e2e']['value'])"
cat gpurun_out/summary.txt
