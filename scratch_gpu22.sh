#!/bin/bash
mkdir -p gpurun_out
timeout 900 python tools/parity_report.py > gpurun_out/parity.json 2> gpurun_out/parity.err; cat gpurun_out/parity.json; tail -3 gpurun_out/parity.err
timeout 600 python -m pytest tests/test_gpu_tc.py tests/test_gpu_serving.py tests/test_gpu_spectral.py -q -m gpu --tb=line > gpurun_out/t.log 2>&1; echo "tests exit $?" >> gpurun_out/summary.txt; tail -3 gpurun_out/t.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_tf32.json 2> gpurun_out/bench_tf32.err; echo "bench exit $?" >> gpurun_out/summary.txt
python -c "
import json; d=json.load(open('gpurun_out/bench_tf32.json')); print('tf32', d['ms_per_step'], d['value'], d['e2e']['value'], d['gpu_launches'])"
cat gpurun_out/summary.txt
