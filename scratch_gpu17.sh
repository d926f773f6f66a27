#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_sht.py -q -m gpu --tb=short > gpurun_out/test_gpu_sht.log 2>&1; echo "sht exit $?" >> gpurun_out/summary.txt
tail -5 gpurun_out/test_gpu_sht.log
timeout 300 python tools/time_stages.py --precision tf32 > gpurun_out/stages_tf32.json 2> gpurun_out/stages_tf32.err
grep "sht_fwd\|specattn_mlp\"" gpurun_out/stages_tf32.json | grep -v GBps
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_tf32.json 2> gpurun_out/bench_tf32.err; echo "bench exit $?" >> gpurun_out/summary.txt
python -c "
import json; d=json.load(open('gpurun_out/bench_tf32.json')); print('tf32', d['ms_per_step'], d['value'], d['e2e']['value'])"
cat gpurun_out/summary.txt
