#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_tc.py -q -m gpu --tb=line > gpurun_out/test_gpu_tc.log 2>&1; echo "tc exit $?" >> gpurun_out/summary.txt
tail -5 gpurun_out/test_gpu_tc.log
timeout 600 python tools/time_convs.py > gpurun_out/convs.json 2> gpurun_out/convs.err; echo "convs exit $?" >> gpurun_out/summary.txt
grep -v fp32 gpurun_out/convs.json; tail -3 gpurun_out/convs.err
timeout 300 python tools/time_stages.py --precision tf32 > gpurun_out/stages_tf32.json 2> gpurun_out/stages_tf32.err
grep "specattn\|sht_fwd_full\"" gpurun_out/stages_tf32.json
timeout 600 python bench.py --steps 10 --warmup 3 --precision tf32 --no-cpu-baseline > gpurun_out/bench_tf32.json 2> gpurun_out/bench_tf32.err; echo "bench exit $?" >> gpurun_out/summary.txt
python -c "
import json; d=json.load(open('gpurun_out/bench_tf32.json')); print('tf32', d['ms_per_step'], d['value'], d['e2e'], d['gpu_launches'], d['roofline']['achieved'])"
cat gpurun_out/summary.txt
