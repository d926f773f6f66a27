#!/bin/bash
mkdir -p gpurun_out
timeout 600 python bench.py --steps 10 --warmup 3 --precision tf32 --no-cpu-baseline > gpurun_out/bench_graph.json 2> gpurun_out/bench_graph.err; echo "bench graph exit $?" >> gpurun_out/summary.txt
tail -5 gpurun_out/bench_graph.err
python -c "
import json; d=json.load(open('gpurun_out/bench_graph.json')); print('graph', d['ms_per_step'], d['value'], d['e2e'], d['gpu_launches'], d['output_finite'])"
timeout 600 python bench.py --steps 10 --warmup 3 --precision tf32 --no-cpu-baseline --no-graph > gpurun_out/bench_eager.json 2> gpurun_out/bench_eager.err; echo "bench eager exit $?" >> gpurun_out/summary.txt
python -c "
import json; d=json.load(open('gpurun_out/bench_eager.json')); print('eager', d['ms_per_step'], d['value'], d['e2e'], d['gpu_launches'])"
cat gpurun_out/summary.txt
