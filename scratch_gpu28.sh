#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/bench_train_step.py --batch 8 --film-layers 1 --steps 3 > gpurun_out/train_fl1.json 2> gpurun_out/train_fl1.err; echo "train fl1 exit $?" >> gpurun_out/summary.txt; cat gpurun_out/train_fl1.json; tail -4 gpurun_out/train_fl1.err
timeout 900 python tools/bench_train_step.py --batch 2 --film-layers 12 --steps 3 > gpurun_out/train_fl12.json 2> gpurun_out/train_fl12.err; echo "train fl12 exit $?" >> gpurun_out/summary.txt; cat gpurun_out/train_fl12.json; tail -4 gpurun_out/train_fl12.err
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_tf32.json 2> gpurun_out/bench_tf32.err; echo "bench exit $?" >> gpurun_out/summary.txt; tail -2 gpurun_out/bench_tf32.err
python -c "
import json; d=json.load(open('gpurun_out/bench_tf32.json')); print('tf32', d['ms_per_step'], d['value'], d['e2e']['value']); print(json.dumps(d['roofline']['stages']))"
cat gpurun_out/summary.txt
