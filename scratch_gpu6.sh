#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpus.txt
timeout 600 python -m pytest tests/test_gpu_distributed.py -q -m gpu --tb=short > gpurun_out/test_gpu_dist.log 2>&1; echo "dist exit $?" >> gpurun_out/summary.txt
tail -15 gpurun_out/test_gpu_dist.log
timeout 600 python bench.py --steps 10 --warmup 3 --precision tf32 --no-cpu-baseline > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench1 exit $?" >> gpurun_out/summary.txt
cat gpurun_out/bench_n1.json; tail -3 gpurun_out/bench_n1.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 --precision tf32 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo "bench2 exit $?" >> gpurun_out/summary.txt
cat gpurun_out/bench_n2.json; tail -3 gpurun_out/bench_n2.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tools/bench_sharded_sht.py > gpurun_out/sharded_n2.json 2> gpurun_out/sharded_n2.err; echo "sharded exit $?" >> gpurun_out/summary.txt
cat gpurun_out/sharded_n2.json; tail -5 gpurun_out/sharded_n2.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 1 --master-addr 127.0.0.1 --master-port 29513 tools/bench_sharded_sht.py > gpurun_out/sharded_n1.json 2> gpurun_out/sharded_n1.err; echo "sharded1 exit $?" >> gpurun_out/summary.txt
cat gpurun_out/sharded_n1.json; tail -5 gpurun_out/sharded_n1.err
cat gpurun_out/summary.txt
