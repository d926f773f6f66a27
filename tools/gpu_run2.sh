mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_distributed.py -x -q 2>&1 | tail -5
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 tools/bench_sharded_sht.py 2>gpurun_out/sharded_n2.err | tail -1 > gpurun_out/r02_sharded_sht_n2_peer.json; cat gpurun_out/r02_sharded_sht_n2_peer.json; tail -3 gpurun_out/sharded_n2.err
