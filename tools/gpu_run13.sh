mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29515 tools/bench_sharded_sht.py 2>/dev/null | tail -1 > gpurun_out/r02_sharded_sht_n8_peer.json; cat gpurun_out/r02_sharded_sht_n8_peer.json
