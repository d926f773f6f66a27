mkdir -p gpurun_out
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_n8_final.json 2> gpurun_out/r02_bench_n8_final.err
tail -3 gpurun_out/r02_bench_n8_final.err | cut -c1-300
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_bench_n8_final.json').read().strip().splitlines()[-1])
print('value', d['value'], 'ms', d['ms_per_step'], 'fp32', d['tiers']['fp32']['ms_per_step'])
print('e2e', d['e2e'])
m=d.get('multi_gpu')
print({k:(v.get('ms_per_roundtrip') or v.get('ms_per_step') or v.get('value')) for k,v in m.items()})
print(m['sharded_sht'].get('exchange'), m['sharded_sht'].get('fused_into_fft_kernels'), m['sharded_sht'].get('ms_per_roundtrip_nccl_exchange'))
print(m['ddp_train'])
PY
