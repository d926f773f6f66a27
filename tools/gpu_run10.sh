mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --durations=5 2>&1 | grep -v "Warning\|amp.autocast\|^$" | tail -14 > gpurun_out/r02_test_gpu_all_v3.log; tail -8 gpurun_out/r02_test_gpu_all_v3.log
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/r02_bench_v4.json 2> gpurun_out/r02_bench_v4.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_bench_v4.json').read())
for t,v in d['tiers'].items(): print(t, v['ms_per_step'], v['e2e_ms_per_step'], v['gpu_launches'])
PY
python __graft_entry__.py 2>&1 | tail -3
