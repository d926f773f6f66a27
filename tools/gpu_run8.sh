mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_x3.py -x -q 2>&1 | tail -3
timeout 300 python tools/x3_bench.py > gpurun_out/r02_x3_bench_v3.json 2> gpurun_out/r02_x3_bench.err
timeout 900 python -m pytest tests/test_gpu_fullsize.py tests/test_gpu_parity_r2.py -x -q 2>&1 | tail -5
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/r02_bench_v3.json 2> gpurun_out/r02_bench_v3.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_bench_v3.json').read())
for t,v in d['tiers'].items(): print(t, v['ms_per_step'], v['e2e_ms_per_step'], v['gpu_launches'])
PY
export MSFNO_B200_LIB=$PWD/build_trace/libmsfno_b200_trace.so
python tools/x3_one.py 7440 1024 1024 1 1 3 > gpurun_out/r02_x3_timeline_7440x1024x1024.txt 2>&1
python tools/x3_one.py 128 128 1024 1 1 3 > gpurun_out/r02_x3_timeline_single_cta.txt 2>&1
