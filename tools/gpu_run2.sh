mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_x3.py -x -q 2>&1 | tail -5 > gpurun_out/r02_test_x3.log
cat gpurun_out/r02_test_x3.log
timeout 300 python tools/x3_bench.py > gpurun_out/r02_x3_bench.json 2> gpurun_out/r02_x3_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_x3_bench.json'))
for k,v in d.items(): print(k, {e:(r['ms'],r['TFLOPs']) for e,r in v.items()})
PY
