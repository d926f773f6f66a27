mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q -x --durations=8 2>&1 | grep -v "Warning\|amp.autocast\|^$" | tail -30 > gpurun_out/r02_test_gpu_all.log; cat gpurun_out/r02_test_gpu_all.log
python tools/tf32_error_budget.py --perturb 2>&1 | tail -9
python tools/tf32_error_budget.py 2>&1 | tail -9
timeout 600 python bench.py --one-tier --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('tf32 ms', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'])"
