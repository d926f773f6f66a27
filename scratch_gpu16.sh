#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu --tb=short -x > gpurun_out/test_all_gpu.log 2>&1; echo "all gpu tests exit $?" >> gpurun_out/summary.txt
tail -8 gpurun_out/test_all_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/summary.txt; tail -2 gpurun_out/smoke.log
timeout 900 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench default exit $?" >> gpurun_out/summary.txt
cat gpurun_out/bench_default.json; tail -3 gpurun_out/bench_default.err
cat gpurun_out/summary.txt
