#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_net.py tests/test_gpu_sht.py tests/test_gpu_tc.py -q -m gpu --tb=short > gpurun_out/t.log 2>&1; echo "tests exit $?" >> gpurun_out/summary.txt; tail -5 gpurun_out/t.log
timeout 900 python tools/parity_report.py > gpurun_out/parity2.json 2> gpurun_out/parity2.err; cat gpurun_out/parity2.json; tail -3 gpurun_out/parity2.err
cat gpurun_out/summary.txt
