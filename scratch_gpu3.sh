#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_tc.py -q -m gpu --tb=short -x > gpurun_out/test_gpu_tc.log 2>&1; echo "tc exit $?" >> gpurun_out/summary.txt
tail -30 gpurun_out/test_gpu_tc.log
timeout 300 python tools/time_stages.py --precision tf32 > gpurun_out/stages_tf32.json 2> gpurun_out/stages_tf32.err; echo "stages exit $?" >> gpurun_out/summary.txt
cat gpurun_out/stages_tf32.json; tail -3 gpurun_out/stages_tf32.err
cat gpurun_out/summary.txt
