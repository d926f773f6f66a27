mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_x3.py -x -q 2>&1 | tail -30 > gpurun_out/r02_test_x3.log
cat gpurun_out/r02_test_x3.log
timeout 300 python tools/x3_bench.py > gpurun_out/r02_x3_bench.json 2> gpurun_out/r02_x3_bench.err; cat gpurun_out/r02_x3_bench.json
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r02_test_gpu_all.log; cat gpurun_out/r02_test_gpu_all.log
timeout 600 python bench.py --precision fp32 --no-cpu-baseline > gpurun_out/r02_bench_fp32_first.json 2> gpurun_out/r02_bench_fp32_first.err; cat gpurun_out/r02_bench_fp32_first.json
timeout 600 python bench.py --precision tf32 --no-cpu-baseline > gpurun_out/r02_bench_tf32_first.json 2> gpurun_out/r02_bench_tf32_first.err; cat gpurun_out/r02_bench_tf32_first.json
