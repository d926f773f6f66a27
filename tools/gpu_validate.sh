#!/bin/bash
# What the driver runs at round end, on ONE B200: GPU tests, smoke, the default bench line (both tiers + CPU baseline) and
# the reference arm.  Outputs under gpurun_out/ (copy what is to be judged into profiles/).
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --durations=5 2>&1 | grep -v "Warning\|amp.autocast\|^$" | tail -14 > gpurun_out/r02_pytest_gpu_final.txt; tail -4 gpurun_out/r02_pytest_gpu_final.txt
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee gpurun_out/r02_smoke_final.txt
timeout 900 python bench.py > gpurun_out/r02_bench_default_final.json 2> gpurun_out/r02_bench_default_final.err; tail -c 400 gpurun_out/r02_bench_default_final.json; echo
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02_bench_reference_final.json 2> gpurun_out/r02_bench_reference_final.err; tail -c 600 gpurun_out/r02_bench_reference_final.json; echo
