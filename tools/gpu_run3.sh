mkdir -p gpurun_out
nproc; free -g | head -2
timeout 2400 python -m pytest tests -m gpu -q --durations=15 2>&1 | tail -60 > gpurun_out/r02_test_gpu_all.log; tail -45 gpurun_out/r02_test_gpu_all.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke.log 2>&1; tail -3 gpurun_out/r02_smoke.log
timeout 900 python bench.py > gpurun_out/r02_bench_both_tiers.json 2> gpurun_out/r02_bench_both_tiers.err; tail -c 3000 gpurun_out/r02_bench_both_tiers.json; tail -5 gpurun_out/r02_bench_both_tiers.err
