mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_sht.py tests/test_gpu_parity_r2.py -x -q -k "sht or sharded or 2880" 2>&1 | tail -3
python tools/time_sharded_stages.py > gpurun_out/r02_sharded_stages_n1_v3.json 2> gpurun_out/r02_sharded_stages_n1.err; cat gpurun_out/r02_sharded_stages_n1_v3.json | tr -d '\n' | cut -c1-900; echo; tail -3 gpurun_out/r02_sharded_stages_n1.err
