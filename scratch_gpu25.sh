#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/debug_mixed2.py tf32_legfp32,tf32,fp32 > gpurun_out/dbg1.log 2>&1; tail -4 gpurun_out/dbg1.log
timeout 600 python tools/debug_mixed2.py fp32,tf32,tf32_legfp32 > gpurun_out/dbg2.log 2>&1; tail -4 gpurun_out/dbg2.log
