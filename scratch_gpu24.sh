#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/debug_mixed.py > gpurun_out/debug_mixed.log 2>&1; cat gpurun_out/debug_mixed.log | tail -12
