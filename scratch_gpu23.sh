#!/bin/bash
mkdir -p gpurun_out
timeout 900 python tools/parity_report.py > gpurun_out/parity2.json 2> gpurun_out/parity2.err; cat gpurun_out/parity2.json; tail -3 gpurun_out/parity2.err
