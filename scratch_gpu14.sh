#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/conv_experiments.py > gpurun_out/conv_exp.json 2> gpurun_out/conv_exp.err; cat gpurun_out/conv_exp.json; tail -3 gpurun_out/conv_exp.err
