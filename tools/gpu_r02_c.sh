#!/bin/bash
# Round 2, GPU visit C: whole GPU suite after the streaming / arbitrary-C / generic-filter-bank work, then the step timing.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -rfE 2>&1 | tail -60 > gpurun_out/c_pytest_gpu.txt; tail -50 gpurun_out/c_pytest_gpu.txt
timeout 300 python tools/prof_step.py 10 | tee gpurun_out/c_step.json
