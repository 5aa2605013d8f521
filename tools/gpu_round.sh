#!/bin/bash
# One GPU visit: parity report, gpu tests, bench, ncu launch list + full capture of the dominant kernel.
mkdir -p gpurun_out
python tools/parity_report.py > gpurun_out/parity.json 2> gpurun_out/parity.err
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -15 > gpurun_out/pytest_gpu.txt
timeout 900 python bench.py --gpus 1 --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err
tail -3 gpurun_out/bench.err
cat gpurun_out/bench.json
timeout 600 python bench.py --impl reference --gpus 1 --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2>> gpurun_out/bench.err
cat gpurun_out/bench_ref.json
