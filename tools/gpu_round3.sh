#!/bin/bash
# Final GPU visit of the round: gpu tests, parity report, bench (+ reference arm), full-size configs, ncu launch list of the bench.
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -5 > gpurun_out/pytest_gpu.txt; cat gpurun_out/pytest_gpu.txt
timeout 300 python tools/parity_report.py > gpurun_out/parity.json 2> gpurun_out/parity.err; tail -2 gpurun_out/parity.err
timeout 400 python bench.py --gpus 1 --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -3 gpurun_out/bench.err; cat gpurun_out/bench.json
timeout 400 python bench.py --impl reference --gpus 1 --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2>> gpurun_out/bench.err; cat gpurun_out/bench_ref.json
timeout 600 python tools/bench_configs.py > gpurun_out/configs.json 2> gpurun_out/configs.err; tail -3 gpurun_out/configs.err; cat gpurun_out/configs.json
timeout 300 python tools/bench_cov64.py > gpurun_out/cov64.json 2>&1; cat gpurun_out/cov64.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches.csv python bench.py --gpus 1 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
ls -la gpurun_out | head -40
