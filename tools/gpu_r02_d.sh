#!/bin/bash
# Round 2, GPU visit D: whole GPU suite, smoke, bench line (ShardedBatchBeamformer e2e arm, copy-only ceiling, parity at bench size), reference arm.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -rfE 2>&1 | tail -15 > gpurun_out/d_pytest_gpu.txt; cat gpurun_out/d_pytest_gpu.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 900 python bench.py --gpus 1 --steps 10 --warmup 3 > gpurun_out/d_bench.json 2> gpurun_out/d_bench.err; tail -5 gpurun_out/d_bench.err; cat gpurun_out/d_bench.json
timeout 600 python bench.py --impl reference --gpus 1 --steps 2 --warmup 1 > gpurun_out/d_bench_ref.json 2>> gpurun_out/d_bench.err; cat gpurun_out/d_bench_ref.json
