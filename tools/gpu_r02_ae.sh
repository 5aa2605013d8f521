#!/bin/bash
# Round 2, GPU visit AE (state after the WPE rework and the boundary completions): the whole GPU suite, smoke, bench (the NCCL-free
# N = 1 line) — the same checks the driver runs at round end.
cd /root/repo
mkdir -p gpurun_out
python -c "from distant_speech_recognition_b200 import _capi" || exit 1
timeout 1500 python -m pytest tests -m gpu -q --timeout 300 2>&1 | tail -6 | tee gpurun_out/ae_pytest_gpu.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee gpurun_out/ae_smoke.txt
timeout 900 python bench.py --gpus 1 --steps 10 --warmup 3 > gpurun_out/ae_bench.json 2> gpurun_out/ae_bench.err; tail -3 gpurun_out/ae_bench.err; cut -c1-600 gpurun_out/ae_bench.json
