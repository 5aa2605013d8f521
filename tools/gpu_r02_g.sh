#!/bin/bash
# Round 2, GPU visit G: K1 reading 16-bit PCM directly — parity / identity tests, A/B against the float path, ncu, bench line.
mkdir -p gpurun_out
python -c "from distant_speech_recognition_b200 import _capi" || exit 1
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -8 | tee gpurun_out/g_pytest_gpu.txt
for i16 in 0 1; do echo "== PROF_I16=$i16"; PROF_I16=$i16 timeout 300 python tools/prof_step.py 10 | tee -a gpurun_out/g_i16_ab.jsonl; done
timeout 900 python bench.py --gpus 1 --steps 10 --warmup 3 > gpurun_out/g_bench.json 2> gpurun_out/g_bench.err; tail -3 gpurun_out/g_bench.err; cat gpurun_out/g_bench.json | cut -c1-1500
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_analysis -s 2 -c 1 -f -o gpurun_out/g_prof_k_analysis_i16 python tools/prof_step.py 1 > gpurun_out/g_ncu.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 30 --csv --log-file gpurun_out/g_launches.csv python tools/prof_step.py 3 > gpurun_out/g_under_ncu.log 2>&1
ls -la gpurun_out | tail -8
