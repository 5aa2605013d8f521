#!/bin/bash
# Round 2, GPU visit M (final evidence of the headline step): GPU suite, smoke, bench + reference arm, ncu launch list and --set full
# captures of the three kernels as they ship.
mkdir -p gpurun_out
python -c "from distant_speech_recognition_b200 import _capi" || exit 1
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -5 | tee gpurun_out/m_pytest_gpu.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py --gpus 1 --steps 10 --warmup 3 > gpurun_out/m_bench.json 2> gpurun_out/m_bench.err; tail -3 gpurun_out/m_bench.err; cut -c1-400 gpurun_out/m_bench.json
timeout 600 python bench.py --impl reference --gpus 1 --steps 2 --warmup 1 > gpurun_out/m_bench_ref.json 2>> gpurun_out/m_bench.err; cut -c1-300 gpurun_out/m_bench_ref.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 30 --csv --log-file gpurun_out/m_launches.csv python tools/prof_step.py 3 > gpurun_out/m_under_ncu.log 2>&1
for k in k_analysis k_perbin k_synthesis; do
timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 2 -c 1 -f -o gpurun_out/m_prof_$k python tools/prof_step.py 1 > gpurun_out/m_ncu_$k.log 2>&1
done
ls gpurun_out | grep "^m_"
