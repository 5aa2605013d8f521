#!/bin/bash
# One GPU visit (round 1, second session): gpu tests, bench, full-size configs, timings + ncu of the new kernels.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -25 > gpurun_out/pytest_gpu.txt
cat gpurun_out/pytest_gpu.txt | tail -8
timeout 400 python bench.py --gpus 1 --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err
tail -3 gpurun_out/bench.err; cat gpurun_out/bench.json
timeout 500 python tools/bench_configs.py > gpurun_out/configs.json 2> gpurun_out/configs.err
tail -3 gpurun_out/configs.err; cat gpurun_out/configs.json
timeout 300 python tools/profile_extra.py sos > gpurun_out/sos_timing.json 2> gpurun_out/sos.err
tail -3 gpurun_out/sos.err; cat gpurun_out/sos_timing.json
for k in k_wpe_corr k_wpe_chol; do
timeout 300 ncu --set full --clock-control none --import-source on -k regex:$k -s 1 -c 1 -f -o gpurun_out/prof_$k python tools/profile_extra.py wpe > gpurun_out/ncu_$k.log 2>&1
done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_sos_cov -s 2 -c 1 -f -o gpurun_out/prof_k_sos_cov python tools/profile_extra.py sos > gpurun_out/ncu_k_sos_cov.log 2>&1
ls -la gpurun_out
