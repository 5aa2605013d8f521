#!/bin/bash
# ncu evidence (never a bench value): launch list of one short bench run + one --set full capture of each hot kernel.
mkdir -p gpurun_out
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches.csv \
    python bench.py --gpus 1 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
for k in k_perbin k_analysis k_synthesis; do
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:$k -s 3 -c 1 -f -o gpurun_out/prof_$k \
    python bench.py --gpus 1 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_$k.log 2>&1
done
ls -la gpurun_out
