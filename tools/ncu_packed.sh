#!/bin/bash
# Round 2, after tools/gpu_r02_first.sh has shown the packed 2 x fp32 kernels bit-identical and faster: the same ncu evidence as
# tools/ncu_round.sh for the packed variants (never a bench value) — launch list + one --set full capture per hot kernel, to be
# summarised under profiles/r02a_* next to the r01d captures of the scalar kernels (issue slots busy, FMA pipe, registers, IPC).
mkdir -p gpurun_out
export BTKB_ANALYSIS_PACKED=1 BTKB_SYNTHESIS_PACKED=1 BTKB_PERBIN_PACKED=1
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_packed.csv \
    python bench.py --gpus 1 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_packed_under_ncu.log 2>&1
for k in k_analysis k_perbin k_synthesis; do
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:$k -s 3 -c 1 -f -o gpurun_out/prof_packed_$k \
    python bench.py --gpus 1 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_packed_$k.log 2>&1
done
ls -la gpurun_out
