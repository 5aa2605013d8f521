#!/bin/bash
# quick iteration: gpu tests + bench variants (no CPU baseline)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -8
for fr in 12 16; do
  echo "== BTKB_ANALYSIS_FR=$fr"
  BTKB_ANALYSIS_FR=$fr timeout 600 python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline 2>gpurun_out/bench_fr$fr.err | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print({k:d[k] for k in ('value','ms_per_step','kernel_ms_per_step')}, d['roofline']['all_kernels_frac'], d['e2e']['ms_per_step'])"
done
