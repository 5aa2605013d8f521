#!/bin/bash
for dbg in 0 1 2 4 3 7; do
  echo "== BTKB_ANALYSIS_DEBUG=$dbg"
  BTKB_ANALYSIS_DEBUG=$dbg timeout 300 python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['kernel_ms_per_step']['analysis_ms'])"
done
