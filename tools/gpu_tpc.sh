#!/bin/bash
timeout 600 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
for tpc in 1 2 4 8; do
  echo "== BTKB_ANALYSIS_TPC=$tpc"
  BTKB_ANALYSIS_TPC=$tpc timeout 300 python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['value'], d['kernel_ms_per_step'], d['e2e']['ms_per_step'])"
done
