#!/bin/bash
# Round 2, GPU visit I: the fused analysis + NLMS kernel — identity with the two-kernel path, then its time at configs[1] size.
mkdir -p gpurun_out
python -c "from distant_speech_recognition_b200 import _capi" || exit 1
timeout 600 python -m pytest tests/test_parity_gpu_r2.py -m gpu -q -k fused 2>&1 | tail -12
for f in 0 1; do echo "== BTKB_FUSED=$f"; BTKB_FUSED=$f timeout 300 python tools/prof_step.py 10 | tee -a gpurun_out/i_fused.jsonl; done
BTKB_FUSED=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_fused -s 2 -c 1 -f -o gpurun_out/i_prof_k_fused python tools/prof_step.py 1 > gpurun_out/i_ncu.log 2>&1
