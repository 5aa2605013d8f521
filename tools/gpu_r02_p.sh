#!/bin/bash
# Round 2, GPU visit P: warp-per-chain Cholesky for the wide MVDR solve — tests, then configs[3] + covariance timing (A/B with BTKB_SOLVE_CHOL=0).
python -c "from distant_speech_recognition_b200 import _capi" || exit 1
timeout 900 python -m pytest tests/test_parity_gpu_r2.py tests/test_parity_gpu.py -m gpu -q -k "wide_mvdr or 64_mic" 2>&1 | tail -12
for c in 1 0; do echo "== BTKB_SOLVE_CHOL=$c"; BTKB_SOLVE_CHOL=$c timeout 600 python tools/bench_cov64.py 2>&1 | tail -3; done
