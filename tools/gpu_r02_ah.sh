#!/bin/bash
# Round 2, GPU visit AH: blocked tensor-core Cholesky for the wide MVDR solve (BTKB_SOLVE_CHOL=2) — tests, then configs[3] timing A/B
cd /root/repo
mkdir -p gpurun_out
python -c "from distant_speech_recognition_b200 import _capi" || exit 1
timeout 600 python -m pytest tests/test_parity_gpu_r2.py tests/test_parity_gpu.py -m gpu -q --timeout 120 -k "wide_mvdr or 64_mic" 2>&1 | tail -12 | tee gpurun_out/ah_tests.txt
: > gpurun_out/ah_solve.txt
for c in 2 1 0; do echo "== BTKB_SOLVE_CHOL=$c" >> gpurun_out/ah_solve.txt; BTKB_SOLVE_CHOL=$c timeout 300 python tools/bench_cov64.py 2>&1 | tail -3 >> gpurun_out/ah_solve.txt; done
cat gpurun_out/ah_solve.txt | cut -c1-700
