#!/bin/bash
# Round 2, GPU visit Q: implicit-pivoting LU (one barrier per column) for the wide MVDR solve — tests, racecheck, A/B timing.
python -c "from distant_speech_recognition_b200 import _capi" || exit 1
timeout 900 python -m pytest tests/test_parity_gpu_r2.py tests/test_parity_gpu.py -m gpu -q -k "wide_mvdr or 64_mic" 2>&1 | tail -8
timeout 900 compute-sanitizer --tool racecheck --print-limit 3 python -m pytest tests/test_parity_gpu_r2.py -m gpu -q -k "wide_mvdr and 16-0" 2>&1 | grep -E "RACECHECK|hazard|passed|failed" | head -5
for c in 1 0; do echo "== BTKB_SOLVE_IP=$c"; BTKB_SOLVE_IP=$c timeout 600 python tools/bench_cov64.py 2>&1 | grep "mvdr solve" | cut -c1-220; done
