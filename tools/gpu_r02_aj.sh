#!/bin/bash
# Round 2, GPU visit AJ: wide MVDR solve with the list-driven LU fallback and the per-call solver choice — tests, then configs[3] timing
cd /root/repo
mkdir -p gpurun_out
python -c "from distant_speech_recognition_b200 import _capi" || exit 1
timeout 600 python -m pytest tests/test_parity_gpu_r2.py tests/test_parity_gpu.py tests/test_zz_host_surface.py -m gpu -q --timeout 120 -k "wide or 64_mic or mvdr" 2>&1 | tail -12 | tee gpurun_out/aj_tests.txt
: > gpurun_out/aj_solve.txt
echo "== default (auto)" >> gpurun_out/aj_solve.txt; timeout 300 python tools/bench_cov64.py 2>&1 | tail -3 >> gpurun_out/aj_solve.txt
for c in 2 1 0; do echo "== BTKB_SOLVE_CHOL=$c" >> gpurun_out/aj_solve.txt; BTKB_SOLVE_CHOL=$c timeout 300 python tools/bench_cov64.py 2>&1 | tail -2 >> gpurun_out/aj_solve.txt; done
cat gpurun_out/aj_solve.txt | cut -c1-400
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_parity_gpu_r2.py -q -x -m gpu -k "wide_mvdr" > gpurun_out/aj_memcheck.txt 2>&1; echo "memcheck rc=$?" >> gpurun_out/aj_memcheck.txt
tail -4 gpurun_out/aj_memcheck.txt
timeout 600 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_parity_gpu_r2.py -q -x -m gpu -k "wide_mvdr and 64" > gpurun_out/aj_racecheck.txt 2>&1; echo "racecheck rc=$?" >> gpurun_out/aj_racecheck.txt
tail -4 gpurun_out/aj_racecheck.txt
