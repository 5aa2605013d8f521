#!/bin/bash
# round 2, call af: WPE Cholesky with look-ahead factorisation of the next diagonal block
cd /root/repo
mkdir -p gpurun_out
python -c "from distant_speech_recognition_b200 import _capi" || exit 1
timeout 600 python -m pytest tests/test_parity_gpu_r2.py tests/test_parity_gpu.py tests/test_zz_host_surface.py tests/test_btk20_api.py -q -x -m gpu --timeout 120 -k "wpe or dereverb" 2>&1 | tail -15 > gpurun_out/af_tests.txt
cat gpurun_out/af_tests.txt
: > gpurun_out/af_wpe.jsonl
WPE_PREC=fp64 timeout 300 python tools/bench_wpe.py >> gpurun_out/af_wpe.jsonl 2> gpurun_out/af_wpe.err
cat gpurun_out/af_wpe.jsonl; tail -3 gpurun_out/af_wpe.err
timeout 600 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_parity_gpu_r2.py -q -x -m gpu -k "wpe" > gpurun_out/af_racecheck.txt 2>&1; echo "racecheck rc=$?" >> gpurun_out/af_racecheck.txt
tail -4 gpurun_out/af_racecheck.txt
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_parity_gpu_r2.py tests/test_parity_gpu.py -q -x -m gpu -k "wpe" > gpurun_out/af_memcheck.txt 2>&1; echo "memcheck rc=$?" >> gpurun_out/af_memcheck.txt
tail -4 gpurun_out/af_memcheck.txt
