#!/bin/bash
# round 2, call s: the frame-domain WPE form: parity tests, timing against the lag-domain form, ncu of the S x S Cholesky
cd /root/repo
mkdir -p gpurun_out
python -c "from distant_speech_recognition_b200 import _capi" || exit 1
timeout 900 python -m pytest tests/test_parity_gpu_r2.py tests/test_parity_gpu.py tests/test_zz_host_surface.py -q -x -m gpu -k "wpe" 2>&1 | tail -15 > gpurun_out/s_tests.txt
cat gpurun_out/s_tests.txt
timeout 600 python tools/bench_wpe.py > gpurun_out/s_wpe.jsonl 2> gpurun_out/s_wpe.err
cat gpurun_out/s_wpe.jsonl; tail -3 gpurun_out/s_wpe.err
WPE_FORMS=frame WPE_U=2 timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_wpe_chol|k_wpe_gram_dual" -c 4 -o gpurun_out/s_ncu_wpe_dual -f python tools/bench_wpe.py > gpurun_out/s_ncu.log 2>&1
ncu -i gpurun_out/s_ncu_wpe_dual.ncu-rep --page details 2>/dev/null > gpurun_out/s_ncu_wpe_dual_details.txt
tail -2 gpurun_out/s_ncu.log
