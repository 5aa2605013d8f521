#!/bin/bash
# round 2, call aa: WPE Cholesky with the register-resident diagonal-block factorisation
cd /root/repo
mkdir -p gpurun_out
python -c "from distant_speech_recognition_b200 import _capi" || exit 1
timeout 600 python -m pytest tests/test_parity_gpu_r2.py tests/test_parity_gpu.py tests/test_zz_host_surface.py tests/test_btk20_api.py -q -x -m gpu --timeout 120 -k "wpe or dereverb" 2>&1 | tail -15 > gpurun_out/aa_tests.txt
cat gpurun_out/aa_tests.txt
: > gpurun_out/aa_wpe.jsonl
timeout 300 python tools/bench_wpe.py >> gpurun_out/aa_wpe.jsonl 2> gpurun_out/aa_wpe.err
cat gpurun_out/aa_wpe.jsonl; tail -3 gpurun_out/aa_wpe.err
WPE_FORMS=frame WPE_PREC=fp64 WPE_U=2 timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_wpe_chol" -c 2 -o gpurun_out/aa_ncu_wpe_chol -f python tools/bench_wpe.py > gpurun_out/aa_ncu.log 2>&1
ncu -i gpurun_out/aa_ncu_wpe_chol.ncu-rep --page details 2>/dev/null > gpurun_out/aa_ncu_wpe_chol_details.txt
tail -2 gpurun_out/aa_ncu.log
