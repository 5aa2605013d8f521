#!/bin/bash
# round 2, call u: WPE Cholesky with L1 prefetch of the trailing entries; chunk size of the frame-domain form
cd /root/repo
mkdir -p gpurun_out
python -c "from distant_speech_recognition_b200 import _capi" || exit 1
timeout 900 python -m pytest tests/test_parity_gpu_r2.py tests/test_parity_gpu.py tests/test_zz_host_surface.py tests/test_btk20_api.py -q -x -m gpu -k "wpe or dereverb" 2>&1 | tail -15 > gpurun_out/u_tests.txt
cat gpurun_out/u_tests.txt
: > gpurun_out/u_wpe.jsonl
for kn in "BTKB_WPE_PREFETCH=1" "BTKB_WPE_PREFETCH=0" "BTKB_WPE_CHUNK_FRAME=148" "BTKB_WPE_CHUNK_FRAME=148 BTKB_WPE_PREFETCH=0" "BTKB_WPE_CHUNK_FRAME=592"; do
  env $kn WPE_FORMS=frame WPE_PREC=fp64 timeout 300 python tools/bench_wpe.py >> gpurun_out/u_wpe.jsonl 2>> gpurun_out/u_wpe.err
done
for kn in "BTKB_WPE_PREFETCH=1" "BTKB_WPE_PREFETCH=0"; do
  env $kn WPE_FORMS=lag WPE_PREC=fp64 timeout 300 python tools/bench_wpe.py >> gpurun_out/u_wpe.jsonl 2>> gpurun_out/u_wpe.err
done
WPE_FORMS=frame WPE_PREC=fp32 timeout 300 python tools/bench_wpe.py >> gpurun_out/u_wpe.jsonl 2>> gpurun_out/u_wpe.err
cat gpurun_out/u_wpe.jsonl; tail -3 gpurun_out/u_wpe.err
WPE_FORMS=frame WPE_PREC=fp64 WPE_U=2 timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_wpe_chol" -c 2 -o gpurun_out/u_ncu_wpe_chol -f python tools/bench_wpe.py > gpurun_out/u_ncu.log 2>&1
ncu -i gpurun_out/u_ncu_wpe_chol.ncu-rep --page details 2>/dev/null > gpurun_out/u_ncu_wpe_chol_details.txt
tail -2 gpurun_out/u_ncu.log
