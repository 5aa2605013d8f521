#!/bin/bash
# round 2, call t: reworked WPE Cholesky (warp-parallel diagonal block, register TRSM, 128-thread CTAs in the frame-domain form)
cd /root/repo
mkdir -p gpurun_out
python -c "from distant_speech_recognition_b200 import _capi" || exit 1
timeout 900 python -m pytest tests/test_parity_gpu_r2.py tests/test_parity_gpu.py tests/test_zz_host_surface.py tests/test_btk20_api.py -q -x -m gpu -k "wpe or dereverb" 2>&1 | tail -15 > gpurun_out/t_tests.txt
cat gpurun_out/t_tests.txt
: > gpurun_out/t_wpe.jsonl
timeout 300 python tools/bench_wpe.py >> gpurun_out/t_wpe.jsonl 2> gpurun_out/t_wpe.err
for kn in "BTKB_WPE_CHOL_THREADS=256 BTKB_WPE_CHUNK_FRAME=37" "BTKB_WPE_CHOL_THREADS=256 BTKB_WPE_CHUNK_FRAME=55" "BTKB_WPE_CHOL_THREADS=128 BTKB_WPE_CHUNK_FRAME=55" "BTKB_WPE_CHOL_THREADS=128 BTKB_WPE_CHUNK_FRAME=148" "BTKB_WPE_CHOL_THREADS=64 BTKB_WPE_CHUNK_FRAME=74" "BTKB_WPE_CHOL_THREADS=192 BTKB_WPE_CHUNK_FRAME=55"; do
  env $kn WPE_FORMS=frame WPE_PREC=fp64 timeout 300 python tools/bench_wpe.py >> gpurun_out/t_wpe.jsonl 2>> gpurun_out/t_wpe.err
done
for kn in "BTKB_WPE_CHOL_THREADS=128" "BTKB_WPE_CHOL_THREADS=128 BTKB_WPE_CHUNK=49"; do
  env $kn WPE_FORMS=lag WPE_PREC=fp64 timeout 300 python tools/bench_wpe.py >> gpurun_out/t_wpe.jsonl 2>> gpurun_out/t_wpe.err
done
cat gpurun_out/t_wpe.jsonl; tail -3 gpurun_out/t_wpe.err
WPE_FORMS=frame WPE_PREC=fp64 WPE_U=2 timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_wpe_chol" -c 2 -o gpurun_out/t_ncu_wpe_chol -f python tools/bench_wpe.py > gpurun_out/t_ncu.log 2>&1
ncu -i gpurun_out/t_ncu_wpe_chol.ncu-rep --page details 2>/dev/null > gpurun_out/t_ncu_wpe_chol_details.txt
tail -2 gpurun_out/t_ncu.log
