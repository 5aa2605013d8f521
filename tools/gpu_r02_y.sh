#!/bin/bash
# round 2, call y: WPE Cholesky with the fp64 tensor-core trailing update (planar swizzled panel); boundary tests
cd /root/repo
mkdir -p gpurun_out
python -c "from distant_speech_recognition_b200 import _capi" || exit 1
timeout 600 python -m pytest tests/test_parity_gpu_r2.py tests/test_parity_gpu.py tests/test_zz_host_surface.py tests/test_btk20_api.py -q -x -m gpu --timeout 120 -k "wpe or dereverb or mvdrgsc or input_source_vector" 2>&1 | tail -15 > gpurun_out/y_tests.txt
cat gpurun_out/y_tests.txt
: > gpurun_out/y_wpe.jsonl
timeout 300 python tools/bench_wpe.py >> gpurun_out/y_wpe.jsonl 2> gpurun_out/y_wpe.err
for kn in "BTKB_WPE_MMA=0"; do
  env $kn WPE_PREC=fp64 timeout 300 python tools/bench_wpe.py >> gpurun_out/y_wpe.jsonl 2>> gpurun_out/y_wpe.err
done
cat gpurun_out/y_wpe.jsonl; tail -3 gpurun_out/y_wpe.err
WPE_FORMS=frame WPE_PREC=fp64 WPE_U=2 timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_wpe_chol" -c 2 -o gpurun_out/y_ncu_wpe_chol -f python tools/bench_wpe.py > gpurun_out/y_ncu.log 2>&1
ncu -i gpurun_out/y_ncu_wpe_chol.ncu-rep --page details 2>/dev/null > gpurun_out/y_ncu_wpe_chol_details.txt
tail -2 gpurun_out/y_ncu.log
