#!/bin/bash
# Round 2, GPU visit E: the fp64 C++ SubbandGSCRLS kernel, then the occupancy experiment that backs DESIGN.md §10's fusion note:
# K1 with 3 (default), 2 and 1 CTAs per SM (unused shared-memory padding), i.e. with the warps a fused kernel could give its FFT part.
mkdir -p gpurun_out
python -c "from distant_speech_recognition_b200.btk20.beamformer import SubbandGSCRLSPtr" || exit 1
timeout 800 python -m pytest tests/test_parity_gpu_r2.py tests/test_btk20_api.py -m gpu -q -k "cpp_subband" 2>&1 | tail -15
for pad in 0 40000 120000; do echo "== BTKB_ANALYSIS_SMEM_PAD=$pad"; BTKB_ANALYSIS_SMEM_PAD=$pad timeout 300 python tools/prof_step.py 10 | tee -a gpurun_out/e_occupancy.jsonl; done
