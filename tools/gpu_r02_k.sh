#!/bin/bash
# Round 2, GPU visit K: per-group named barriers in the K1 / K5 transforms — whole suite, racecheck of a small pipe, step timing.
mkdir -p gpurun_out
python -c "from distant_speech_recognition_b200 import _capi" || exit 1
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -6
timeout 900 compute-sanitizer --tool racecheck --print-limit 5 python -m pytest tests/test_parity_gpu.py -m gpu -q -x -k "gsc_lms_pipe_golden or ds_pipe_golden" 2>&1 | grep -E "RACECHECK|Error|hazard|passed|failed" | head -12
for i in 1 2; do timeout 300 python tools/prof_step.py 10 | tee -a gpurun_out/k_step.jsonl; done
