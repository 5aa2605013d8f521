#!/bin/bash
# Round 2, GPU visit H: every channel count (WPE / SOS for C = 3, 6; wide arrays 9, 12, 20, 48), then the whole suite.
python -c "from distant_speech_recognition_b200 import _capi" || exit 1
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -25
