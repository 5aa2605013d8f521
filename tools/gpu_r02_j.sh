#!/bin/bash
# Round 2, GPU visit J: host-mirror WPE filter reuse, then the whole suite.
python -c "from distant_speech_recognition_b200 import _capi" || exit 1
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -25
