#!/bin/bash
# Round 2, GPU visit AL: narrow MVDR solve with the cheap singular-value bracket (configs[2] regression), the per-config bench of every BASELINE config
cd /root/repo
mkdir -p gpurun_out
python -c "from distant_speech_recognition_b200 import _capi" || exit 1
timeout 600 python -m pytest tests -m gpu -q --timeout 300 -k "mvdr or smi or threshold or cfg3 or zelinski" 2>&1 | tail -6 | tee gpurun_out/al_tests.txt
timeout 900 python tools/bench_configs.py > gpurun_out/al_configs_full_size.json 2> gpurun_out/al_err.txt; tail -3 gpurun_out/al_err.txt
python - <<'PY'
import json
d = json.load(open("gpurun_out/al_configs_full_size.json"))
for k, v in d.items():
    print(k, {kk: (round(vv, 3) if isinstance(vv, float) else vv) for kk, vv in v.items() if kk in ("ms", "frames_per_s", "covariance_ms", "solve_ms", "apply_synthesis_ms", "wpe_ms")}, (v.get("parity_check") or {}).get("pass"))
PY
