#!/usr/bin/env python
"""Small driver for ncu captures and quick A/B timings of the headline step (configs[1]: 8 mics, M = 512, 256 x 5 s, GSC-NLMS):
   python tools/prof_step.py [steps] [distinct]     prints one JSON line with the CUDA-event kernel times (median over the steps).
Kernel variants are selected by the BTKB_* environment variables the library reads at every launch.  Never a bench value when run
under ncu."""
import json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))
from distant_speech_recognition_b200 import _capi
from bench_configs import proto, tiled_batch


def main():
    steps = int(sys.argv[1]) if len(sys.argv) > 1 else 5
    distinct = int(sys.argv[2]) if len(sys.argv) > 2 else 16
    C, M, U, n = int(os.environ.get("PROF_C", 8)), int(os.environ.get("PROF_M", 512)), int(os.environ.get("PROF_U", 256)), 80000
    h, g = proto(M); x, d = tiled_batch(U, C, n, distinct)
    p = _capi.Pipeline(C, M, 4, 1, beamformer=_capi.BF_GSC_LMS, max_utterances=U, max_samples=n)
    p.set_prototypes(h, g); p.set_delays(d)
    if os.environ.get("PROF_I16", "0") != "0":
        p.submit_i16(x.astype(np.int16))   # 16-bit PCM resident in HBM (PROF_I16=1); default float32, the bench's device-resident arm
    else:
        p.submit(x)
    p.synchronize()
    rows = []
    for i in range(steps + 2):
        p.run(True); p.synchronize()
        if i >= 2: rows.append(p.last_timing())
    med = {k: float(np.median([r[k] for r in rows])) for k in ("total_ms", "analysis_ms", "perbin_ms", "synthesis_ms")}
    med["frames"] = U * p.num_frames
    med["env"] = {k: v for k, v in os.environ.items() if k.startswith("BTKB_")}
    print(json.dumps(med))


if __name__ == "__main__":
    main()
