#!/usr/bin/env python
"""configs[4] sample (8 mics, M = 1024, confs/wpe.json: 33 lags, 2 iterations): device time of the WPE pass for U utterances."""
import json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from distant_speech_recognition_b200 import _capi
from bench_configs import proto, tiled_batch, timed
C, M, U, n = 8, 1024, int(os.environ.get("WPE_U", "8")), 80000
h, g = proto(M); x, d = tiled_batch(U, C, n, 4)
for tag, fp32 in (("fp64", 0), ("fp32", 1)):
    wpe = dict(lower_num=0, upper_num=32, iterations_num=2, load_db=-18.0, band_width=0.0, diagonal_bias=1e-4, fp32_normal_equations=fp32)
    p = _capi.Pipeline(C, M, 4, 1, beamformer=_capi.BF_GSC_LMS, max_utterances=U, max_samples=n, wpe=wpe)
    p.set_prototypes(h, g); p.set_delays(d); p.submit(x); p.synchronize()
    s = timed(lambda: (p.run(True), p.synchronize()), steps=1, warm=1)
    print(json.dumps({"wpe chain %s, %d utterances" % (tag, U): dict(ms=1e3 * s, wpe_ms=p.last_timing_wpe(), chunk=os.environ.get("BTKB_WPE_CHUNK", "55 (default)"), s_per_1024_utt=s * 1024 / U)}))
    p.close()
