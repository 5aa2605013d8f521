#!/usr/bin/env python
"""64-mic covariance at configs[3] size (256 x 5 s, M = 512): device time of btkb_accumulate_covariance.
BTKB_COV_TC=0 python tools/bench_cov64.py  -> CUDA-core kernel;  default -> tcgen05 kernel."""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from distant_speech_recognition_b200 import _capi
from bench_configs import proto, tiled_batch, timed

C, M, U, n = 64, 512, 256, 80000
h, g = proto(M); x, d = tiled_batch(U, C, n, 4)
p = _capi.Pipeline(C, M, 4, 1, beamformer=_capi.BF_MVDR, max_utterances=U, max_samples=n)
p.set_prototypes(h, g); p.set_delays(d); p.submit(x); p.run_analysis(); p.synchronize()
T = p.num_frames; K = M // 2 + 1
for name, labels in (("all frames (label never fires)", np.tile(np.array([[100.0, 200.0]]), (U, 1))), ("noise frames outside [1, 5] s", np.tile(np.array([[1.0, 5.0]]), (U, 1)))):
    def f(): p.accumulate_covariance(labels, 10.0); p.synchronize()
    s = timed(f, steps=3, warm=1)
    xbytes = C * K * 8 * U * T; rbytes = C * C * K * 8 * U
    print(json.dumps({"cov64 " + name: dict(tc=os.environ.get("BTKB_COV_TC", "1"), ms=1e3 * s, hbm_frac_read_X_plus_write_R=(xbytes + rbytes) / s / 1e9 / 6530.3,
                                            tflops_complex_gram=8.0 * C * C * K * U * T / s / 1e12)}))
s = timed(lambda: (p.calc_mvdr_weights(1e-4), p.synchronize()), steps=3, warm=1)
solver = os.environ.get("BTKB_SOLVE_CHOL", "auto")
print(json.dumps({"mvdr solve 64 x 64, 65 792 matrices, covariance of 62 noise frames + 1e-4 (rank deficient, numerically not positive definite)": dict(ms=1e3 * s, BTKB_SOLVE_CHOL=solver)}))
p.accumulate_covariance(np.tile(np.array([[100.0, 200.0]]), (U, 1)), 10.0)
s = timed(lambda: (p.calc_mvdr_weights(1e-4), p.synchronize()), steps=3, warm=1)
print(json.dumps({"mvdr solve 64 x 64, 65 792 matrices, covariance of all 317 frames + 1e-4 (positive definite)": dict(ms=1e3 * s, BTKB_SOLVE_CHOL=solver)}))
