#!/usr/bin/env python
"""A/B check at scale: 64-mic covariance of a 64-utterance batch (16 448 chains, 317 frames) with the tcgen05 kernel and with the
CUDA-core kernel (BTKB_COV_TC=0), per-chain comparison.  python tools/dbg/cov_ab.py write|compare"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))
from distant_speech_recognition_b200 import _capi
from bench_configs import proto, tiled_batch
C, M, U, n = 64, 512, 64, 80000
h, g = proto(M); x, d = tiled_batch(U, C, n, 8)
p = _capi.Pipeline(C, M, 4, 1, beamformer=_capi.BF_MVDR, max_utterances=U, max_samples=n)
p.set_prototypes(h, g); p.set_delays(d); p.submit(x); p.run_analysis()
labels = np.stack([np.array([0.5 + 0.05 * (u % 7), 2.0 + 0.1 * (u % 5)]) for u in range(U)])
out = []
for rep in range(3):
    p.accumulate_covariance(labels, 10.0)
    out.append(p.get_covariance())
assert np.array_equal(out[0], out[1]) and np.array_equal(out[0], out[2]), "not reproducible across calls"
path = os.path.join(ROOT, "gpurun_out", "cov_ab_%s.npy" % os.environ.get("BTKB_COV_TC", "1"))
if sys.argv[1] == "write":
    np.save(path, out[0])
else:
    ref = np.load(os.path.join(ROOT, "gpurun_out", "cov_ab_0.npy"))
    R = out[0].astype(np.complex128); Q = ref.astype(np.complex128)
    err = np.sqrt(np.sum(np.abs(R - Q) ** 2, axis=(2, 3)) / np.maximum(np.sum(np.abs(Q) ** 2, axis=(2, 3)), 1e-300))
    print("tcgen05 vs CUDA-core covariance over %d chains: max rel err %.2e, median %.2e, chains > 1e-5: %d" % (err.size, err.max(), np.median(err), int((err > 1e-5).sum())))
