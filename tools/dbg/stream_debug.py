#!/usr/bin/env python
"""Where does a streamed run differ from the whole-utterance run?  python tools/dbg/stream_debug.py [kind]"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from distant_speech_recognition_b200 import _capi as capi, synthetic
kind = sys.argv[1] if len(sys.argv) > 1 else "ds"
M, m, r, C, U = 256, 4, 1, 4, 3
D = M >> r
pr = np.load(os.path.join(ROOT, "tests", "golden", "prototype_M256_m4_r1.npz")); h, g = pr["h"], pr["g"]
cuts = np.cumsum([0, 2, 1, 17, 2, 40]) * D
n_final = 9 * D
tail = np.array([n_final, 3 * D + 17, 0], np.int32)
n = int(cuts[-1]) + n_final
x, d = synthetic.make_batch(U, C, n, first=700)
lengths = (cuts[-1] + tail).astype(np.int32)
kw = dict(ds=dict(beamformer=capi.BF_DS), nlms=dict(beamformer=capi.BF_GSC_LMS, lms=dict(min_frames=9, slowdown_after=16)))[kind]
w = capi.Pipeline(C, M, m, r, max_utterances=U, max_samples=n, **kw); w.set_prototypes(h, g); w.set_delays(d); w.submit(x, lengths); w.run(True)
RX, RY, Ry = w.fetch_snapshots(), w.fetch_subband(), w.fetch_time(); w.close()
p = capi.Pipeline(C, M, m, r, max_utterances=U, max_samples=41 * D, **kw); p.set_prototypes(h, g); p.set_delays(d); p.stream_begin(U)
bounds = list(cuts) + [n]
t_off = b_off = 0
for j in range(len(bounds) - 1):
    a, b = int(bounds[j]), int(bounds[j + 1]); final = j == len(bounds) - 2
    xc = np.ascontiguousarray(x[:, :, a:b])
    p.stream_submit(xc, tail if final else None, final=final); p.synchronize()
    T, nb = p.num_frames, p.num_blocks
    print("chunk", j, "blocks", (b - a) // D, "-> frames", T, "blocks out", nb, "pos", p.stream_position())
    if T > 0:
        X, Y = p.fetch_snapshots(), p.fetch_subband()
        for u in range(U):
            bx = [t for t in range(T) if not np.array_equal(X[u, t].view(np.uint32), RX[u, t_off + t].view(np.uint32))]
            by = [t for t in range(T) if not np.array_equal(Y[u, t].view(np.uint32), RY[u, t_off + t].view(np.uint32))]
            if bx or by:
                t = (bx or by)[0]
                print("  u", u, "X frames differ:", bx[:8], "Y frames differ:", by[:8], " max |dX| at first:", np.abs(X[u, t] - RX[u, t_off + t]).max(), "ref max", np.abs(RX[u, t_off + t]).max())
    if nb > 0:
        y = p.fetch_time()
        for u in range(U):
            bb = [t for t in range(nb) if not np.array_equal(y[u, t * D:(t + 1) * D].view(np.uint32), Ry[u, (b_off + t) * D:(b_off + t + 1) * D].view(np.uint32))]
            if bb:
                print("  u", u, "time blocks differ:", bb[:8])
    t_off += T; b_off += nb
print("done")
