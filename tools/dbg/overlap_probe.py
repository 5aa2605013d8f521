#!/usr/bin/env python
"""Does running the three kernels of different sub-batches concurrently (one stream per sub-batch) beat the serial step?  K1 / K5 are
bound by the shared-memory data pipe, K4 by HBM: complementary resources.   python tools/dbg/overlap_probe.py [steps]"""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))
from distant_speech_recognition_b200 import _capi
from bench_configs import proto, tiled_batch

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 30
C, M, U, n = 8, 512, 256, 80000
h, g = proto(M); x, d = tiled_batch(U, C, n, 16)
out = {}
for NP in [int(v) for v in os.environ.get("PROBE_NP", "1,2,4,8").split(",")]:
    Us = U // NP
    pipes = []
    for i in range(NP):
        p = _capi.Pipeline(C, M, 4, 1, beamformer=_capi.BF_GSC_LMS, max_utterances=Us, max_samples=n)
        p.set_prototypes(h, g); p.set_delays(d[i * Us:(i + 1) * Us]); p.submit(np.ascontiguousarray(x[i * Us:(i + 1) * Us])); p.synchronize()
        pipes.append(p)
    for _ in range(3):
        for p in pipes: p.run(True)
    for p in pipes: p.synchronize()
    t0 = time.perf_counter()
    for _ in range(steps):
        for p in pipes: p.run(True)
    for p in pipes: p.synchronize()
    dt = (time.perf_counter() - t0) / steps
    out["NP=%d" % NP] = {"ms_per_step_wall": 1e3 * dt, "sum_of_kernel_ms": sum(p.last_timing()["total_ms"] for p in pipes)}
    for p in pipes: p.close()
print(json.dumps(out))
