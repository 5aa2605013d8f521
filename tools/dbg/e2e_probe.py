"""Where does the end-to-end step time go?  Variants of bench.py's e2e loop on configs[1]."""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import bench
from distant_speech_recognition_b200 import _capi
c = bench.CFG; U, C, n, M, m, r = c["U"], c["C"], c["n"], c["M"], c["m"], c["r"]
T = bench.frames_per_utt(n, M, m, r); h, g = bench.load_proto(M)
x_np, delays = bench.make_inputs(0)
x16 = torch.from_numpy(x_np).to(torch.int16).pin_memory(); del x_np
nb = (T - m * (1 << r) // 2) * (M >> r)
out_pin = torch.empty((U, nb), dtype=torch.float32).pin_memory()

def probe(NP, do_run=True, do_fetch=True, do_delays=True, steps=8):
    Us = U // NP
    subs = []
    for i in range(NP):
        q = _capi.Pipeline(C, M, m, r, beamformer=_capi.BF_GSC_LMS, lms={}, max_utterances=Us, max_samples=n); q.set_prototypes(h, g); q.set_delays(delays[i * Us:(i + 1) * Us]); subs.append(q)
    pending = [False] * NP
    def collect(i):
        q = subs[i]
        if do_fetch and do_run:
            q.fetch_time_into(out_pin[i * Us:(i + 1) * Us].data_ptr()); q.fetch_stats()
        else:
            q.synchronize()
        pending[i] = False
    def step():
        for i, q in enumerate(subs):
            if pending[i]: collect(i)
            q.submit_i16_pointer(x16[i * Us:(i + 1) * Us].data_ptr(), Us, n)
            if do_delays: q.set_delays(delays[i * Us:(i + 1) * Us])
            if do_run: q.run(True)
            pending[i] = True
    for _ in range(2): step()
    for i in range(NP):
        if pending[i]: collect(i)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(steps): step()
    for i in range(NP):
        if pending[i]: collect(i)
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / steps
    for q in subs: q.close()
    return 1e3 * dt

for NP in (2, 4, 8, 16):
    print("NP=%2d  full %.2f ms | no fetch %.2f | upload+convert only %.2f" % (NP, probe(NP), probe(NP, do_fetch=False), probe(NP, do_run=False, do_delays=False)))
