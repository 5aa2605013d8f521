#!/usr/bin/env python
"""Streamed chunks at configs[1] shape (8 mics, M = 512, GSC-NLMS, 256 utterances advancing in lockstep): what does frame-batch
granularity cost?  For chunk sizes of 1 .. 64 blocks (16 ms .. 1 s of audio per utterance): host-to-result latency of one chunk through the
C-ABI (pageable int16 PCM up, subband + time signal left on the device, stream synchronised), the resulting frames/s and the real-time
factor, next to the whole-utterance step.   python tools/bench_stream.py"""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))
from distant_speech_recognition_b200 import _capi
from bench_configs import proto, tiled_batch

C, M, U, n, D = 8, 512, 256, 80000, 256
h, g = proto(M); x, d = tiled_batch(U, C, n, 16)
x16 = np.ascontiguousarray(x.astype(np.int16))
out = {"workload": "configs[1] shape, %d utterances in lockstep, 16-bit PCM chunks from pageable host memory" % U}
p = _capi.Pipeline(C, M, 4, 1, beamformer=_capi.BF_GSC_LMS, max_utterances=U, max_samples=n)
p.set_prototypes(h, g); p.set_delays(d)
p.submit_i16(x16); p.run(True); p.synchronize()
t0 = time.perf_counter()
for _ in range(3):
    p.submit_i16(x16); p.run(True); p.synchronize()
whole = (time.perf_counter() - t0) / 3
T = p.num_frames
out["whole utterances"] = {"ms_per_pass_incl_upload": 1e3 * whole, "frames_per_s": U * T / whole}
p.close()
for blocks in (1, 4, 16, 64):
    nc = blocks * D
    p = _capi.Pipeline(C, M, 4, 1, beamformer=_capi.BF_GSC_LMS, max_utterances=U, max_samples=nc)
    p.set_prototypes(h, g); p.set_delays(d)
    nchunks = min(n // nc, 40)
    lat = []
    for rep in range(2):
        p.stream_begin(U)
        lat = []
        for j in range(nchunks):
            xc = np.ascontiguousarray(x16[:, :, j * nc:(j + 1) * nc])
            t0 = time.perf_counter()
            p.stream_submit_i16(xc); p.synchronize()
            lat.append(time.perf_counter() - t0)
    lat = np.array(lat[3:])
    audio_s = nc / 16000.0
    out["chunks of %d blocks (%.0f ms of audio)" % (blocks, 1e3 * audio_s)] = {
        "median_ms_per_chunk": 1e3 * float(np.median(lat)), "p95_ms_per_chunk": 1e3 * float(np.percentile(lat, 95)),
        "frames_per_s": U * blocks / float(np.median(lat)), "real_time_factor_all_utterances": U * audio_s / float(np.median(lat)), "kernel_ms_last_chunk": p.last_timing()}
    p.close()
print(json.dumps(out, indent=1))
