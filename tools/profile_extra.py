#!/usr/bin/env python
"""Timing / ncu driver for the kernels outside the bench.py headline: SOS statistics + solves at configs[1] size, WPE at a small
size.  python tools/profile_extra.py sos|wpe   (JSON on stdout; run under ncu for captures — never a bench value then)."""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from distant_speech_recognition_b200 import _capi, synthetic
from bench_configs import proto, tiled_batch, timed


def sos():
    C, M, U, n = 8, 512, 256, 80000
    h, g = proto(M); x, d = tiled_batch(U, C, n, 8)
    p = _capi.Pipeline(C, M, 4, 1, beamformer=_capi.BF_DS, max_utterances=U, max_samples=n)
    p.set_prototypes(h, g); p.submit(x); p.run_analysis(); p.synchronize()
    T = p.num_frames; K = M // 2 + 1
    labels = np.tile(np.array([[[1.0, 3.0]]]), (U, 1, 1))
    def acc(): p.sos_reset_stats(); p.sos_accumulate_from_label(labels, 10.0)
    s_acc = timed(acc, steps=3, warm=1)
    mt = (np.random.default_rng(0).uniform(size=(U, T, K)) > 0.5).astype(np.float32); mj = 1.0 - mt
    def accm(): p.sos_reset_stats(); p.sos_accumulate_from_tfmask(mt, mj, 10.0)
    s_accm = timed(accm, steps=2, warm=1)
    s_b = timed(lambda: (p.sos_calc_weights(_capi.SOS_BMVDR), p.synchronize()), steps=3, warm=1)
    s_g = timed(lambda: (p.sos_calc_weights(_capi.SOS_GEV), p.synchronize()), steps=3, warm=1)
    s_app = timed(lambda: (p.run_beamformer(True), p.synchronize()), steps=3, warm=1)
    xbytes = C * K * 8 * U * T
    print(json.dumps({"sos configs[1] size (8 mics, M=512, 256 x 5 s)": dict(
        frames=U * T, accumulate_label_ms=1e3 * s_acc, accumulate_label_hbm_frac_two_reads_of_X=2 * xbytes / s_acc / 1e9 / 6566.7,
        accumulate_tfmask_ms_incl_mask_upload=1e3 * s_accm, bmvdr_solve_ms=1e3 * s_b, gev_solve_ms=1e3 * s_g, apply_synthesis_ms=1e3 * s_app)}, indent=1))


def wpe():
    C, M, U, n = 8, 512, 2, 40000
    h, g = proto(M); x, d = tiled_batch(U, C, n, 2)
    wpe = dict(lower_num=0, upper_num=32, iterations_num=2, load_db=-18.0, band_width=0.0, diagonal_bias=1e-4)
    p = _capi.Pipeline(C, M, 4, 1, beamformer=_capi.BF_GSC_LMS, max_utterances=U, max_samples=n, wpe=wpe)
    p.set_prototypes(h, g); p.set_delays(d); p.submit(x)
    p.run(True); p.synchronize()
    print(json.dumps({"wpe small": dict(frames=U * p.num_frames, wpe_ms=p.last_timing_wpe(), kernels=p.last_timing())}))


if __name__ == "__main__":
    {"sos": sos, "wpe": wpe}[sys.argv[1]]()
