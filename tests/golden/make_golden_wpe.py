#!/usr/bin/env python
"""Golden vectors of the multi-channel WPE dereverberator from the REFERENCE's own C++ (dereverberation/dereverberation.cc,
compiled unmodified into oracle/_ref/libbtkref.so), wired as unit_test/test_subband_dereverberator.py:114-170: analysis-bank
snapshots in, MultiChannelWPEDereverberation::estimate_filter, MultiChannelWPEDereverberationFeature::next per channel out.
Stored: the input samples and X'[T][C][0..M/2].

  golden_wpe_c4_m256.npz   4 mics, M = 256: (a) lower 0 / upper 5 / 2 iterations / load -18 dB / bias 1e-4 (confs/wpe.json with a
                           shorter filter), (b) lower 2 / upper 8 / 3 iterations / -20 dB / band_width 3000 Hz, estimation on
                           estimate_filter(2, 42), (c) lower 1 / upper 5 / 2 iterations / load -40 dB (light loading: large filters)
  golden_wpe_c8_m512.npz   8 mics, M = 512, lower 1 / upper 8 (L = 64), 2 iterations, load -35 dB
  golden_wpe_single_m256.npz  SingleChannelWPEDereverberationFeature, M = 256: (a) lags 0..16, -20 dB; (b) lags 2..12, 3 iterations,
                           -25 dB, band_width 3000 Hz, estimate_filter(2, 42); plus the resynthesised signal of (a)

Usage: python tests/golden/make_golden_wpe.py
"""
import os, sys
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import ref  # noqa: E402
from distant_speech_recognition_b200 import synthetic  # noqa: E402
from make_golden import proto, save  # noqa: E402


def reverberant(u, C, n, seed):
    """make_utterance + a decaying multi-tap tail per channel, so that the linear prediction has something to remove."""
    x, d, mpos, _ = synthetic.make_utterance(u, C, n, target_start_s=0.05)
    rng = np.random.default_rng(seed)
    y = x.astype(np.float64).copy()
    for c in range(C):
        for tap in range(1, 7):
            dl = 300 * tap + int(rng.integers(0, 120))
            y[c, dl:] += (0.55 ** tap) * rng.choice([-1.0, 1.0]) * x[c, :-dl]
    return y.astype(np.float32)


def main():
    M = 256; K = M // 2 + 1; h, _ = proto(M)
    x = reverberant(6, 4, 8000, 1)
    X = np.stack([ref.analysis(x[c], h, M, 4, 1) for c in range(4)], axis=1)
    ka = dict(lower_num=0, upper_num=5, iterations_num=2, load_db=-18.0, band_width=0.0, diagonal_bias=1e-4)
    kb = dict(lower_num=2, upper_num=8, iterations_num=3, load_db=-20.0, band_width=3000.0, diagonal_bias=1e-3, start_frame_no=2, end_frame_no=42)
    kc = dict(lower_num=1, upper_num=5, iterations_num=2, load_db=-40.0, band_width=0.0, diagonal_bias=1e-4)
    Xa, ua = ref.wpe(X, **ka)
    Xb, ub = ref.wpe(X, **kb)
    Xc, uc = ref.wpe(X, **kc)
    save("wpe_c4_m256", x=x, Xa=Xa[:, :, :K], Xb=Xb[:, :, :K], Xc=Xc[:, :, :K], used_a=ua, used_b=ub)

    M = 512; K = M // 2 + 1; h, _ = proto(M)
    x = reverberant(7, 8, 12000, 2)
    X = np.stack([ref.analysis(x[c], h, M, 4, 1) for c in range(8)], axis=1)
    Xa, ua = ref.wpe(X, lower_num=1, upper_num=8, iterations_num=2, load_db=-35.0, band_width=0.0, diagonal_bias=1e-4)
    save("wpe_c8_m512", x=x, Xa=Xa[:, :, :K], used_a=ua)

    # SingleChannelWPEDereverberationFeature (dereverberation.cc:24-310), wired as test_subband_dereverberator.py:53-92
    M = 256; K = M // 2 + 1; h, g = proto(M)
    x = reverberant(9, 1, 8000, 3)
    X = ref.analysis(x[0], h, M, 4, 1)
    Xa, ua = ref.wpe_single(X, lower_num=0, upper_num=16, iterations_num=2, load_db=-20.0, band_width=0.0)
    Xb, ub = ref.wpe_single(X, lower_num=2, upper_num=12, iterations_num=3, load_db=-25.0, band_width=3000.0, start_frame_no=2, end_frame_no=42)
    ta = ref.synthesis(Xa, g, M, 4, 1)
    save("wpe_single_m256", x=x, Xa=Xa[:, :K], Xb=Xb[:, :K], used_a=ua, used_b=ub, time_a=ta)


def kinect():
    """unit_test/test_subband_dereverberator.py on its default inputs (the whole 4-channel Kinect recording, shipped M = 256 prototypes)
    with confs/wpe.json read where it lies, multi-channel and single-channel (channel 1) variants, estimate_filter() over the whole
    utterance.  The samples are the x16 array of golden_online_kinect_c4_m256 and are not stored again; stored: the resynthesised
    output of every channel and frames 240..299 of the dereverberated subband signals."""
    import json, pickle, wave
    base = "/root/reference/btk20_src/unit_test/"
    d = base + "data/CMU/R1/M1005/KINECT/RAW/segmented/"
    M, K = 256, 129
    xs = []
    for c in range(1, 5):
        w = wave.open(d + "U1001_1M_16k_b16_c%d.wav" % c); xs.append(np.frombuffer(w.readframes(w.getnframes()), np.int16)); w.close()
    x = np.stack(xs).astype(np.float32)
    h = np.asarray(pickle.load(open(base + "prototype.ny/h-M256-m4-r1.pickle", "rb"), encoding="latin1"), np.float64)
    g = np.asarray(pickle.load(open(base + "prototype.ny/g-M256-m4-r1.pickle", "rb"), encoding="latin1"), np.float64)
    conf = json.load(open(base + "confs/wpe.json"))
    X = np.stack([ref.analysis(x[c], h, M, 4, 1) for c in range(4)], axis=1)
    # multi_channel_wpe (:114-170): defaults of the script where the file is silent
    km = dict(lower_num=conf.get("lower_num", 0), upper_num=conf.get("upper_num", 32), iterations_num=conf.get("iterations_num", 2),
              load_db=conf.get("load_db", -20.0), band_width=conf.get("band_width", 0.0), diagonal_bias=conf.get("diagonal_bias", 0.001), samplerate=16000.0)
    Xm, um = ref.wpe(X, **km)
    tm = np.stack([ref.synthesis(Xm[:, c, :], g, M, 4, 1) for c in range(4)])
    # single_channel_wpe (:53-92) on the first file
    ks = dict(lower_num=conf.get("lower_num", 0), upper_num=conf.get("upper_num", 64), iterations_num=conf.get("iterations_num", 2),
              load_db=conf.get("load_db", -20.0), band_width=conf.get("band_width", 0.0), samplerate=16000.0)
    Xs, us = ref.wpe_single(X[:, 0, :], **ks)
    ts = ref.synthesis(Xs, g, M, 4, 1)
    red = 1.0 - np.linalg.norm(Xm) ** 2 / np.linalg.norm(X) ** 2
    print("kinect WPE: %d / %d frames used, energy removed %.1f %%" % (um, us, 100 * red))
    save("wpe_kinect_c4_m256", frames=np.array([240, 300]), used_multi=um, used_single=us, conf=json.dumps(conf),
         X_multi=Xm[240:300, :, :K].astype(np.complex64), time_multi=tm.astype(np.float32),
         X_single=Xs[240:300, :K].astype(np.complex64), time_single=ts.astype(np.float32))


if __name__ == "__main__":
    main()
    kinect()
