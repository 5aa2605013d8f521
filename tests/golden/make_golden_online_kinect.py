#!/usr/bin/env python
"""Golden vectors of the reference's online beamforming test ON ITS OWN FIXTURES AND CONFIGURATIONS:
unit_test/test_online_beamforming.py with its default inputs — the 4-channel Kinect recording
unit_test/data/CMU/R1/M1005/KINECT/RAW/segmented/U1001_1M_16k_b16_c{1..4}.wav, the shipped prototypes
unit_test/prototype.ny/{h,g}-M256-m4-r1.pickle (M = 256, m = 4, r = 1, delay-compensation type 2) — and the parameter files
unit_test/confs/{ds, ds_and_zelinski, sd, sd_and_zelinski, sd_and_mccowan, sd_and_lefkimmiatis, gsclms, gscrls}.json, read where
they lie (microphone positions, look direction, every hyper-parameter); of confs/lcmv_and_zelinski.json the LCMV weights.

Who computes what (same split as the other goldens): the filter banks, D&S / super-directive weights and the three post-filters
are the reference's C++ compiled unmodified (oracle/_ref via oracle/ref.py, wired like test_online_beamforming.py:51-228); the two
adaptive beamformers are the reference's own Python loops (lib/pybeamformer.py through oracle/pyref.py); the delays come from the
reference's own calc_delays (pybeamformer.py:41-153).

Stored: the int16 samples; per configuration the subband output (bins 0..M/2, complex64), the resynthesised signal (float32) and
the script's own report value total_energy = sum(buf . buf).  The static configurations run on a 160-frame excerpt (19 968 samples
from sample 23 040 on, where the talker starts; the first 1.5 s are room noise) to keep the file small; the adaptive ones
(adaptation starts at frame 128) on the whole 5 s recording.

Usage: python tests/golden/make_golden_online_kinect.py     (needs /root/reference and oracle/_ref)
"""
import json
import os
import pickle
import sys
import wave

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import ref, pyref  # noqa: E402
from make_golden import save  # noqa: E402

BASE = "/root/reference/btk20_src/unit_test/"
FS, M, m, r = 16000, 256, 4, 1
K, D = M // 2 + 1, M >> r
NF_STATIC, S0_STATIC = 160, 180 * 128
SSPEED = 343740.0


def load_fixtures():
    d = BASE + "data/CMU/R1/M1005/KINECT/RAW/segmented/"
    xs = []
    for c in range(1, 5):
        w = wave.open(d + "U1001_1M_16k_b16_c%d.wav" % c)
        assert w.getframerate() == FS and w.getsampwidth() == 2 and w.getnchannels() == 1
        xs.append(np.frombuffer(w.readframes(w.getnframes()), np.int16)); w.close()
    x16 = np.stack(xs)
    h = np.asarray(pickle.load(open(BASE + "prototype.ny/h-M256-m4-r1.pickle", "rb"), encoding="latin1"), np.float64)
    g = np.asarray(pickle.load(open(BASE + "prototype.ny/g-M256-m4-r1.pickle", "rb"), encoding="latin1"), np.float64)
    return x16, h, g


def conf(name):
    return json.load(open(BASE + "confs/%s.json" % name))


def main():
    x16, h, g = load_fixtures()
    xfull = x16.astype(np.float32)
    xs = xfull[:, S0_STATIC: S0_STATIC + (NF_STATIC - m * (1 << r) // 2) * D]
    mod = pyref.load()
    out = dict(x16=x16, s0_static=S0_STATIC, n_static=xs.shape[1])
    delays = None
    for name in ("ds", "ds_and_zelinski", "sd", "sd_and_zelinski", "sd_and_mccowan", "sd_and_lefkimmiatis", "gsclms", "gscrls"):
        c = conf(name)
        mpos = np.asarray(c["microphone_positions"], np.float64)
        d = np.asarray(mod.calc_delays(c["array_type"], c["microphone_positions"], c["target"]["positions"][0][1], sspeed=SSPEED), np.float64)
        if delays is None:
            delays = d; out["delays"] = d; out["mpos"] = mpos
        assert np.array_equal(d, delays) and np.array_equal(mpos, out["mpos"])      # every file describes the same array and look direction
        bf, pf = c["beamformer"], c.get("postfilter")
        out["conf_" + name] = json.dumps(c)      # the parameter file itself, so that tests can configure the front end from it
        if bf["type"] in ("delay_and_sum", "super_directive"):
            pfd = None
            if pf is not None:   # test_online_beamforming.py:132-156
                if pf["type"] == "zelinski":
                    pfd = dict(kind="zelinski", alpha=pf.get("alpha", 0.6), type=pf.get("subtype", 2))
                elif pf["type"] == "mccowan":
                    pfd = dict(kind="mccowan", alpha=pf.get("alpha", 0.6), type=pf.get("subtype", 2), diag_load=bf.get("diagonal_load", 0.01))
                else:
                    pfd = dict(kind="lefkimmiatis", min_sv=pf.get("min_sv", 1e-8), fbin1=pf.get("fbin_no1", 128), alpha=pf.get("alpha", 0.8),
                               type=pf.get("subtype", 2), diag_load=bf.get("diagonal_load", 0.1))
            if bf["type"] == "delay_and_sum":   # SubbandGSCBeamformer(afbs, Nc=1) with zero active weights (:98-99)
                res = ref.beamform(xs, h, g, d, M, m, r, samplerate=float(FS), bf_kind=ref.BF_GSC, mpos=mpos, pf=pfd)
            else:                               # SubbandMVDRBeamformer.calc_sd_beamformer_weights(mu = diagonal_load or 0.01) (:189-192)
                res = ref.beamform(xs, h, g, d, M, m, r, samplerate=float(FS), bf_kind=ref.BF_MVDR_SD, mpos=mpos, mvdr_mu=bf.get("diagonal_load", 0.01),
                                   sspeed=SSPEED, pf=pfd)
                out["w_" + name] = res["w"]
            Y, t = res["Y"], res["time"]
            assert Y.shape[0] == NF_STATIC
        else:
            X = np.stack([ref.analysis(xfull[ch], h, M, m, r) for ch in range(4)], axis=1)
            params = {k: v for k, v in bf.items() if k != "type"}
            Y, waH, nu = pyref.run_adaptive("lms" if bf["type"] == "gsclms" else "rls", X, float(FS), d, D, **params)
            t = ref.synthesis(Y, g, M, m, r)
            out["waH_" + name] = waH; out["n_updates_" + name] = nu
        out["Y_" + name] = Y[:, :K].astype(np.complex64)
        out["time_" + name] = t.astype(np.float32)
        out["energy_" + name] = float(np.inner(t.astype(np.float64), t.astype(np.float64)))
        print(name, Y.shape, t.shape, "total_energy/frames = %.3f" % (out["energy_" + name] / (len(t) // D)))
    # confs/lcmv_and_zelinski.json: SubbandGSCBeamformer(afbs, Nc=2).calc_beamformer_weights_n -> C++ calcMainlobeN (beamformer.cc:573-721).
    # The harness wires the LCMV weights only (ref_lcmv_weights), so the quiescent vectors and blocking matrices are the reference's and
    # the tests run the (elsewhere pinned) static-GSC + Zelinski chain on them.
    c = conf("lcmv_and_zelinski")
    assert np.array_equal(np.asarray(c["microphone_positions"], np.float64), out["mpos"])
    dT = np.asarray(mod.calc_delays(c["array_type"], c["microphone_positions"], c["target"]["positions"][0][1], sspeed=SSPEED), np.float64)
    dJ = np.stack([np.asarray(mod.calc_delays(c["array_type"], c["microphone_positions"], nz["positions"][0][1], sspeed=SSPEED), np.float64) for nz in c["noises"]])
    wl, Bl = ref.lcmv_weights(M, 4, 1 + len(dJ), float(FS), dT, dJ)
    out.update(lcmv_dT=dT, lcmv_dJ=dJ, lcmv_w=wl, lcmv_B=Bl, conf_lcmv_and_zelinski=json.dumps(c))
    print("lcmv: target %s rad, jammer %s rad" % (c["target"]["positions"][0][1][0], c["noises"][0]["positions"][0][1][0]))
    save("online_kinect_c4_m256", **out)


if __name__ == "__main__":
    main()
