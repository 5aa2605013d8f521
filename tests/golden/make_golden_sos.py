#!/usr/bin/env python
"""Golden vectors for the SOS batch beamformers from the reference's OWN Python (lib/pybeamformer.py:1026-1357, run through
oracle/pyref.py) on the analysis-bank output of the reference's C++ (oracle/_ref), in the call order of
unit_test/test_sos_batch_beamforming.py:186-233.

  golden_bmvdr_vad_c8_m512.npz     SubbandBlindMVDRBeamformer, VAD label (confs/bmvdr_vad.json shape: one segment), two segments here
  golden_bmvdr_tfmask_c4_m256.npz  SubbandBlindMVDRBeamformer, TF masks with FRACTIONAL values (the integer-count truncation quirk)
  golden_gev_vad_c8_m512.npz       SubbandGEVBeamformer, VAD label (confs/gev_vad.json)
  golden_gev_tfmask_c4_m256.npz    SubbandGEVBeamformer, binary TF masks (confs/gev_tfmask.json)

Usage: python tests/golden/make_golden_sos.py
"""
import os, sys
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import ref, pyref, restate  # noqa: E402
from distant_speech_recognition_b200 import synthetic  # noqa: E402
from make_golden import proto, save  # noqa: E402

FS = 16000.0


def snapshots(x, h, M):
    return np.stack([ref.analysis(x[c], h, M, 4, 1) for c in range(x.shape[0])], axis=1)


def tf_masks(X, K, seed, fractional):
    """Deterministic TF masks from the snapshot energies: target where channel-0 magnitude is above the bin median."""
    rng = np.random.default_rng(seed)
    mag = np.abs(X[:, 0, :K])
    med = np.median(mag, axis=0, keepdims=True)
    mt = (mag > med).astype(np.float64)
    mj = 1.0 - mt
    if fractional:
        mt = mt * rng.uniform(0.3, 1.7, mt.shape)
        mj = mj * rng.uniform(0.3, 1.7, mj.shape)
        mt[rng.uniform(size=mt.shape) < 0.1] = 0.0
    # the C-ABI takes float32 masks: round first so the reference sees exactly the stored values
    return mt.astype(np.float32).astype(np.float64), mj.astype(np.float32).astype(np.float64)


def one(name, kind, C, M, n, useed, labels=None, mask=None, **kw):
    K = M // 2 + 1; h, g = proto(M)
    x, d, _, _ = synthetic.make_utterance(useed, C, n, target_start_s=0.3)
    X = snapshots(x, h, M)
    mt = mj = None
    if mask is not None:
        mt, mj = tf_masks(X, K, useed, fractional=(mask == "fractional"))
    res = pyref.run_sos(kind, X, FS, M // 2, labels=labels, mask_t=mt, mask_j=mj, **kw)
    w = np.conj(res["wqH"])
    time = ref.synthesis(res["Y"], g, M, 4, 1)
    # cross-check the fp64 restatement against the reference's own output before saving
    Rt, Rn, ct, cn = restate.sos_accumulate(X, FS, M // 2, target_labs=labels, mask_t=mt, mask_j=mj, energy_threshold=kw.get("energy_threshold", 10))
    assert np.array_equal(ct, res["ct"]) and np.array_equal(cn, res["cn"]), (ct[:5], res["ct"][:5])
    if kind == "bmvdr":
        w2 = restate.sos_bmvdr_weights(Rt, Rn, ct, cn, gamma=kw.get("gamma", 1e-6), ref_micx=kw.get("ref_micx", 0), offset=kw.get("offset", 0.0))
        sgn = 1.0
    else:
        w2 = restate.sos_gev_weights(Rt, Rn, cn, gamma=kw.get("gamma", 1e-6))
        sgn = np.sign(np.real(np.vdot(w2[0], w[0])))
    err = np.linalg.norm(sgn * w2 - w) / np.linalg.norm(w)
    Y2 = restate.sos_apply(X, sgn * w2)
    errY = np.linalg.norm(Y2 - res["Y"]) / np.linalg.norm(res["Y"])
    print("%s: restatement vs reference python: weights %.2e, Y %.2e (global sign %+d), counts t %d..%d n %d..%d" %
          (name, err, errY, sgn, ct.min(), ct.max(), cn.min(), cn.max()))
    assert err < 1e-8 and errY < 1e-8
    extra = {}
    if labels is not None:
        extra["labels"] = np.asarray(labels, np.float64)
    if mt is not None:
        extra["mask_t"] = mt.astype(np.float32); extra["mask_j"] = mj.astype(np.float32)
    save(name, x=x, Y=res["Y"][:, :K].astype(np.complex64), w=w, time=time.astype(np.float32), ct=res["ct"], cn=res["cn"],
         Rn=res["Rn"].astype(np.complex64), gamma=kw.get("gamma", 1e-6), ref_micx=kw.get("ref_micx", 0), offset=kw.get("offset", 0.0),
         energy_threshold=kw.get("energy_threshold", 10), **extra)


def kinect():
    """The reference's OWN fixtures for this path (unit_test/confs/{bmvdr,gev}_tfmask.json): the 4-channel Kinect recording
    unit_test/data/CMU/R1/M1005/KINECT/RAW/segmented/U1001_1M_16k_b16_c{1..4}.wav, its speech / noise TF-mask pickles (622 x 129)
    and the shipped M = 256 prototype pickles; first 240 frames.  Stored: the int16 samples, the masks, and the outputs of the
    reference's SubbandBlindMVDRBeamformer / SubbandGEVBeamformer."""
    import pickle, wave
    base = "/root/reference/btk20_src/unit_test/"
    d = base + "data/CMU/R1/M1005/KINECT/RAW/segmented/"
    M, K, NF = 256, 129, 240
    xs = []
    for c in range(1, 5):
        w = wave.open(d + "U1001_1M_16k_b16_c%d.wav" % c); xs.append(np.frombuffer(w.readframes(w.getnframes()), np.int16)); w.close()
    x16 = np.stack(xs)[:, : (NF - 4) * (M // 2)]

    def load_mask(path):
        rows = []
        with open(path, "rb") as f:
            while True:
                try:
                    rows.append(pickle.load(f, encoding="latin1"))
                except EOFError:
                    break
        return np.array(rows)
    mt = load_mask(d + "U1001_1M_16k.speech.tfmask.pickle")[:NF].astype(np.uint8)
    mj = load_mask(d + "U1001_1M_16k.noise.tfmask.pickle")[:NF].astype(np.uint8)
    h = np.asarray(pickle.load(open(base + "prototype.ny/h-M256-m4-r1.pickle", "rb"), encoding="latin1"), np.float64)
    g = np.asarray(pickle.load(open(base + "prototype.ny/g-M256-m4-r1.pickle", "rb"), encoding="latin1"), np.float64)
    x = x16.astype(np.float32)
    X = snapshots(x, h, M)
    assert X.shape[0] == NF
    out = dict(x16=x16, mask_t=mt, mask_j=mj)
    for kind in ("bmvdr", "gev"):
        res = pyref.run_sos(kind, X, FS, M // 2, mask_t=mt.astype(np.float64), mask_j=mj.astype(np.float64), energy_threshold=10, gamma=1e-6)
        out["w_" + kind] = np.conj(res["wqH"]); out["Y_" + kind] = res["Y"][:, :K].astype(np.complex64)
        out["time_" + kind] = ref.synthesis(res["Y"], g, M, 4, 1).astype(np.float32)
        out["ct"] = res["ct"]; out["cn"] = res["cn"]
    save("sos_kinect_c4_m256", **out)


def kinect_vad():
    """unit_test/test_sos_batch_beamforming.py on its default inputs (the whole Kinect recording) with confs/{smimvdr, bmvdr_vad,
    gev_vad}.json read where they lie: VAD label [[1.5, 4.0]], energy_threshold 10, SMI-MVDR with mu 1e-4 towards the look direction
    of the file.  The samples are the x16 array of golden_online_kinect_c4_m256 (same wav files) and are not stored again; stored per
    configuration: weights, frame counts, the resynthesised signal and frames 200..359 (speech) of the subband output."""
    import json, pickle, wave
    base = "/root/reference/btk20_src/unit_test/"
    d = base + "data/CMU/R1/M1005/KINECT/RAW/segmented/"
    M, K, D = 256, 129, 128
    xs = []
    for c in range(1, 5):
        w = wave.open(d + "U1001_1M_16k_b16_c%d.wav" % c); xs.append(np.frombuffer(w.readframes(w.getnframes()), np.int16)); w.close()
    x = np.stack(xs).astype(np.float32)
    h = np.asarray(pickle.load(open(base + "prototype.ny/h-M256-m4-r1.pickle", "rb"), encoding="latin1"), np.float64)
    g = np.asarray(pickle.load(open(base + "prototype.ny/g-M256-m4-r1.pickle", "rb"), encoding="latin1"), np.float64)
    X = snapshots(x, h, M)
    out = dict(frames=np.array([200, 360]))
    for name in ("bmvdr_vad", "gev_vad"):
        c = json.load(open(base + "confs/%s.json" % name)); bf = c["beamformer"]
        labels = [tuple(l) for l in c["target"]["vad_label"]]
        res = pyref.run_sos(bf["type"], X, FS, D, labels=labels, energy_threshold=bf.get("energy_threshold", 10), gamma=bf.get("gamma", 1e-6),
                            ref_micx=bf.get("ref_micx", 0), offset=bf.get("offset", 0.0))
        out["labels"] = np.asarray(labels, np.float64)
        out["w_" + name] = np.conj(res["wqH"]); out["ct"] = res["ct"]; out["cn"] = res["cn"]
        out["Y_" + name] = res["Y"][200:360, :K].astype(np.complex64)
        out["time_" + name] = ref.synthesis(res["Y"], g, M, 4, 1).astype(np.float32)
        print(name, res["Y"].shape, "counts", res["ct"][:3], res["cn"][:3])
    c = json.load(open(base + "confs/smimvdr.json")); bf = c["beamformer"]
    mod = pyref.load()
    delays = np.asarray(mod.calc_delays(c["array_type"], c["microphone_positions"], c["target"]["positions"][0][1], sspeed=343740.0), np.float64)
    lab = c["target"]["vad_label"]
    assert len(lab) == 1
    res = ref.beamform(x, h, g, delays, M, 4, 1, samplerate=FS, bf_kind=ref.BF_SMI_MVDR, mvdr_mu=bf.get("mu", 1e-4), smi_label=tuple(lab[0]),
                       smi_energy_threshold=bf.get("energy_threshold", 10))
    out["delays"] = delays; out["mu_smimvdr"] = bf.get("mu", 1e-4)
    out["cov_smimvdr"] = res["cov"].astype(np.complex64); out["w_smimvdr"] = res["w"]
    out["Y_smimvdr"] = res["Y"][200:360, :K].astype(np.complex64); out["time_smimvdr"] = res["time"].astype(np.float32)
    print("smimvdr", res["Y"].shape, "cond(R) median %.1e max %.1e" % (np.median([np.linalg.cond(r) for r in res["cov"]]), max(np.linalg.cond(r) for r in res["cov"])))
    save("sos_kinect_vad_c4_m256", **out)


def main():
    kinect()
    kinect_vad()
    one("bmvdr_vad_c8_m512", "bmvdr", 8, 512, 16000, 11, labels=[(0.3, 0.55), (0.7, -1)], ref_micx=2, offset=0.01)
    one("bmvdr_tfmask_c4_m256", "bmvdr", 4, 256, 8000, 12, mask="fractional", gamma=1e-4)
    one("gev_vad_c8_m512", "gev", 8, 512, 16000, 13, labels=[(0.3, 0.8)])
    one("gev_tfmask_c4_m256", "gev", 4, 256, 8000, 14, mask="binary")


if __name__ == "__main__":
    main()
