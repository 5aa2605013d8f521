#!/usr/bin/env python
"""GPU debug aid: per-bin comparison of the SOS statistics / GEV weights with the fp64 restatement."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from distant_speech_recognition_b200 import _capi
from oracle import restate
from conftest import load_golden, rel_l2

for name, C, M in (("gev_tfmask_c4_m256", 4, 256), ("gev_vad_c8_m512", 8, 512)):
    g = load_golden(name); pr = np.load(os.path.join(ROOT, "tests", "golden", "prototype_M%d_m4_r1.npz" % M)); K = M // 2 + 1
    x = g["x"]
    p = _capi.Pipeline(C, M, 4, 1, beamformer=_capi.BF_DS, max_utterances=1, max_samples=x.shape[1]); p.set_prototypes(pr["h"], pr["g"])
    p.submit(x[None]); p.run_analysis()
    if "mask_t" in g.files:
        p.sos_accumulate_from_tfmask(g["mask_t"], g["mask_j"], 10.0)
    else:
        p.sos_accumulate_from_label(g["labels"], 10.0)
    Rt, Rn, cnt = p.sos_get_stats()
    X = np.stack([restate.analysis(x[c], pr["h"], M, 4, 1) for c in range(C)], axis=1)
    labels = [tuple(r) for r in g["labels"]] if "labels" in g.files else None
    Rto, Rno, ct, cn = restate.sos_accumulate(X, 16000.0, M // 2, target_labs=labels, mask_t=g["mask_t"] if "mask_t" in g.files else None,
                                              mask_j=g["mask_j"] if "mask_j" in g.files else None, energy_threshold=10.0)
    print(name, "counts equal", np.array_equal(cnt[0, :, 0], ct), np.array_equal(cnt[0, :, 1], cn))
    eT = np.array([rel_l2(Rt[0, k], Rto[k]) for k in range(K)]); eN = np.array([rel_l2(Rn[0, k], Rno[k]) for k in range(K)])
    print(" Rt err max %.2e at %d, Rn err max %.2e at %d" % (eT.max(), eT.argmax(), eN.max(), eN.argmax()))
    p.sos_calc_weights(_capi.SOS_GEV, gamma=float(g["gamma"]))
    w = p.get_weights()[0].astype(np.complex128)
    wo = restate.sos_gev_weights(Rto, Rno, cn, gamma=float(g["gamma"]))
    wg = restate.sos_gev_weights(Rt[0], Rn[0], cnt[0, :, 1], gamma=float(g["gamma"]))   # restatement on the GPU's own statistics
    print(" w vs restatement %.2e ; w vs restatement-on-GPU-stats %.2e" % (rel_l2(w, wo), rel_l2(w, wg)))
    ph = np.array([np.angle(np.vdot(wo[k], w[k])) for k in range(K)]); mg = np.array([np.linalg.norm(w[k]) / np.linalg.norm(wo[k]) for k in range(K)])
    dirn = np.array([abs(np.vdot(wo[k], w[k])) / np.linalg.norm(w[k]) / np.linalg.norm(wo[k]) for k in range(K)])
    np.set_printoptions(precision=3, suppress=True, linewidth=200)
    print(" phase offset per bin (deg):", np.degrees(ph)[:40])
    print(" |w|/|wo| min max:", mg.min(), mg.max(), " direction cos min:", dirn.min(), dirn.argmin())
    p.close()
