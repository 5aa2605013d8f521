#!/usr/bin/env python
"""Golden vectors from the reference's OWN pure-Python adaptive beamformers (lib/pybeamformer.py, run through
oracle/pyref.py's in-memory Python-2 -> 3 shim) on the analysis-bank output of the reference's C++ (oracle/_ref).

  golden_gscrls_c8_m512.npz   SubbandGSCRLSBeamformer (pybeamformer.py:765-928), confs/gscrls.json defaults, min_frames 10
  golden_gscrls_c4_m256.npz   the same with the quadratic constraint and the norm reset firing (alpha2 0.05, max_wa_l2norm 0.2)
  golden_pyref_gsclms_c8_m512.npz   SubbandGSCLMSBeamformer (pybeamformer.py:588-762) on the g2 input of make_golden.py:
                              pins the C++ re-statement in oracle/ref_harness.cc that produced golden_gsclms_c8_m512.npz

Usage: python tests/golden/make_golden_pyref.py
"""
import os, sys
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import ref, pyref  # noqa: E402
from distant_speech_recognition_b200 import synthetic  # noqa: E402
from make_golden import proto, save  # noqa: E402

FS = 16000.0


def snapshots(x, h, M):
    return np.stack([ref.analysis(x[c], h, M, 4, 1) for c in range(x.shape[0])], axis=1)


def main():
    M = 512; K = M // 2 + 1; h, g = proto(M)
    x, d, _, _ = synthetic.make_utterance(1, 8, 12000, target_start_s=0.25)
    X = snapshots(x, h, M)
    Y, waH, nu = pyref.run_adaptive("lms", X, FS, d, M // 2, min_frames=10)
    save("pyref_gsclms_c8_m512", Y=Y[:, :K], waH=waH, n_updates=nu)
    Y, waH, nu = pyref.run_adaptive("rls", X, FS, d, M // 2, min_frames=10)
    time = ref.synthesis(Y, g, M, 4, 1)
    save("gscrls_c8_m512", x=x, delays=d, Y=Y[:, :K], waH=waH, n_updates=nu, time=time, min_frames=10)

    M = 256; K = M // 2 + 1; h, g = proto(M)
    x, d, _, _ = synthetic.make_utterance(4, 4, 8000, target_start_s=0.1)
    X = snapshots(x, h, M)
    kw = dict(min_frames=3, alpha2=0.05, max_wa_l2norm=0.2, gamma=0.5, mu=0.95, regularization_param=0.02)
    Y, waH, nu = pyref.run_adaptive("rls", X, FS, d, M // 2, **kw)
    time = ref.synthesis(Y, g, M, 4, 1)
    save("gscrls_c4_m256", x=x, delays=d, Y=Y[:, :K], waH=waH, n_updates=nu, time=time, **kw)


if __name__ == "__main__":
    main()
