#!/usr/bin/env python
"""Golden vectors of the reference's C++ SubbandGSCRLS (beamformer/beamformer.cc:1447-1699) from the compiled reference
(oracle/_ref, bf_kind 5 of oracle/ref_harness.cc): SubbandGSCRLS(M, False, myu, sigma2); calc_gsc_weights; init_precision_matrix;
set_quadratic_constraint.  No CUDA kernel consumes them: they pin oracle/restate.py::gsc_rls_cpp.

Usage: python tests/golden/make_golden_rls_cpp.py
"""
import os, sys
import numpy as np
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import ref  # noqa: E402
from distant_speech_recognition_b200 import synthetic  # noqa: E402
from make_golden import proto, save  # noqa: E402

CASES = (dict(mu=0.9, sigma2=0.01, init_sigma2=0.01),                                   # the constructor / init_precision_matrix defaults
         dict(mu=0.97, sigma2=0.0, init_sigma2=1e6, alpha=0.5, qctype=2),                 # THRESHOLD_LIMITATION
         dict(mu=0.95, sigma2=1e-3, init_sigma2=1.0, alpha=0.3, qctype=1))                # CONSTANT_NORM


def main():
    M = 256; K = M // 2 + 1; h, g = proto(M)
    x, d, _, _ = synthetic.make_utterance(21, 4, 8000, target_start_s=0.1)
    out = dict(x=x, delays=d)
    for i, kw in enumerate(CASES):
        res = ref.beamform(x, h, g, d, M, bf_kind=ref.BF_GSC_RLS_CPP, rls=kw, do_synthesis=True)
        out["Y%d" % i] = res["Y"][:, :K]
    save("gscrls_cpp_c4_m256", **out)


if __name__ == "__main__":
    main()
