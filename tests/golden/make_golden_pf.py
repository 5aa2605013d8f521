#!/usr/bin/env python
"""Golden vectors of the McCowan and Lefkimmiatis post-filters from the REFERENCE's own C++ (oracle/_ref/libbtkref.so),
wired as unit_test/test_online_beamforming.py:137-151 does and with the parameters of unit_test/confs/sd_and_mccowan.json
and sd_and_lefkimmiatis.json (plus one type-1 / warm-up variant each).  Subband outputs: bins 0..M/2 only.

Usage: python tests/golden/make_golden_pf.py
"""
import os, sys
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import ref  # noqa: E402
from distant_speech_recognition_b200 import synthetic  # noqa: E402
from make_golden import proto, save  # noqa: E402


def main():
    m, r = 4, 1
    # super-directive MVDR (mu 0.01) + McCowan, 4 mics, M = 256
    M = 256; K = M // 2 + 1; h, g = proto(M)
    x, d, mpos, _ = synthetic.make_utterance(2, 4, 8000, target_start_s=0.1)
    a = ref.beamform(x, h, g, d, M, m, r, bf_kind=ref.BF_MVDR_SD, mpos=mpos, mvdr_mu=0.01,
                     pf=dict(kind="mccowan", alpha=0.7, type=2, diag_load=0.01))
    b = ref.beamform(x, h, g, d, M, m, r, bf_kind=ref.BF_MVDR_SD, mpos=mpos, mvdr_mu=0.01,
                     pf=dict(kind="mccowan", alpha=0.6, type=1, min_frames=3, diag_load=0.0, threshold=0.9))
    save("mccowan_c4_m256", x=x, delays=d, mpos=mpos, mu=0.01, w=a["w"], Ya=a["Y"][:, :K], timea=a["time"], Yb=b["Y"][:, :K], timeb=b["time"],
         upper_a=a["Y"][:2, K:], upper_b=b["Y"][:5, K:])
    # delay-and-sum + Lefkimmiatis, 8 mics, M = 512
    M = 512; K = M // 2 + 1; h, g = proto(M)
    x, d, mpos, _ = synthetic.make_utterance(3, 8, 8000, target_start_s=0.1)
    a = ref.beamform(x, h, g, d, M, m, r, bf_kind=ref.BF_DS, mpos=mpos,
                     pf=dict(kind="lefkimmiatis", alpha=0.8, type=2, min_sv=1e-4, diag_load=0.1, fbin1=100))
    b = ref.beamform(x, h, g, d, M, m, r, bf_kind=ref.BF_DS, mpos=mpos,
                     pf=dict(kind="lefkimmiatis", alpha=0.6, type=1, min_frames=2, min_sv=1e-8, diag_load=0.01, fbin1=0))
    save("lefkimmiatis_c8_m512", x=x, delays=d, mpos=mpos, Ya=a["Y"][:, :K], timea=a["time"], Yb=b["Y"][:, :K], timeb=b["time"])


if __name__ == "__main__":
    main()
