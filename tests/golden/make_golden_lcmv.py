#!/usr/bin/env python
"""Golden LCMV quiescent weights from the reference's own calcMainlobeN / calc_null_beamformer_ (beamformer.cc:299-363,573-721)
via oracle/_ref (build container only).  Output: tests/golden/golden_lcmv.npz."""
import os, sys
import numpy as np
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import ref
from distant_speech_recognition_b200 import synthetic

M, C, FS = 512, 8, 16000.0
_, mpos = synthetic.array_for_channels(C)
dT = synthetic.far_field_delays("linear", mpos, np.pi / 3)
dJ1 = synthetic.far_field_delays("linear", mpos, 2 * np.pi / 3)
dJ2 = synthetic.far_field_delays("linear", mpos, 0.3)
w2, B2 = ref.lcmv_weights(M, C, 2, FS, dT, dJ1[None])
w3, B3 = ref.lcmv_weights(M, C, 3, FS, dT, np.stack([dJ1, dJ2]))
np.savez_compressed(os.path.join(HERE, "golden_lcmv.npz"), dT=dT, dJ1=dJ1, dJ2=dJ2, w2=w2, B2=B2, w3=w3, B3=B3)
print("ok", w2.shape, B2.shape, w3.shape, B3.shape)
