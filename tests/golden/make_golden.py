#!/usr/bin/env python
"""Generate golden input/output vectors from the REFERENCE's own C++ (oracle/_ref/libbtkref.so, built by
`make -C oracle` from /root/reference in the build container) on seeded synthetic inputs.

Committed outputs: tests/golden/golden_*.npz.  They pin (a) the NumPy restatement oracle/restate.py and (b) the CUDA
path (tests/test_parity_gpu.py) without needing /root/reference at test time.  Subband outputs are stored for the
K = M/2+1 unique bins only (the reference fills the rest by conjugate symmetry).

Usage: python tests/golden/make_golden.py
"""
import os, sys
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import ref  # noqa: E402
from distant_speech_recognition_b200 import synthetic  # noqa: E402

FS = 16000.0


def proto(M):
    p = np.load(os.path.join(HERE, "prototype_M%d_m4_r1.npz" % M))
    return p["h"], p["g"]


def save(name, **kw):
    path = os.path.join(HERE, "golden_%s.npz" % name)
    np.savez_compressed(path, **kw)
    print("wrote", path, os.path.getsize(path) // 1024, "KiB")


def main():
    m, r = 4, 1
    # g1: configs[0] shape (2-mic D&S, M=256), shortened to 1 s
    M = 256; K = M // 2 + 1; h, g = proto(M)
    x, d, mpos, _ = synthetic.make_utterance(0, 2, 16000)
    X0 = ref.analysis(x[0], h, M, m, r)
    res = ref.beamform(x, h, g, d, M, m, r, bf_kind=ref.BF_DS)
    save("ds_c2_m256", x=x, delays=d, X0=X0[:, :K], Y=res["Y"][:, :K], time=res["time"], w=res["w"])

    # g2: configs[1] shape (8-mic GSC, M=512) with the NLMS sidelobe canceller, 0.75 s, short warm-up so adaptation shows
    M = 512; K = M // 2 + 1; h, g = proto(M)
    x, d, mpos, _ = synthetic.make_utterance(1, 8, 12000, target_start_s=0.25)
    lms = dict(min_frames=10)
    res = ref.beamform(x, h, g, d, M, m, r, bf_kind=ref.BF_GSC_LMS, lms=lms)
    save("gsclms_c8_m512", x=x, delays=d, Y=res["Y"][:, :K], time=res["time"], waH=res["w"], stats=res["stats"], min_frames=10)
    # same input through the static GSC with non-zero active weights + Zelinski (type 2, alpha .7)
    rng = np.random.default_rng(7)
    wa = 0.05 * (rng.standard_normal((K, 7)) + 1j * rng.standard_normal((K, 7)))
    wap = np.stack([wa.real, wa.imag], axis=-1).reshape(K, 14)
    res = ref.beamform(x, h, g, d, M, m, r, bf_kind=ref.BF_GSC, wa=wap, pf=dict(alpha=0.7, type=2))
    wq, B = ref.gsc_weights(M, 8, FS, d)
    save("gsc_zelinski_c8_m512", x=x, delays=d, wa=wa, Y=res["Y"][:, :K], time=res["time"], wq=wq[:K], B=B[:K])

    # g3: configs[2] shape (8-mic SMI-MVDR + Zelinski), 0.75 s, target from 0.25 s, VAD label [[0.25, 0.75]]
    res = ref.beamform(x, h, g, d, M, m, r, bf_kind=ref.BF_SMI_MVDR, smi_label=(0.25, 0.75), mvdr_mu=1e-4,
                       pf=dict(alpha=0.7, type=2))
    save("smimvdr_zelinski_c8_m512", x=x, delays=d, Y=res["Y"][:, :K], time=res["time"], cov=res["cov"], w=res["w"],
         label=np.array([0.25, 0.75]), mu=1e-4)

    # g4: super-directive MVDR (diffuse model), 4 mics M=256, Zelinski type 1 with warm-up
    M = 256; K = M // 2 + 1; h, g = proto(M)
    x, d, mpos, _ = synthetic.make_utterance(2, 4, 8000, target_start_s=0.1)
    res = ref.beamform(x, h, g, d, M, m, r, bf_kind=ref.BF_MVDR_SD, mpos=mpos, mvdr_mu=0.01, pf=dict(alpha=0.6, type=1, min_frames=5))
    save("mvdrsd_zelinski1_c4_m256", x=x, delays=d, mpos=mpos, Y=res["Y"][:, :K], time=res["time"], w=res["w"], mu=0.01)

    # g5: pseudoinverse known answers (float LINPACK SVD)
    rng = np.random.default_rng(3)
    As, invs = [], []
    for _ in range(6):
        a = rng.standard_normal((8, 8)) + 1j * rng.standard_normal((8, 8))
        A = a @ a.conj().T + 0.1 * np.eye(8)
        inv, ok = ref.pseudoinverse(A, 1e-8)
        As.append(A); invs.append(inv)
    save("pseudoinverse", A=np.array(As), inv=np.array(invs))


if __name__ == "__main__":
    main()
