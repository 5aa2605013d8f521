import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from distant_speech_recognition_b200 import _capi, synthetic
from oracle import restate
M, C, U, n = 512, 64, 3, 30000
pr = np.load(os.path.join(ROOT, "tests", "golden", "prototype_M512_m4_r1.npz"))
x, d = synthetic.make_batch(U, C, n, first=410)
lengths = np.array([30000, 21000, 26500], np.int32)
labels = np.array([[0.5, 0.9], [0.2, 0.6], [1.0, 1.4]])
q = _capi.Pipeline(C, M, 4, 1, beamformer=_capi.BF_MVDR, max_utterances=U, max_samples=n); q.set_prototypes(pr["h"], pr["g"])
q.set_delays(d); q.submit(x, lengths); q.run_analysis()
X = q.fetch_snapshots().astype(np.complex128)   # [U][T][C][K]
q.accumulate_covariance(labels=labels, energy_threshold=10.0)
cov = q.get_covariance().astype(np.complex128)
T = X.shape[1]; K = 257
Rref = []
for u in range(U):
    Tu = q.num_frames_of(u)
    e = np.array([abs(np.vdot(np.concatenate([X[u, t, 0, :], np.conj(X[u, t, 0, 1:256][::-1])]), np.concatenate([X[u, t, 0, :], np.conj(X[u, t, 0, 1:256][::-1])]))) / M for t in range(T)])
    wt, wn = restate.sos_label_weights(T, e, 16000.0, 256, [tuple(labels[u])], 10.0)
    wn[Tu:] = 0
    Rref.append(np.einsum("t,tck,tdk->kcd", wn, X[u], np.conj(X[u])) / wn.sum())
nbad = 0
for rep in range(6):
    q.accumulate_covariance(labels=labels, energy_threshold=10.0)
    cv = q.get_covariance().astype(np.complex128)
    bad = [(u, k) for u in range(U) for k in range(K) if np.linalg.norm(cv[u, k] - Rref[u][k]) > 1e-4 * np.linalg.norm(Rref[u][k])]
    nbad += len(bad)
    print("rep", rep, "bad chains", len(bad), bad[:10])
print("DEBUG", os.environ.get("BTKB_COV_TC_DEBUG", "0"), "total bad", nbad)
