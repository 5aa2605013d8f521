"""Host-side interface containers of the btk20 mirror that carry no GPU work (CPU only): SnapShotArray / SpectralMatrixArray
(beamformer/spectralinfoarray.h, beamformer.cc:95-143; SWIG surface beamformer/beamformer.i:46-114)."""
import numpy as np
import pytest

from oracle import restate


def test_spectral_matrix_array_recursion():
    from distant_speech_recognition_b200.btk20.beamformer import SpectralMatrixArrayPtr, SnapShotArrayPtr
    M, C, mu = 16, 3, 0.9
    s = SpectralMatrixArrayPtr(M, C, mu)
    rng = np.random.default_rng(0)
    R = np.zeros((M, C, C), complex)
    for it in range(4):
        x = rng.standard_normal((C, M)) + 1j * rng.standard_normal((C, M))
        for c in range(C):
            s.set_samples(x[c], c)
        s.update()
        R = restate.spectral_matrix_update(R, x.T, float(np.float32(mu)), legacy_noconj=True)   # forgetFact is a float (beamformer.cc:97-100)
        assert np.allclose(np.array(s.snapshot(5)), x[:, 5])                                    # SnapShotArray::update ran too
    got = np.stack([np.array(s.matrix_f(f)) for f in range(M)])
    assert np.abs(got - R).max() < 1e-14
    with pytest.raises(Exception):
        s.matrix_f(M)
    s.zero()
    assert np.all(np.array(s.matrix_f(3)) == 0) and np.all(np.array(s.snapshot(3)) == 0)
    assert isinstance(s, SnapShotArrayPtr)
