"""oracle/pyref.py — run the reference's OWN pure-Python beamformers (lib/pybeamformer.py) in this container.

TEST INFRASTRUCTURE ONLY (build container; /root/reference does not exist on the GPU box).  The reference module is
Python 2 (print statements, numpy.complex) and star-imports the SWIG modules btk20.*, which cannot be built here.  This
loader reads the source text where it lies, applies three kinds of mechanical Python-2 -> 3 token fixes IN MEMORY (nothing is
copied into the repo):

    print <expr>            ->  print(<expr>)
    numpy.complex / .float / .int  ->  complex / float / int
    raise StopIteration     ->  return            (PEP 479, inside generators)

(the file is executed up to, not including, `class SubbandHOSBatchBeamformer`, whose section uses Python-2-only tuple
parameters and is out of scope) and executes it with stub `btk20.*` modules whose only content is a NumPy SnapShotArrayPtr (set_samples / update /
snapshot, beamformer/beamformer.cc:56-70).  The algorithmic code of SubbandGSCLMSBeamformer.__iter__ (pybeamformer.py:
659-734), SubbandGSCRLSBeamformer.__iter__ (:817-901), calc_blocking_matrix (:309-341) etc. then runs unmodified on
array-backed spectral sources.  tests/golden/make_golden_pyref.py uses it to generate golden vectors.
"""
import os
import re
import sys
import types
import numpy as np

REF_PY = "/root/reference/btk20_src/lib/pybeamformer.py"


def available():
    return os.path.exists(REF_PY)


class _SnapShotArray:
    """SnapShotArray (beamformer/beamformer.cc:56-70): per-channel spectra in, per-bin snapshots out."""

    def __init__(self, fftlen, chan_num):
        self._s = np.zeros((chan_num, fftlen), np.complex128)
        self._x = np.zeros((fftlen, chan_num), np.complex128)

    def set_samples(self, samp, chanX):
        self._s[chanX] = samp

    def update(self):
        self._x = self._s.T.copy()

    def snapshot(self, fbinX):
        return self._x[fbinX]


class ArraySpectralSource:
    """A channel's analysis-bank output held as an array X[T][M] (what OverSampledDFTAnalysisBankPtr.next() returns)."""

    def __init__(self, X, shiftlen):
        self._X = np.asarray(X, np.complex128)
        self._shiftlen = shiftlen
        self._t = -1

    def next(self, frame_no=-5):
        if self._t + 1 >= self._X.shape[0]:
            raise StopIteration
        self._t += 1
        return self._X[self._t]

    def size(self):
        return self._X.shape[1]

    def shiftlen(self):
        return self._shiftlen

    def reset(self):
        self._t = -1


_mod = None


def load():
    """Returns the executed reference module (cached)."""
    global _mod
    if _mod is not None:
        return _mod
    src = open(REF_PY).read()
    # the higher-order-statistics beamformers further down (out of scope) use Python-2-only tuple parameters
    cut = src.find("class SubbandHOSBatchBeamformer")
    if cut > 0:
        src = src[:cut]
    out = []
    for line in src.split("\n"):
        m = re.match(r"^(\s*)print (.*)$", line)
        if m and not line.lstrip().startswith("#"):
            line = "%sprint(%s)" % (m.group(1), m.group(2))
        line = re.sub(r"\bnumpy\.complex\b(?!\d)", "complex", line)
        line = re.sub(r"\bnumpy\.float\b(?!\d)", "float", line)
        line = re.sub(r"\bnumpy\.int\b(?!\d)", "int", line)
        line = re.sub(r"^(\s*)raise StopIteration\s*$", r"\1return", line)
        out.append(line)
    code = "\n".join(out)
    stubs = {}
    for name in ("btk20", "btk20.common", "btk20.stream", "btk20.feature", "btk20.modulated", "btk20.beamformer"):
        stubs[name] = types.ModuleType(name)
    stubs["btk20.beamformer"].SnapShotArrayPtr = _SnapShotArray
    saved = {k: sys.modules.get(k) for k in stubs}
    sys.modules.update(stubs)
    try:
        mod = types.ModuleType("ref_pybeamformer")
        mod.__file__ = REF_PY
        exec(compile(code, REF_PY, "exec"), mod.__dict__)
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    _mod = mod
    return mod


def run_adaptive(kind, X, samplerate, delays, D, **params):
    """Run the reference's SubbandGSCLMSBeamformer ('lms') or SubbandGSCRLSBeamformer ('rls') over snapshots X[T][C][M].
    Returns (Y[T][M], waH[K][C-1], ttl_updates)."""
    mod = load()
    X = np.asarray(X)
    T, C, M = X.shape
    srcs = [ArraySpectralSource(X[:, c, :], D) for c in range(C)]
    cls = mod.SubbandGSCLMSBeamformer if kind == "lms" else mod.SubbandGSCRLSBeamformer
    bf = cls(srcs, **params)
    bf.calc_beamformer_weights(samplerate, np.asarray(delays, np.float64))
    Y = []
    it = iter(bf)
    while True:  # Python 2: the sources' StopIteration ends the generator; Python 3 (PEP 479) wraps it in a RuntimeError
        try:
            Y.append(next(it))
        except StopIteration:
            break
        except RuntimeError as e:
            if isinstance(e.__cause__, StopIteration):
                break
            raise
    Y = np.array(Y)
    K = M // 2 + 1
    return Y, np.array(bf._waH)[:K].copy(), bf._ttl_updates


def _drain(bf):
    """Iterate a reference beamformer to the end of its sources (Python 2 semantics of a StopIteration inside a generator)."""
    Y = []
    it = iter(bf)
    while True:
        try:
            Y.append(next(it))
        except StopIteration:
            break
        except RuntimeError as e:
            if isinstance(e.__cause__, StopIteration):
                break
            raise
    return np.array(Y)


def run_sos(kind, X, samplerate, D, labels=None, mask_t=None, mask_j=None, energy_threshold=10, gamma=1e-6, ref_micx=0, offset=0.0):
    """Run the reference's SubbandBlindMVDRBeamformer ('bmvdr', pybeamformer.py:1243-1295) or SubbandGEVBeamformer ('gev',
    :1298-1357) over snapshots X[T][C][M], statistics from a VAD label (accu_stats_from_label, :1063-1127) or TF masks
    (accu_stats_from_tfmask, :1129-1183), exactly in the order of unit_test/test_sos_batch_beamforming.py:186-210.
    Returns dict(Y[T][M], wqH[K][C], Rt, Rn [K][C][C] after finalize_stats, ct, cn [K])."""
    mod = load()
    X = np.asarray(X)
    T, C, M = X.shape
    srcs = [ArraySpectralSource(X[:, c, :], D) for c in range(C)]
    bf = (mod.SubbandBlindMVDRBeamformer if kind == "bmvdr" else mod.SubbandGEVBeamformer)(srcs)
    if mask_t is not None:
        bf.accu_stats_from_tfmask(samplerate, mask_t, mask_j, energy_threshold=energy_threshold)
    else:
        bf.accu_stats_from_label(samplerate, target_labs=labels, energy_threshold=energy_threshold)
    ct, cn = np.array(bf._target_frame_counts), np.array(bf._noise_frame_counts)
    bf.finalize_stats(gamma=gamma)
    if kind == "bmvdr":
        bf.calc_beamformer_weights(ref_micx=ref_micx, offset=offset)
    else:
        bf.calc_beamformer_weights()
    bf.reset()
    Y = _drain(bf)
    return dict(Y=Y, wqH=np.array(bf._wqH), Rt=np.array(bf._target_covariance_matrices), Rn=np.array(bf._noise_covariance_matrices), ct=ct, cn=cn)
