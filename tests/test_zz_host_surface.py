"""The rest of the reference's SWIG surface for the hot-path classes: stream handle classes of every element
type, SubbandDS LCMV manifolds, SubbandGSC.normalize_weight / set_quiescent_weights_f / write_fir_coeff, SubbandMVDR per-bin
covariance edits, btkb_get_sidelobe_weights.

CPU part: adapters, argument checks and call-order errors (no pipeline is created).  GPU part (-m gpu): each method against
oracle/restate.py or the reference's goldens.  The file sorts last on purpose: these GPU tests were added after the round's
GPU budget was spent; they ran green on a B200 at the end of round 1."""
import os
import numpy as np
import pytest

from conftest import load_golden, rel_l2

btk20 = pytest.importorskip("distant_speech_recognition_b200.btk20")
from distant_speech_recognition_b200.btk20 import stream  # noqa: E402
from distant_speech_recognition_b200.btk20.feature import SampleFeaturePtr  # noqa: E402
from distant_speech_recognition_b200.btk20.modulated import OverSampledDFTAnalysisBankPtr, get_window  # noqa: E402
from distant_speech_recognition_b200.btk20.beamformer import SubbandDSPtr, SubbandGSCPtr, SubbandMVDRPtr, SubbandMVDRGSCPtr  # noqa: E402

FS = 16000
TOL = 1e-4


class _Seq:
    """A Python feature object as stream/pyStream.h expects it: size(), __iter__/__next__, reset()."""
    def __init__(self, rows):
        self.rows, self.resets = rows, 0
    def size(self):
        return len(self.rows[0])
    def __iter__(self):
        return iter(self.rows)
    def reset(self):
        self.resets += 1


@pytest.mark.parametrize("cls,dtype", [("PyVectorFloatFeatureStreamPtr", np.float32), ("PyVectorFeatureStreamPtr", np.float64),
                                       ("PyVectorShortFeatureStreamPtr", np.int16), ("PyVectorComplexFeatureStreamPtr", np.complex128)])
def test_python_stream_adapters_of_every_element_type(cls, dtype):
    rows = [(np.arange(6) + 10 * i).astype(dtype) for i in range(3)]
    src = _Seq(rows)
    s = getattr(stream, cls)(src, "adapter")
    assert s.size() == 6 and s.name() == "adapter" and s.frame_no() == -1
    with pytest.raises(Exception, match="Frame index"):  # current() before the first frame: jconsistency_error (stream.h:31-36)
        s.current()
    got = [np.array(v) for v in s]                       # __iter__ resets the stream (stream.i), then pulls to StopIteration
    assert len(got) == 3 and all(np.array_equal(a, b) for a, b in zip(got, rows)) and got[0].dtype == dtype
    assert src.resets == 1 and s.is_end() and s.frame_no() == 2
    assert np.array_equal(np.array(s.next(2)), rows[2])  # frame_no == frame_no(): the cached frame again
    assert np.array_equal(np.array(s.current()), rows[2])
    s.reset()
    assert src.resets == 2 and not s.is_end() and s.frame_no() == -1
    bad = getattr(stream, cls)(_Seq([np.zeros(6, dtype), np.zeros(5, dtype)]), "bad")
    bad.next()
    with pytest.raises(Exception):                       # a frame of the wrong length: jdimension_error
        bad.next()


def test_stream_handle_classes_exist():
    for n in ("VectorCharFeatureStreamPtr", "VectorShortFeatureStreamPtr", "VectorFloatFeatureStreamPtr", "VectorFeatureStreamPtr",
              "VectorComplexFeatureStreamPtr"):
        c = getattr(stream, n)
        for meth in ("next", "reset", "size", "is_end", "frame_no", "name", "__iter__"):
            assert hasattr(c, meth), (n, meth)


def _afbs(x, h, M, D):
    out = []
    for c in range(x.shape[0]):
        s = SampleFeaturePtr(block_len=D, shift_len=D, pad_zeros=True)
        s.setSamples(np.asarray(x[c], np.float64), FS)
        out.append(OverSampledDFTAnalysisBankPtr(s, prototype=h, M=M, m=4, r=1, delay_compensation_type=2))
    return out


def test_argument_checks_of_the_added_methods(protos, tmp_path):
    h, _ = protos[256]; M, D, C = 256, 128, 4
    afbs = _afbs(np.zeros((C, 1000)), h, M, D)
    ds = SubbandDSPtr(fftlen=M)
    for a in afbs:
        ds.set_channel(a)
    d = np.zeros(C)
    with pytest.raises(Exception):            # 1 < NC <= chanN (beamformer.cc:592-594)
        ds.calc_array_manifold_vectors_n(FS, d, np.zeros((4, C)), NC=5)
    with pytest.raises(Exception):            # delays_j must be [NC-1][C]
        ds.calc_array_manifold_vectors_n(FS, d, np.zeros((1, C)), NC=3)
    with pytest.raises(Exception):            # target delays / channels mismatch (beamformer.cc:504-506)
        ds.calc_array_manifold_vectors_2(FS, np.zeros(3), d)
    ds.calc_array_manifold_vectors_2(FS, d, d + 1e-4)
    ds.calc_array_manifold_vectors_n(FS, d, np.stack([d + 1e-4, d - 1e-4]), NC=3)

    gsc = SubbandGSCPtr(fftlen=M)
    for a in afbs:
        gsc.set_channel(a)
    assert gsc.write_fir_coeff(str(tmp_path / "fir.txt")) is False      # "call calc_array_manifold_vectorsX() once" (beamformer.cc:1366-1369)
    with pytest.raises(Exception, match="calc_gsc_weights"):
        gsc.set_active_weights_f(3, np.zeros(2 * (C - 1)))
    with pytest.raises(Exception):            # wrong length of the quiescent vector
        gsc.set_quiescent_weights_f(3, np.ones(C + 1, complex))
    with pytest.raises(IndexError):
        gsc.set_quiescent_weights_f(M, np.ones(C, complex))
    gsc.set_quiescent_weights_f(3, np.ones(C, complex) / C)
    gsc.set_active_weights_f(3, np.zeros(2 * (C - 1)))                   # allowed now: a weight object exists
    gsc.normalize_weight(True); gsc.normalize_weight(False)

    mv = SubbandMVDRPtr(fftlen=M)
    for a in afbs:
        mv.set_channel(a)
    for call in (lambda: mv.set_diagonal_looading(2, 0.1), lambda: mv.divide_nondiagonal_elements(2, 0.1), lambda: mv.divide_all_nondiagonal_elements(0.1)):
        with pytest.raises(Exception, match="noise covariance"):         # "Construct first a noise covariance matrix" (beamformer.cc:2527-2529)
            call()
    assert mv.set_noise_spatial_spectral_matrix(2, np.eye(C, dtype=complex))
    mv.set_diagonal_looading(2, 0.1); mv.divide_nondiagonal_elements(2, 0.5); mv.divide_all_nondiagonal_elements(0.01)
    with pytest.raises(IndexError):
        mv.set_diagonal_looading(M // 2 + 1, 0.1)
    mg = SubbandMVDRGSCPtr(fftlen=M)
    for a in afbs:
        mg.set_channel(a)
    assert mg.calc_blocking_matrix2() is False                           # no MVDR weights yet (beamformer.cc:2651-2653)
    assert mg.calc_blocking_matrix1(FS, d) is True


# ---------------------------------------------------------------------------------------------------------------- GPU
# (round 1 carried these behind a non-strict xfail marker; all of them passed on the B200 at the end of round 1 — GPUTEST_r01.json —
# and the marker is gone: a failure here fails the suite.)


@pytest.fixture(scope="module")
def capi():
    from distant_speech_recognition_b200 import _capi
    assert _capi.device_count() >= 1, "no CUDA device: the product has no CPU path"
    return _capi


def _restate_static(g, h, M, C):
    from oracle import restate
    X = np.stack([restate.analysis(g["x"][c], h, M, 4, 1) for c in range(C)], axis=1)
    wq = restate.calc_mainlobe(M, C, FS, g["delays"])
    K = M // 2 + 1
    wl = np.zeros_like(wq)
    wl[:K] = restate.active_to_wl(np.stack([restate.calc_blocking_matrix(wq[k]) for k in range(K)]), g["wa"])
    for k in range(1, M // 2):
        wl[M - k] = np.conj(wl[k])
    return X, wq, wl


@pytest.mark.gpu
def test_normalize_weight_and_sidelobe_weights(capi, protos):
    """calc_gsc_output(normalizeWeight = true), beamformer.cc:1230-1236: w <- w / (||w|| C) for bins >= 1, DC bin untouched;
    btkb_get_sidelobe_weights = wl = B wa."""
    from oracle import restate
    g = load_golden("gsc_zelinski_c8_m512"); h, gg = protos[512]; M, C, K = 512, 8, 257
    X, wq, wl = _restate_static(g, h, M, C)
    for flag in (False, True):
        p = capi.Pipeline(C, M, 4, 1, beamformer=capi.BF_GSC, max_utterances=1, max_samples=g["x"].shape[1], normalize_weight=flag)
        p.set_prototypes(h, gg)
        p.set_delays(g["delays"][None])
        assert np.all(p.get_sidelobe_weights() == 0)        # zero_active_weights state
        p.set_active_weights(g["wa"][None])
        assert rel_l2(p.get_sidelobe_weights()[0], wl[:K]) < 1e-5
        p.submit(g["x"][None]); p.run(True)
        Yo = restate.subband_gsc(X, wq, wl, normalize_weight=flag)
        assert rel_l2(p.fetch_subband()[0], Yo[:, :K]) < TOL, flag
        assert rel_l2(p.fetch_time()[0], restate.synthesis(Yo, gg, M, 4, 1)) < TOL, flag
        p.close()
    # the host mirror's switch reaches the device configuration
    afbs = _afbs(g["x"], h, M, 256)
    bf = SubbandGSCPtr(fftlen=M)
    for a in afbs:
        bf.set_channel(a)
    bf.calc_gsc_weights(FS, g["delays"])
    for k in range(K):
        bf.set_active_weights_f(k, np.stack([g["wa"][k].real, g["wa"][k].imag], axis=1).ravel())
    bf.normalize_weight(True)
    Y = np.array([np.array(v) for v in bf])
    assert rel_l2(Y[:, :K], Yo[:, :K]) < TOL


@pytest.mark.gpu
def test_ds_lcmv_manifolds_and_mvdr_guard(protos):
    """SubbandDS::calc_array_manifold_vectors_2 / _n (beamformer.cc:1057-1074): LCMV weights without a sidelobe canceller, vs the
    reference's calcMainlobe2 / calcMainlobeN golden; y = w^H x for every bin."""
    from oracle import restate
    from distant_speech_recognition_b200 import synthetic
    g = load_golden("lcmv"); h, _ = protos[512]; M, D, C, K = 512, 256, 8, 257
    x, _, _, _ = synthetic.make_utterance(33, C, 6000)
    afbs = _afbs(x, h, M, D)
    X = np.stack([restate.analysis(x[c], h, M, 4, 1) for c in range(C)], axis=1)
    for name, call in (("w2", lambda b: b.calc_array_manifold_vectors_2(FS, g["dT"], g["dJ1"])),
                       ("w3", lambda b: b.calc_array_manifold_vectors_n(FS, g["dT"], np.stack([g["dJ1"], g["dJ2"]]), NC=3))):
        ds = SubbandDSPtr(fftlen=M)
        for a in afbs:
            ds.set_channel(a)
        call(ds)
        W = np.array([np.array(ds.get_weights(k)) for k in range(K)])
        if name == "w2":
            assert rel_l2(W, g[name]) < 1e-5                       # 2 x 2 closed-form inverse: every bin
        else:                                                      # NC = 3: the reference's float SVD is rounding noise at the lowest bins
            assert rel_l2(W[16:256], g[name][16:256]) < 2e-5       # (tests/test_parity_gpu.py::test_lcmv_quiescent_weights_golden)
        Y = np.array([np.array(v) for v in ds])
        wfull = np.concatenate([W, np.conj(W[1:256][::-1])])       # y = w^H x with the weights the object reports, DC bin included
        assert rel_l2(Y[:, :K], restate.subband_ds(X, wfull)[:, :K]) < TOL, name
    mv = SubbandMVDRPtr(fftlen=M)
    for a in afbs:
        mv.set_channel(a)
    mv.calc_array_manifold_vectors_2(FS, g["dT"], g["dJ1"])
    mv.set_noise_spatial_spectral_matrix(0, np.eye(C, dtype=complex))
    mv.calc_mvdr_weights(FS)
    with pytest.raises(Exception, match="LCMV"):
        mv.next()


@pytest.mark.gpu
def test_write_fir_coeff(protos, tmp_path):
    """BeamformerWeights::write_fir_coeff (beamformer.cc:775-828): header, one row per channel, coefficient n =
    window[n] Re(IDFT_f(conj(wq - wl) e^{j pi (f+1)}))."""
    g = load_golden("gsc_zelinski_c8_m512"); h, _ = protos[512]; M, C, K = 512, 8, 257
    _, wq, wl = _restate_static(g, h, M, C)
    afbs = _afbs(g["x"], h, M, 256)
    bf = SubbandGSCPtr(fftlen=M)
    for a in afbs:
        bf.set_channel(a)
    bf.calc_gsc_weights(FS, g["delays"])
    for k in range(K):
        bf.set_active_weights_f(k, np.stack([g["wa"][k].real, g["wa"][k].imag], axis=1).ravel())
    for win in (1, 0, 2):
        fn = str(tmp_path / ("fir%d.txt" % win))
        assert bf.write_fir_coeff(fn, win) is True
        lines = open(fn).read().strip().split("\n")
        assert lines[0].split() == [str(C), str(M)] and len(lines) == C + 1
        got = np.array([[float(v) for v in ln.split()] for ln in lines[1:]])
        spec = np.zeros((M, C), complex)
        f = np.arange(K)
        spec[:K] = np.exp(1j * np.pi * (f + 1))[:, None] * np.conj(wq[:K] - wl[:K])
        spec[K:] = np.conj(spec[1:M // 2][::-1])
        want = np.real(np.fft.ifft(spec, axis=0)).T * get_window(win, M)[None]
        assert got.shape == want.shape and rel_l2(got, want) < 1e-5, win


@pytest.mark.gpu
def test_set_quiescent_weights_f_keeps_only_the_last_bin(protos):
    """SubbandGSC::set_quiescent_weights_f re-allocates the weight object on every call (beamformer.cc:1318-1324 -> alloc_bfweight_):
    after two calls only the second bin carries a quiescent vector, every other bin outputs zero."""
    from oracle import restate
    from distant_speech_recognition_b200 import synthetic
    h, _ = protos[256]; M, D, C, K = 256, 128, 4, 129
    x, _, _, _ = synthetic.make_utterance(35, C, 4000)
    afbs = _afbs(x, h, M, D)
    bf = SubbandGSCPtr(fftlen=M)
    for a in afbs:
        bf.set_channel(a)
    rng = np.random.default_rng(3)
    w5 = rng.standard_normal(C) + 1j * rng.standard_normal(C); w9 = rng.standard_normal(C) + 1j * rng.standard_normal(C)
    bf.set_quiescent_weights_f(5, w5)
    bf.set_quiescent_weights_f(9, w9)
    assert np.all(np.array(bf.get_weights(5)) == 0) and rel_l2(np.array(bf.get_weights(9)), w9) < 1e-6
    Y = np.array([np.array(v) for v in bf])
    X = np.stack([restate.analysis(x[c], h, M, 4, 1) for c in range(C)], axis=1)
    assert rel_l2(Y[:, 9], X[:, :, 9] @ np.conj(w9)) < TOL
    keep = np.ones(M, bool); keep[[9, M - 9]] = False
    assert np.all(Y[:, keep] == 0)


@pytest.mark.gpu
def test_mvdr_per_bin_covariance_edits(protos):
    """set_diagonal_looading(fbinX, w), divide_nondiagonal_elements(fbinX, mu), divide_all_nondiagonal_elements(mu)
    (beamformer.cc:2525-2535, 2589-2599, beamformer.h:357-360) on the diffuse noise model, then calc_mvdr_weights: weights vs the
    fp64 solve of the edited matrices."""
    from oracle import restate
    g = load_golden("mvdrsd_zelinski1_c4_m256"); h, _ = protos[256]; M, D, C, K = 256, 128, 4, 129
    afbs = _afbs(g["x"], h, M, D)
    mv = SubbandMVDRPtr(fftlen=M)
    for a in afbs:
        mv.set_channel(a)
    mv.calc_array_manifold_vectors(FS, g["delays"])
    assert mv.set_diffuse_noise_model(g["mpos"], FS, 343740.0)
    mv.set_all_diagonal_loading(0.01)
    mv.divide_all_nondiagonal_elements(0.05)
    mv.set_diagonal_looading(7, 0.5)
    mv.divide_nondiagonal_elements(11, 1.0)
    assert mv.calc_mvdr_weights(FS, 1e-8, True)
    R = restate.diffuse_noise_model(M, g["mpos"], FS)[:K].astype(complex)
    eye = np.eye(C, dtype=bool)
    R[:, eye] += float(np.float32(0.01))
    R[:, ~eye] /= 1.0 + float(np.float32(0.05))
    R[7][eye] += 0.5
    R[11][~eye] /= 2.0
    wq = restate.calc_mainlobe(M, C, FS, g["delays"])
    want = restate.calc_mvdr_weights(R, wq[:K], single=False)
    W = np.array([np.array(mv.mvdr_weights(k)) for k in range(K)])
    assert rel_l2(W[1:], want[1:]) < 1e-4
    assert rel_l2(W[7], want[7]) < 1e-4 and rel_l2(W[11], want[11]) < 1e-4


@pytest.mark.gpu
def test_mvdrgsc_blocking_matrix_from_the_mvdr_weights(protos):
    """SubbandMVDRGSC::calc_blocking_matrix2 (beamformer.cc:2649-2672): B orthogonal to w_mvdr instead of the delay-and-sum weights;
    y[f >= 1] = (w_mvdr - B wa)^H x, the DC bin uses w_mvdr only (beamformer.cc:2750-2768)."""
    from oracle import restate
    g = load_golden("mvdrsd_zelinski1_c4_m256"); h, _ = protos[256]; M, D, C, K = 256, 128, 4, 129
    afbs = _afbs(g["x"], h, M, D)
    X = np.stack([restate.analysis(g["x"][c], h, M, 4, 1) for c in range(C)], axis=1)
    rng = np.random.default_rng(7)
    wa = 0.05 * (rng.standard_normal((K, C - 1)) + 1j * rng.standard_normal((K, C - 1)))
    wq = restate.calc_mainlobe(M, C, FS, g["delays"])
    R = restate.diffuse_noise_model(M, g["mpos"], FS)[:K].astype(complex)
    R[:, np.eye(C, dtype=bool)] += float(np.float32(g["mu"]))
    wm = np.zeros((M, C), complex); wm[:K] = restate.calc_mvdr_weights(R, wq[:K], single=False)
    outs = {}
    for variant in (1, 2):
        bf = SubbandMVDRGSCPtr(fftlen=M)
        for a in afbs:
            bf.set_channel(a)
        bf.calc_array_manifold_vectors(FS, g["delays"])
        bf.set_diffuse_noise_model(g["mpos"], FS, 343740.0)
        bf.set_all_diagonal_loading(float(g["mu"]))
        bf.calc_mvdr_weights(FS, 1e-8, True)
        if variant == 1:
            assert bf.calc_blocking_matrix1(FS, g["delays"])
            src = wq
        else:
            assert bf.calc_blocking_matrix2()
            src = wm
        for k in range(K):
            bf.set_active_weights_f(k, np.stack([wa[k].real, wa[k].imag], axis=1).ravel())
        wl = np.zeros((M, C), complex)
        wl[:K] = restate.active_to_wl(np.stack([restate.calc_blocking_matrix(src[k]) for k in range(K)]), wa)
        Y = np.array([np.array(v) for v in bf])
        outs[variant] = Y
        assert rel_l2(Y[:, :K], restate.subband_mvdr(X, wm, wl)[:, :K]) < TOL, variant
    assert rel_l2(outs[1][:, 1:K], outs[2][:, 1:K]) > 1e-3          # the two blocking matrices really differ


@pytest.mark.gpu
def test_online_beamforming_on_the_references_own_fixtures(capi):
    """unit_test/test_online_beamforming.py on its own inputs (Kinect recording as 16-bit PCM, shipped M = 256 prototypes) with its own
    parameter files confs/{ds, ds_and_zelinski, sd, sd_and_zelinski, sd_and_mccowan, sd_and_lefkimmiatis, gsclms, gscrls}.json, through
    the C-ABI, against the reference's outputs (golden_online_kinect_c4_m256; the oracle is pinned on the same file in
    tests/test_oracle.py::test_online_beamforming_on_the_references_own_fixtures)."""
    import os
    from conftest import GOLDEN
    g = load_golden("online_kinect_c4_m256")
    pr = np.load(os.path.join(GOLDEN, "prototype_shipped_M256_m4_r1.npz"))
    M, C = 256, 4
    d, mpos = g["delays"], g["mpos"]
    s0, n = int(g["s0_static"]), int(g["n_static"])
    xs = np.ascontiguousarray(g["x16"][:, s0:s0 + n])

    def run(name, x16, setup=None, **kw):
        p = capi.Pipeline(C, M, 4, 1, max_utterances=1, max_samples=x16.shape[1], **kw)
        p.set_prototypes(pr["h"], pr["g"])
        p.set_delays(d[None])
        if setup:
            setup(p)
        p.submit_i16(x16[None])
        p.run(True)
        assert rel_l2(p.fetch_subband()[0], g["Y_" + name]) < TOL, name
        t = p.fetch_time()[0]
        assert rel_l2(t, g["time_" + name]) < TOL, name
        assert abs(p.fetch_stats()[0, 0] / float(g["energy_" + name]) - 1.0) < 1e-3, name     # the script's report: total_energy
        return p

    def sd(p):
        p.set_diffuse_noise_model(1, mpos)
        p.calc_mvdr_weights(0.01)

    def coherence(load):
        def f(p):
            sd(p)
            p.pf_set_diffuse_noise_model(mpos, float(FS))
            p.pf_set_diagonal_loading(load)
        return f

    run("ds", xs, beamformer=capi.BF_GSC).close()
    run("ds_and_zelinski", xs, beamformer=capi.BF_GSC, postfilter=capi.PF_ZELINSKI, pf_alpha=0.7, pf_type=2).close()
    p = run("sd", xs, sd, beamformer=capi.BF_MVDR)
    assert rel_l2(p.get_weights()[0][1:], g["w_sd"][1:]) < TOL
    p.close()
    run("sd_and_zelinski", xs, sd, beamformer=capi.BF_MVDR, postfilter=capi.PF_ZELINSKI, pf_alpha=0.7, pf_type=2).close()
    run("sd_and_mccowan", xs, coherence(0.01), beamformer=capi.BF_MVDR, postfilter=capi.PF_MCCOWAN, pf_alpha=0.7, pf_type=2).close()
    run("sd_and_lefkimmiatis", xs, coherence(0.1), beamformer=capi.BF_MVDR, postfilter=capi.PF_LEFKIMMIATIS, pf_alpha=0.8, pf_type=2, pf_min_sv=1e-4,
        pf_fbin1=100).close()
    xf = np.ascontiguousarray(g["x16"])
    p = run("gsclms", xf, beamformer=capi.BF_GSC_LMS)               # confs/gsclms.json = the defaults of btkb_default_config
    assert int(p.fetch_stats()[0, 2]) == int(g["n_updates_gsclms"])
    assert rel_l2(p.get_active_weights()[0], g["waH_gsclms"]) < 1e-3
    p.close()
    p = run("gscrls", xf, beamformer=capi.BF_GSC_RLS)               # confs/gscrls.json = the defaults
    assert int(p.fetch_stats()[0, 2]) == int(g["n_updates_gscrls"])
    p.close()
    # confs/lcmv_and_zelinski.json: LCMV quiescent weights vs the reference's calcMainlobeN (every bin, incl. its f = M/2 cascade), then
    # static GSC + Zelinski on the excerpt vs the restatement run on the reference's weights
    from oracle import restate
    p = capi.Pipeline(C, M, 4, 1, max_utterances=1, max_samples=xs.shape[1], beamformer=capi.BF_GSC, postfilter=capi.PF_ZELINSKI, pf_alpha=0.7, pf_type=2)
    p.set_prototypes(pr["h"], pr["g"])
    p.set_delays_lcmv(g["lcmv_dT"][None], g["lcmv_dJ"][None])
    assert rel_l2(p.get_weights()[0], g["lcmv_w"]) < 1e-5
    p.submit_i16(xs[None]); p.run(True)
    X = np.stack([restate.analysis(xs[c].astype(np.float32), pr["h"], M, 4, 1) for c in range(C)], axis=1)
    wfull = np.concatenate([g["lcmv_w"], np.conj(g["lcmv_w"][1:M // 2][::-1])])
    Yo = restate.zelinski_postfilter(restate.subband_gsc(X, wfull, np.zeros_like(wfull)), X, restate.calc_mainlobe(M, C, float(FS), g["lcmv_dT"]), 0.7, 2, 0)[0]
    assert rel_l2(p.fetch_subband()[0], Yo[:, :M // 2 + 1]) < TOL
    assert rel_l2(p.fetch_time()[0], restate.synthesis(Yo, pr["g"], M, 4, 1)) < TOL
    p.close()


@pytest.mark.gpu
def test_sos_batch_beamforming_vad_on_the_references_own_fixtures(capi):
    """unit_test/test_sos_batch_beamforming.py on the whole Kinect recording with confs/{bmvdr_vad, gev_vad, smimvdr}.json (VAD label
    [[1.5, 4.0]]) through the C-ABI, against the reference's outputs (golden_sos_kinect_vad_c4_m256)."""
    import os
    from conftest import GOLDEN
    g = load_golden("sos_kinect_vad_c4_m256"); x16 = np.ascontiguousarray(load_golden("online_kinect_c4_m256")["x16"])
    pr = np.load(os.path.join(GOLDEN, "prototype_shipped_M256_m4_r1.npz"))
    M, C = 256, 4
    f0, f1 = [int(v) for v in g["frames"]]
    for name, kind in (("bmvdr_vad", capi.SOS_BMVDR), ("gev_vad", capi.SOS_GEV)):
        p = capi.Pipeline(C, M, 4, 1, beamformer=capi.BF_DS, max_utterances=1, max_samples=x16.shape[1])
        p.set_prototypes(pr["h"], pr["g"])
        p.submit_i16(x16[None])
        p.run_analysis()
        p.sos_accumulate_from_label(g["labels"], 10.0)
        _, _, cnt = p.sos_get_stats()
        assert np.array_equal(cnt[0, :, 0], g["ct"]) and np.array_equal(cnt[0, :, 1], g["cn"])
        p.sos_calc_weights(kind, gamma=1e-6, ref_micx=0, offset=0.0)
        w = p.get_weights()[0]
        sgn = 1.0 if kind == capi.SOS_BMVDR else float(np.sign(np.real(np.vdot(w[0], g["w_" + name][0]))))
        assert rel_l2(sgn * w, g["w_" + name]) < 5e-4, name
        if sgn < 0:
            p.set_weights((sgn * w)[None])
        p.run_beamformer(True)
        assert rel_l2(p.fetch_subband()[0][f0:f1], g["Y_" + name]) < TOL, name
        assert rel_l2(p.fetch_time()[0], g["time_" + name]) < TOL, name
        p.close()
    p = capi.Pipeline(C, M, 4, 1, beamformer=capi.BF_MVDR, max_utterances=1, max_samples=x16.shape[1])
    p.set_prototypes(pr["h"], pr["g"])
    p.set_delays(g["delays"][None])
    p.submit_i16(x16[None])
    p.run_analysis()
    p.accumulate_covariance(labels=g["labels"][:1], energy_threshold=10.0)
    assert rel_l2(p.get_covariance()[0], g["cov_smimvdr"]) < 1e-5
    p.calc_mvdr_weights(float(g["mu_smimvdr"]))
    assert rel_l2(p.get_weights()[0][1:], g["w_smimvdr"][1:]) < 2e-4      # fp64 LU here, float LINPACK SVD there
    p.run_beamformer(True)
    assert rel_l2(p.fetch_subband()[0][f0:f1], g["Y_smimvdr"]) < TOL
    assert rel_l2(p.fetch_time()[0], g["time_smimvdr"]) < TOL
    p.close()


@pytest.mark.gpu
def test_wpe_on_the_references_own_fixtures(capi):
    """unit_test/test_subband_dereverberator.py on the whole Kinect recording with confs/wpe.json (lags 0..32: 132 x 132 normal
    equations per bin and channel) through the C-ABI, multi- and single-channel, against the compiled reference's outputs
    (golden_wpe_kinect_c4_m256)."""
    import json
    import os
    from conftest import GOLDEN
    g = load_golden("wpe_kinect_c4_m256"); x16 = np.ascontiguousarray(load_golden("online_kinect_c4_m256")["x16"])
    pr = np.load(os.path.join(GOLDEN, "prototype_shipped_M256_m4_r1.npz"))
    M = 256
    conf = json.loads(str(g["conf"]))
    f0, f1 = [int(v) for v in g["frames"]]
    p = capi.Pipeline(4, M, 4, 1, beamformer=capi.BF_DS, max_utterances=1, max_samples=x16.shape[1], wpe=conf)
    p.set_prototypes(pr["h"], pr["g"])
    p.submit_i16(x16[None]); p.run_analysis(); p.run_wpe()
    Xd = p.fetch_snapshots()
    assert rel_l2(Xd[0][f0:f1], g["X_multi"]) < TOL
    q = capi.Pipeline(1, M, 4, 1, beamformer=capi.BF_DS, max_utterances=4, max_samples=x16.shape[1])   # every channel's stream into a synthesis bank
    q.set_prototypes(pr["h"], pr["g"])
    q.set_subband(np.ascontiguousarray(np.transpose(Xd[0], (1, 0, 2)))); q.run_synthesis()
    assert rel_l2(q.fetch_time(), g["time_multi"]) < TOL
    q.close(); p.close()
    ks = dict(conf); ks["diagonal_bias"] = 0.0
    p = capi.Pipeline(1, M, 4, 1, beamformer=capi.BF_DS, max_utterances=1, max_samples=x16.shape[1], wpe=ks)
    p.set_prototypes(pr["h"], pr["g"])
    p.submit_i16(x16[None, :1]); p.run_analysis(); p.run_wpe()
    Xs = p.fetch_snapshots()
    assert rel_l2(Xs[0][f0:f1, 0], g["X_single"]) < TOL
    q = capi.Pipeline(1, M, 4, 1, beamformer=capi.BF_DS, max_utterances=1, max_samples=x16.shape[1])
    q.set_prototypes(pr["h"], pr["g"])
    q.set_subband(Xs[:, :, 0, :]); q.run_synthesis()
    assert rel_l2(q.fetch_time()[0], g["time_single"]) < TOL
    q.close(); p.close()


# ------------------------------------------------------------------------------------------- a C++ user of the host mirror
def _cpp_frontend(tmp_path):
    """tests/host/cpp_frontend.cc: SampleFeature -> analysis banks -> SubbandGSCLMS -> synthesis bank driven from C++ exactly like the
    reference's C++ programs (src/filterBankTest.cc:189-196), linked against the host mirror sources and libbtkb.so."""
    import os
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    pkg = os.path.join(root, "distant_speech_recognition_b200")
    exe = str(tmp_path / "cpp_frontend")
    subprocess.check_call(["g++", "-O2", "-std=c++17", os.path.join(root, "tests", "host", "cpp_frontend.cc"), os.path.join(pkg, "csrc", "host", "btk20_host.cc"),
                           "-o", exe, "-L" + pkg, "-lbtkb", "-Wl,-rpath," + pkg])
    return exe


def _cpp_inputs(tmp_path, h, g, M, x, delays):
    pf, sf = str(tmp_path / "proto.txt"), str(tmp_path / "samples.txt")
    with open(pf, "w") as f:
        f.write("%d 4 1\n" % M); f.write(" ".join("%.17g" % v for v in h) + "\n"); f.write(" ".join("%.17g" % v for v in g) + "\n")
    with open(sf, "w") as f:
        f.write("%d %d\n" % x.shape)
        for c in range(x.shape[0]):
            f.write(" ".join("%d" % v for v in x[c]) + "\n")
    return [pf, sf] + ["%.17g" % d for d in delays]


def test_cpp_user_of_the_host_mirror_builds_and_fails_loudly_without_a_gpu(tmp_path, protos):
    import subprocess
    from distant_speech_recognition_b200 import _capi
    exe = _cpp_frontend(tmp_path)
    if _capi.device_count() > 0:
        pytest.skip("a GPU is present: the run itself is covered by the gpu-marked test")
    h, g = protos[256]
    args = _cpp_inputs(tmp_path, h, g, 256, np.zeros((2, 1000), np.int16), [0.0, 1e-4])
    r = subprocess.run([exe] + args, capture_output=True, text=True)
    assert r.returncode == 1 and r.stdout.startswith("j_error:") and "no CUDA device" in r.stdout   # no CPU path: the library refuses


@pytest.mark.gpu
def test_cpp_user_of_the_host_mirror_on_the_references_own_fixtures(tmp_path):
    """The C++ program on the Kinect recording with confs/gsclms.json defaults: block count, NLMS update count and the script's
    total_energy report equal the reference's (golden_online_kinect_c4_m256)."""
    import os
    import subprocess
    from conftest import GOLDEN
    g = load_golden("online_kinect_c4_m256")
    pr = np.load(os.path.join(GOLDEN, "prototype_shipped_M256_m4_r1.npz"))
    exe = _cpp_frontend(tmp_path)
    r = subprocess.run([exe] + _cpp_inputs(tmp_path, pr["h"], pr["g"], 256, g["x16"], g["delays"]), capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    tok = r.stdout.split()
    frames, blocks, energy, updates = int(tok[1]), int(tok[3]), float(tok[5]), int(tok[7])
    assert frames == g["Y_gsclms"].shape[0] and blocks * 128 == len(g["time_gsclms"]) and updates == int(g["n_updates_gsclms"])
    assert abs(energy / float(g["energy_gsclms"]) - 1.0) < 1e-3


@pytest.mark.gpu
def test_batch_front_end_configured_from_the_references_parameter_files():
    """btk20.batch.BatchBeamformer built from the reference's own unit_test/confs/*.json (carried verbatim inside
    golden_online_kinect_c4_m256) and btk20.pybeamformer.calc_delays, two copies of the Kinect excerpt / recording per submission,
    against the reference's outputs."""
    import json
    import os
    from conftest import GOLDEN
    from distant_speech_recognition_b200.btk20 import pybeamformer
    from distant_speech_recognition_b200.btk20.batch import BatchBeamformer
    g = load_golden("online_kinect_c4_m256")
    pr = np.load(os.path.join(GOLDEN, "prototype_shipped_M256_m4_r1.npz"))
    s0, n = int(g["s0_static"]), int(g["n_static"])
    for name in ("ds", "ds_and_zelinski", "sd", "sd_and_zelinski", "sd_and_mccowan", "sd_and_lefkimmiatis", "gsclms", "gscrls"):
        c = json.loads(str(g["conf_" + name]))
        x = g["x16"] if name.startswith("gsc") else g["x16"][:, s0:s0 + n]
        x = np.ascontiguousarray(np.stack([x, x]).astype(np.float32))
        d = pybeamformer.calc_delays(c["array_type"], c["microphone_positions"], c["target"]["positions"][0][1])
        assert np.allclose(d, g["delays"], rtol=0, atol=1e-15)
        bf = dict(c["beamformer"])
        if bf["type"] == "delay_and_sum":
            bf["type"] = "gsc"          # the script builds SubbandGSCBeamformer(afbs, Nc=1) with zero active weights for 'delay_and_sum'
        bb = BatchBeamformer(4, pr["h"], pr["g"], M=256, m=4, r=1, samplerate=16000, beamformer=bf, postfilter=c.get("postfilter"),
                             max_utterances=2, max_samples=x.shape[2])
        t, Y, st = bb.process(x, np.stack([d, d]), mpos=c["microphone_positions"])
        for u in range(2):
            assert rel_l2(Y[u], g["Y_" + name]) < TOL, (name, u)
            assert rel_l2(t[u], g["time_" + name]) < TOL, (name, u)
        bb.pipe.close()


# ------------------------------------------------------------------------------- packed 2 x fp32 variants, one test per kernel
def _ulp_report(a, b):
    """(max ulp distance, relative L2) between two float32 / complex64 arrays of the same shape."""
    a = np.ascontiguousarray(a); b = np.ascontiguousarray(b)
    assert a.shape == b.shape and a.dtype == b.dtype
    fa = a.view(np.float32).ravel() if a.dtype in (np.complex64, np.float32) else a.view(np.float64).ravel()
    fb = b.view(fa.dtype).ravel()
    it, mask = (np.int32, 0x7fffffff) if fa.dtype == np.float32 else (np.int64, 0x7fffffffffffffff)
    ia = fa.view(it).astype(np.int64); ib = fb.view(it).astype(np.int64)
    ia = np.where(ia < 0, -(ia & mask), ia)   # sign-magnitude float order -> monotone integer order
    ib = np.where(ib < 0, -(ib & mask), ib)
    ulp = int(np.abs(ia - ib).max()) if ia.size else 0
    den = float(np.linalg.norm(fa.astype(np.float64)))
    rel = float(np.linalg.norm(fa.astype(np.float64) - fb.astype(np.float64))) / den if den > 0 else 0.0
    return ulp, rel


def _packed_case(capi, protos, M, case):
    """Run one kernel's default and packed variant on the same ragged batch; returns {name: (default, packed)}.  Only the variable that
    selects the kernel under test is switched, so a difference can be attributed."""
    from distant_speech_recognition_b200 import synthetic
    U, n = 3, 9000
    utts = [synthetic.make_utterance(60 + u, 8, n) for u in range(U)]
    x = np.stack([t[0] for t in utts]); d = np.stack([t[1] for t in utts])
    lengths = np.array([n, n - 1234, 517], np.int32)
    h, g = protos[M]
    var = {"analysis": "BTKB_ANALYSIS_PACKED", "analysis_c3": "BTKB_ANALYSIS_PACKED", "synthesis": "BTKB_SYNTHESIS_PACKED"}.get(case, "BTKB_PERBIN_PACKED")
    res = {}
    try:
        for tag, env in (("default", "0"), ("packed", "1")):
            for v in ("BTKB_ANALYSIS_PACKED", "BTKB_SYNTHESIS_PACKED", "BTKB_PERBIN_PACKED"):
                os.environ[v] = "0"
            os.environ[var] = env
            out = {}
            if case in ("analysis", "synthesis", "nlms"):
                p = capi.Pipeline(8, M, 4, 1, beamformer=capi.BF_GSC_LMS, lms=dict(min_frames=8), max_utterances=U, max_samples=n)
                p.set_prototypes(h, g); p.set_delays(d); p.submit(x, lengths); p.run(True)
                if case == "analysis": out["X"] = p.fetch_snapshots()
                if case == "nlms": out["Y"] = p.fetch_subband(); out["wa"] = p.get_active_weights()
                if case == "synthesis": out["time"] = p.fetch_time()
                p.close()
            elif case == "analysis_c3":       # odd channel count: the unpaired last channel
                p = capi.Pipeline(3, M, 4, 1, beamformer=capi.BF_DS, max_utterances=U, max_samples=n)
                p.set_prototypes(h, g); p.submit(np.ascontiguousarray(x[:, :3]), lengths); p.run_analysis()
                out["X"] = p.fetch_snapshots(); p.close()
            elif case == "zelinski":          # static GSC + Zelinski (packed CSD recursions)
                p = capi.Pipeline(8, M, 4, 1, beamformer=capi.BF_GSC, postfilter=capi.PF_ZELINSKI, pf_alpha=0.7, pf_type=2, max_utterances=U, max_samples=n)
                p.set_prototypes(h, g); p.set_delays(d); p.submit(x, lengths); p.run(False)
                out["Y"] = p.fetch_subband(); out["W"] = p.get_postfilter_weights(); p.close()
            elif case in ("mccowan", "lefkimmiatis"):
                mpos = np.stack([40.0 * (np.arange(8) - 3.5), np.zeros(8), np.zeros(8)], axis=1)
                pfk = capi.PF_MCCOWAN if case == "mccowan" else capi.PF_LEFKIMMIATIS
                p = capi.Pipeline(8, M, 4, 1, beamformer=capi.BF_DS, postfilter=pfk, pf_alpha=0.7, pf_type=2, max_utterances=U, max_samples=n)
                p.set_prototypes(h, g); p.set_delays(d); p.pf_set_diffuse_noise_model(mpos, 16000.0); p.pf_set_diagonal_loading(0.05)
                p.submit(x, lengths); p.run(False)
                out["Y"] = p.fetch_subband(); out["W"] = p.get_postfilter_weights(); p.close()
            elif case == "smi_covariance":
                p = capi.Pipeline(8, M, 4, 1, beamformer=capi.BF_MVDR, max_utterances=U, max_samples=n)
                p.set_prototypes(h, g); p.set_delays(d); p.submit(x, lengths); p.run_analysis()
                p.accumulate_covariance(labels=np.tile([0.1, 0.3], (U, 1)), energy_threshold=10.0)
                out["R"] = p.get_covariance(); p.close()
            elif case == "wide_c16":          # the lane-split kernel for wide arrays
                x16c, d16c = synthetic.make_batch(2, 16, 6000, first=90)
                p = capi.Pipeline(16, M, 4, 1, beamformer=capi.BF_GSC_LMS, lms=dict(min_frames=8), max_utterances=2, max_samples=6000)
                p.set_prototypes(h, g); p.set_delays(d16c); p.submit(x16c, np.array([6000, 4100], np.int32)); p.run(False)
                out["Y"] = p.fetch_subband(); p.close()
            elif case in ("rls", "rls_constrained"):
                rls = dict(min_frames=8) if case == "rls" else dict(min_frames=8, regularization_param=1.0e-2, constraint_option=3, alpha2=1.0e-3)
                p = capi.Pipeline(8, M, 4, 1, beamformer=capi.BF_GSC_RLS, rls=rls, max_utterances=U, max_samples=n)
                p.set_prototypes(h, g); p.set_delays(d); p.submit(x, lengths); p.run(False)
                out["Y"] = p.fetch_subband(); out["wa"] = p.get_active_weights(); p.close()
            else:
                raise AssertionError(case)
            res[tag] = out
    finally:
        for v in ("BTKB_ANALYSIS_PACKED", "BTKB_SYNTHESIS_PACKED", "BTKB_PERBIN_PACKED"):
            os.environ.pop(v, None)
    return {k: (res["default"][k], res["packed"][k]) for k in res["default"]}


PACKED_CASES = [("analysis", 256), ("analysis", 512), ("analysis", 1024), ("analysis_c3", 512), ("synthesis", 256), ("synthesis", 512),
                ("synthesis", 1024), ("nlms", 512), ("nlms", 256), ("zelinski", 512), ("mccowan", 512), ("lefkimmiatis", 512),
                ("smi_covariance", 512), ("wide_c16", 512), ("rls", 512), ("rls_constrained", 512)]


@pytest.mark.gpu
@pytest.mark.parametrize("case,M", PACKED_CASES)
def test_packed_fp32_kernel_equals_the_scalar_kernel(capi, protos, case, M):
    """The FADD2 / FMUL2 / FFMA2 variants (csrc/btkb_f2.cuh, one switch per kernel family: BTKB_ANALYSIS_PACKED / BTKB_SYNTHESIS_PACKED /
    BTKB_PERBIN_PACKED, 0 = scalar, 1 = packed) perform the same IEEE operations per component as the scalar kernels — including the
    multiply-adds nvcc contracts in the scalar source, which are spelled out as fmaf() on both sides (round 1's B200 run failed here:
    the scalar dft8 was contracted to FFMA by -fmad=true, the packed one was not) — so the outputs must be equal BIT FOR BIT.  The
    failure message carries max-ulp and relative L2 per output."""
    got = _packed_case(capi, protos, M, case)
    report = {k: _ulp_report(a, b) for k, (a, b) in got.items()}
    for k, (a, b) in got.items():
        assert np.abs(a.view(np.float32) if a.dtype != np.float64 and a.dtype != np.complex128 else a.view(np.float64)).max() > 0, (case, k, "all-zero output")
        assert np.array_equal(a.view(np.uint8), b.view(np.uint8)), "%s M=%d: packed != scalar, (max ulp, rel L2) per output: %r" % (case, M, report)


@pytest.mark.gpu
def test_mvdrgsc_upgrade_blocking_matrix_and_orthogonalizer(protos):
    """SubbandMVDRGSC::upgrade_blocking_matrix / blocking_matrix_output and SubbandOrthogonalizer (beamformer.cc:2674-2716, 2776-2806):
    after an upgrade the blocking matrix of the bins >= 1 is orthogonal to wq - wl (bin 0 keeps its matrix), active weights set
    afterwards go through it, and branch i of the blocking matrix streams b_i^H x."""
    from oracle import restate
    from distant_speech_recognition_b200.btk20.beamformer import SubbandOrthogonalizerPtr
    g = load_golden("mvdrsd_zelinski1_c4_m256"); h, _ = protos[256]; M, D, C, K = 256, 128, 4, 129
    afbs = _afbs(g["x"], h, M, D)
    X = np.stack([restate.analysis(g["x"][c], h, M, 4, 1) for c in range(C)], axis=1)
    rng = np.random.default_rng(11)
    wa1 = 0.05 * (rng.standard_normal((K, C - 1)) + 1j * rng.standard_normal((K, C - 1)))
    wa2 = 0.05 * (rng.standard_normal((K, C - 1)) + 1j * rng.standard_normal((K, C - 1)))
    wq = restate.calc_mainlobe(M, C, FS, g["delays"])
    R = restate.diffuse_noise_model(M, g["mpos"], FS)[:K].astype(complex)
    R[:, np.eye(C, dtype=bool)] += float(np.float32(g["mu"]))
    wm = np.zeros((M, C), complex); wm[:K] = restate.calc_mvdr_weights(R, wq[:K], single=False)

    def make():
        bf = SubbandMVDRGSCPtr(fftlen=M)
        for a in afbs:
            bf.set_channel(a)
        bf.calc_array_manifold_vectors(FS, g["delays"])
        bf.set_diffuse_noise_model(g["mpos"], FS, 343740.0)
        bf.set_all_diagonal_loading(float(g["mu"]))
        bf.calc_mvdr_weights(FS, 1e-8, True)
        assert bf.calc_blocking_matrix2()
        return bf

    def put(bf, wa):
        for k in range(K):
            bf.set_active_weights_f(k, np.stack([wa[k].real, wa[k].imag], axis=1).ravel())

    B0 = np.stack([restate.calc_blocking_matrix(wm[k]) for k in range(K)])
    wl1 = restate.active_to_wl(B0, wa1)
    B1 = np.stack([B0[0]] + [restate.calc_blocking_matrix(wm[k] - wl1[k]) for k in range(1, K)])   # fbinX = 1 .. (beamformer.cc:2681)
    # (a) set, upgrade: the output still uses the old wl; the branches come from the NEW matrix
    bf = make(); put(bf, wa1); bf.upgrade_blocking_matrix()
    wl = np.zeros((M, C), complex); wl[:K] = wl1
    want_y = restate.subband_mvdr(X, wm, wl)
    # the orthogonalizer with outChanX = 0 pulls the beamformer; the branch streams read the SAME frame (they never advance it,
    # beamformer.cc:2793-2798), so they are driven frame by frame next to it
    so0 = SubbandOrthogonalizerPtr(bf, outChanX=0)
    branches = {i: SubbandOrthogonalizerPtr(bf, outChanX=i + 1) for i in (0, C - 2)}
    Y, Z = [], {i: [] for i in branches}
    for t, y in enumerate(so0):
        Y.append(np.array(y))
        for i, so in branches.items():
            Z[i].append(np.array(so.next(t)))
    Y = np.array(Y)
    assert Y.shape == (X.shape[0], M) and rel_l2(Y[:, :K], want_y[:, :K]) < TOL
    for i in branches:
        Zi = np.array(Z[i])
        want = np.einsum("kc,tck->tk", np.conj(B1[:, :, i]), X[:, :, :K])
        assert Zi.shape == (X.shape[0], M) and rel_l2(Zi[:, :K], want) < TOL, i
        assert rel_l2(Zi[:, K:], want_y[:, K:]) < TOL          # the reference's shared output vector keeps the beamformer's upper half
    # (b) set, upgrade, set again: wl = B1 wa2
    bf = make(); put(bf, wa1); bf.upgrade_blocking_matrix(); put(bf, wa2)
    wl = np.zeros((M, C), complex); wl[:K] = restate.active_to_wl(B1, wa2)
    Y = np.array([np.array(v) for v in bf])
    assert rel_l2(Y[:, :K], restate.subband_mvdr(X, wm, wl)[:, :K]) < TOL
    wl_old = np.zeros((M, C), complex); wl_old[:K] = restate.active_to_wl(B0, wa2)
    assert rel_l2(Y[:, 1:K], restate.subband_mvdr(X, wm, wl_old)[:, 1:K]) > 1e-3      # and that differs from the matrix before the upgrade
    # (c) without active weights the first branch is b_0^H x of calc_blocking_matrix2's matrix, bin 0 included
    bf = make()
    Z = []
    for t, y in enumerate(bf):
        Z.append(np.array(bf.blocking_matrix_output(0)))
    assert rel_l2(np.array(Z)[:, :K], np.einsum("kc,tck->tk", np.conj(B0[:, :, 0]), X[:, :, :K])) < TOL


@pytest.mark.gpu
def test_synthesis_bank_input_source_vector(protos):
    """OverSampledDFTSynthesisBank::input_source_vector (modulated/modulated.h:330): frames pushed before the first next() enter the
    buffer ahead of the source's; the outputs are those of the concatenated sequence from block (number pushed) on (r = 0)."""
    from oracle import restate
    from distant_speech_recognition_b200.btk20.modulated import OverSampledDFTSynthesisBankPtr
    from distant_speech_recognition_b200.btk20.stream import PyVectorComplexFeatureStreamPtr
    M, m, r = 256, 2, 0
    rng = np.random.default_rng(5)
    gproto = rng.standard_normal(M * m) / M
    T, P = 12, 3
    Y = rng.standard_normal((T + P, M // 2 + 1)) + 1j * rng.standard_normal((T + P, M // 2 + 1))
    Y[:, 0] = Y[:, 0].real; Y[:, -1] = Y[:, -1].real
    full = restate._hermitian_fill(Y, M)

    class Src:
        def size(self):
            return M
        def __iter__(self):
            return iter(full[P:])
        def reset(self):
            pass
    syn = OverSampledDFTSynthesisBankPtr(PyVectorComplexFeatureStreamPtr(Src()), prototype=gproto, M=M, m=m, r=r, delay_compensation_type=2)
    for t in range(P):
        syn.input_source_vector(full[t])
    got = np.concatenate([np.array(v) for v in syn])
    want = restate.synthesis(full, gproto, M, m, r, 2).reshape(-1, M)[P:].ravel()
    assert got.shape == want.shape and rel_l2(got, want) < TOL
    with pytest.raises(Exception):
        OverSampledDFTSynthesisBankPtr(PyVectorComplexFeatureStreamPtr(Src()), prototype=gproto, M=M, m=m, r=1, delay_compensation_type=2).input_source_vector(full[0])


@pytest.mark.gpu
def test_zelinski_postfilter_on_a_foreign_stream(protos):
    """ZelinskiPostFilter wired WITHOUT a beamformer object: set_snapshot_array + set_array_manifold_vector per bin
    (postfilter/postfilter.cc:384-417), the output stream being any beamformer that shares the snapshot array.  Here: a Python stream
    that applies its own (non delay-and-sum) weights and updates the shared SnapShotArrayPtr frame by frame.  Expected: the filter gains
    from the snapshots and the manifold, times THAT stream's output (ZelinskiFilter, postfilter.cc:57-219); also C-ABI level:
    btkb_set_snapshots."""
    from oracle import restate
    from distant_speech_recognition_b200.btk20.beamformer import SnapShotArrayPtr
    from distant_speech_recognition_b200.btk20.postfilter import ZelinskiPostFilterPtr
    from distant_speech_recognition_b200.btk20.stream import PyVectorComplexFeatureStreamPtr
    g = load_golden("mvdrsd_zelinski1_c4_m256"); h, _ = protos[256]; M, C, K = 256, 4, 129
    X = np.stack([restate.analysis(g["x"][c], h, M, 4, 1) for c in range(C)], axis=1)[:40]          # [T][C][M]
    T = X.shape[0]
    rng = np.random.default_rng(21)
    wf = np.zeros((M, C), complex)
    wf[:K] = (rng.standard_normal((K, C)) + 1j * rng.standard_normal((K, C))) / C
    Yf = restate.subband_ds(X, wf)                                                                   # the foreign beamformer's output
    wq = restate.calc_mainlobe(M, C, FS, g["delays"])
    snap = SnapShotArrayPtr(M, C)

    class Foreign:
        def size(self):
            return M
        def __iter__(self):
            for t in range(T):
                for c in range(C):
                    snap.set_samples(X[t, c], c)
                snap.update()
                yield Yf[t]
        def reset(self):
            pass
    for alpha, pf_type, min_frames in ((0.6, 2, 0), (0.7, 1, 3)):
        pf = ZelinskiPostFilterPtr(PyVectorComplexFeatureStreamPtr(Foreign()), M, alpha, pf_type, min_frames)
        pf.set_snapshot_array(snap)
        for k in range(M):
            pf.set_array_manifold_vector(k, wq[k], False, 1)
        out = np.array([np.array(v) for v in pf])
        want, gains = restate.zelinski_postfilter(Yf, X, wq, alpha, pf_type, min_frames)
        assert out.shape == want.shape and rel_l2(out[:, :K], want[:, :K]) < TOL, (alpha, pf_type)
        assert rel_l2(out[min_frames + 1:, K:], want[min_frames + 1:, K:]) < TOL
        assert rel_l2(want[:, :K], Yf[:, :K]) > 1e-2                                                 # the filter does something
    with pytest.raises(Exception):                                                                   # neither wiring: "set beamformer's weights"
        next(iter(ZelinskiPostFilterPtr(PyVectorComplexFeatureStreamPtr(Foreign()), M, 0.6, 2, 0)))
