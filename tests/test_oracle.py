"""CPU tests (-m "not gpu"): pin the oracle.

(1) oracle/restate.py (fp64 NumPy restatement) against golden vectors produced by the reference's OWN C++
    (oracle/_ref/libbtkref.so = btk20_src/{stream,modulated,beamformer,postfilter}.cc compiled unmodified against
    oracle/gsl_shim; generator: tests/golden/make_golden.py).
(2) the compiled reference itself against the same goldens when oracle/_ref is present (it travels to the GPU box).
(3) known-answer properties: B^H B = I, B^T v = 0, w^H v = 1, Hermitian symmetry, frame count, the
    shipped M=256 prototypes reproduced by the reference's design tool.
"""
import os
import numpy as np
import pytest

from conftest import GOLDEN, load_golden, rel_l2
from oracle import restate

FS = 16000.0


def _X(x, h, M):
    return np.stack([restate.analysis(x[c], h, M, 4, 1) for c in range(x.shape[0])], axis=1)


def test_design_tool_reproduces_shipped_prototypes():
    a = np.load(os.path.join(GOLDEN, "prototype_M256_m4_r1.npz"))
    b = np.load(os.path.join(GOLDEN, "prototype_shipped_M256_m4_r1.npz"))
    assert np.abs(a["h"] - b["h"]).max() < 1e-12
    assert np.abs(a["g"] - b["g"]).max() < 1e-11


def test_frame_count_closed_form():
    # T = ceil(n/D) + m R / 2 for delay-compensation type 2; cfg sizes of BASELINE.md
    assert restate.num_frames(160000, 256, 4, 1, 2) == 1254
    assert restate.num_frames(80000, 512, 4, 1, 2) == 317
    assert restate.num_frames(80000, 1024, 4, 1, 2) == 161
    assert restate.num_frames(16000, 256, 4, 1, 1) == 125 + 7
    assert restate.num_frames(16000, 256, 3, 0, 0) == 63 + 5


def test_analysis_synthesis_golden_ds(protos):
    g = load_golden("ds_c2_m256"); h, gg = protos[256]
    X = _X(g["x"], h, 256)
    assert rel_l2(X[:, 0, :129], g["X0"]) < 1e-13
    # Hermitian symmetry of the analysis output (real input)
    assert np.abs(X[:, 0, 1:128] - np.conj(X[:, 0, 255:128:-1])).max() < 1e-7 * np.abs(X).max()
    wq = restate.calc_mainlobe(256, 2, FS, g["delays"])
    assert rel_l2(wq[:129], g["w"]) < 1e-14
    Y = restate.subband_ds(X, wq)
    assert rel_l2(Y[:, :129], g["Y"]) < 1e-13
    y = restate.synthesis(Y, gg, 256, 4, 1)
    assert np.array_equal(y, g["time"])  # bit-exact (float32 accumulation order reproduced)


def test_gsc_lms_golden(protos):
    g = load_golden("gsclms_c8_m512"); h, gg = protos[512]
    X = _X(g["x"], h, 512)
    Y, waH, nu = restate.gsc_lms(X, FS, g["delays"], min_frames=int(g["min_frames"]))
    assert rel_l2(Y[:, :257], g["Y"]) < 1e-12
    assert rel_l2(waH, g["waH"]) < 1e-11
    assert nu == g["stats"][2]
    assert rel_l2(restate.synthesis(Y, gg, 512, 4, 1), g["time"]) < 1e-7


def test_projector_form_equals_b_form(protos):
    """carrying u = waH B^T (O(C)) is the same recurrence as the reference's B-form."""
    g = load_golden("gsclms_c8_m512"); h, _ = protos[512]
    X = _X(g["x"], h, 512)
    T, C, M = X.shape; K = M // 2 + 1
    p = dict(restate.DEFAULT_LMS); p["min_frames"] = 10
    vs = np.stack([restate.calc_array_manifold_f(f, M, FS, g["delays"]) for f in range(K)])
    u = np.zeros((K, C), complex); se = np.full(K, p["init_diagonal_load"]); E = p["init_diagonal_load"]; gamma = p["gamma"]
    Yk = np.zeros((T, K), complex)
    for t in range(T):
        energy = abs(np.vdot(X[t, 0], X[t, 0])) / M
        adapt = energy > E / p["sil_thresh"]
        XK = X[t, :, :K].T
        Yc = (np.conj(vs) * XK).sum(1)
        xx = (np.abs(XK) ** 2).sum(1)
        sub = np.maximum(se * p["beta"] + (1 - p["beta"]) * xx if t > 0 else xx, p["energy_floor"])
        if adapt:
            e = Yc - (u * XK).sum(1); a = gamma / sub
            q = XK - (C * Yc)[:, None] * vs
            un = u + (e * a)[:, None] * np.conj(q) - (a * p["regularization_param"])[:, None] * u
            n2 = (np.abs(un) ** 2).sum(1)
            u = np.where((n2 > p["max_wa_l2norm"])[:, None], un * np.sqrt(p["max_wa_l2norm"] / np.maximum(n2, 1e-300))[:, None], un)
            se = sub
        Yk[t] = Yc - (u * XK).sum(1) if t >= p["min_frames"] else Yc
        E = E * p["beta"] + (1 - p["beta"]) * energy
    assert rel_l2(Yk, g["Y"]) < 1e-10


def test_gsc_static_zelinski_golden(protos):
    g = load_golden("gsc_zelinski_c8_m512"); h, gg = protos[512]
    X = _X(g["x"], h, 512); M = 512; K = 257
    wq = restate.calc_mainlobe(M, 8, FS, g["delays"])
    assert rel_l2(wq[:K], g["wq"]) < 1e-14
    B = np.stack([restate.calc_blocking_matrix(wq[f]) for f in range(K)])
    assert rel_l2(B, g["B"]) < 1e-12
    # known answers: orthonormal columns, B^T v = 0 (NOT B^H v), B B^H = P
    for f in (1, 17, 256):
        assert np.abs(B[f].conj().T @ B[f] - np.eye(7)).max() < 1e-12
        assert np.abs(B[f].T @ wq[f]).max() < 1e-14
    wl = np.zeros((M, 8), complex); wl[:K] = restate.active_to_wl(B, g["wa"])
    Y, W = restate.zelinski_postfilter(restate.subband_gsc(X, wq, wl), X, wq, 0.7, 2, 0)
    assert rel_l2(Y[:, :K], g["Y"]) < 1e-12
    assert W.min() >= 1e-4 and W.max() <= 1.0
    assert rel_l2(restate.synthesis(Y, gg, M, 4, 1), g["time"]) < 1e-7


def test_smi_mvdr_golden(protos):
    g = load_golden("smimvdr_zelinski_c8_m512"); h, gg = protos[512]
    X = _X(g["x"], h, 512); M = 512; K = 257
    R, nf = restate.smi_covariance(X, FS, 256, ((0.25, 0.75),), 10.0)
    assert nf > 5
    assert rel_l2(R, g["cov"]) < 1e-13
    wq = restate.calc_mainlobe(M, 8, FS, g["delays"])
    w = restate.calc_mvdr_weights(R + float(np.float32(g["mu"])) * np.eye(8), wq, single=False)
    # distortionless: w^H v = 1 with v = C * wq (unit-modulus manifold)
    for f in (1, 100, 256):
        assert abs(np.vdot(w[f], wq[f] * 8) - 1.0) < 1e-9
    # the reference inverts in complex<float> (beamformer.cc:237-253): on this ill-conditioned sample covariance its
    # weights sit 5.3e-4 (relative L2) from the fp64 solution -- that is the reference's own rounding, not the oracle's
    assert rel_l2(w, g["w"]) < 1e-3
    Y, _ = restate.zelinski_postfilter(restate.subband_mvdr(X, g["w"]), X, wq, 0.7, 2, 0)
    assert rel_l2(Y[:, :K], g["Y"]) < 1e-12   # with the reference's own weights the rest of the chain is exact


def test_mvdr_superdirective_golden(protos):
    g = load_golden("mvdrsd_zelinski1_c4_m256"); h, gg = protos[256]
    X = _X(g["x"], h, 256); M = 256; K = 129
    wq = restate.calc_mainlobe(M, 4, FS, g["delays"])
    R = restate.diffuse_noise_model(M, g["mpos"], FS) + float(np.float32(g["mu"])) * np.eye(4)
    w = restate.calc_mvdr_weights(R, wq, single=False)
    assert rel_l2(w, g["w"]) < 5e-5
    Y, _ = restate.zelinski_postfilter(restate.subband_mvdr(X, g["w"]), X, wq, 0.6, 1, 5)
    assert rel_l2(Y[:, :K], g["Y"]) < 1e-12


def _coherence(M, mpos, load):
    R = restate.diffuse_noise_model(M, mpos, FS)
    return R + float(np.float32(load)) * np.eye(R.shape[1])


def test_mccowan_postfilter_golden(protos):
    """McCowanPostFilter behind the super-directive MVDR (postfilter.cc:496-934): restatement vs the reference's output."""
    g = load_golden("mccowan_c4_m256"); h, gg = protos[256]
    M = 256; K = 129
    X = _X(g["x"], h, M)
    wq = restate.calc_mainlobe(M, 4, FS, g["delays"])
    Ybf = restate.subband_mvdr(X, g["w"])
    Ya, Wa = restate.mccowan_postfilter(Ybf, X, wq, _coherence(M, g["mpos"], 0.01), 0.7, 2, 0, 0.99)
    assert rel_l2(Ya[:, :K], g["Ya"]) < 1e-12
    assert np.array_equal(Ya[:2, K:] == 0, g["upper_a"] == 0)  # frame 0 leaves the upper half of vector_ at zero
    assert rel_l2(restate.synthesis(Ya, gg, M, 4, 1)[: len(g["timea"])], g["timea"]) < 1e-6
    Yb, _ = restate.mccowan_postfilter(Ybf, X, wq, _coherence(M, g["mpos"], 0.0), 0.6, 1, 3, 0.9)
    assert rel_l2(Yb[:, :K], g["Yb"]) < 1e-12
    assert np.array_equal(Yb[:5, K:] == 0, g["upper_b"] == 0)
    assert Wa.min() >= 1e-4 and Wa.max() <= 1.0


def test_lefkimmiatis_postfilter_golden(protos):
    """LefkimmiatisPostFilter behind delay-and-sum (postfilter.cc:935-1200); the coherence inverse is the reference's
    float LINPACK SVD, so agreement is at single precision."""
    g = load_golden("lefkimmiatis_c8_m512"); h, gg = protos[512]
    M = 512; K = 257
    X = _X(g["x"], h, M)
    wq = restate.calc_mainlobe(M, 8, FS, g["delays"])
    Ybf = restate.subband_ds(X, wq)
    Ya, _ = restate.lefkimmiatis_postfilter(Ybf, X, wq, _coherence(M, g["mpos"], 0.1), 0.8, 2, 0, 0.99, 1e-4, 100, single=False)
    assert rel_l2(Ya[:, :K], g["Ya"]) < 2e-6
    Yb, _ = restate.lefkimmiatis_postfilter(Ybf, X, wq, _coherence(M, g["mpos"], 0.01), 0.6, 1, 2, 0.99, 1e-8, 0, single=False)
    assert rel_l2(Yb[:, :K], g["Yb"]) < 2e-5
    assert rel_l2(restate.synthesis(Yb, gg, M, 4, 1)[: len(g["timeb"])], g["timeb"]) < 2e-5


def test_gsc_lms_golden_is_the_reference_pythons_output():
    """golden_gsclms_c8_m512 came from the C++ re-statement in oracle/ref_harness.cc; golden_pyref_gsclms_c8_m512 from the
    reference's own SubbandGSCLMSBeamformer.__iter__ run through oracle/pyref.py on the same snapshots: they must agree."""
    a = load_golden("gsclms_c8_m512"); b = load_golden("pyref_gsclms_c8_m512")
    assert rel_l2(a["Y"], b["Y"]) < 1e-13 and rel_l2(a["waH"], b["waH"]) < 1e-12
    assert int(a["stats"][2]) == int(b["n_updates"])


def test_gsc_rls_golden(protos):
    """SubbandGSCRLSBeamformer (pybeamformer.py:765-928): B-form and projector-form restatements vs the reference Python's output."""
    for name, M in (("gscrls_c8_m512", 512), ("gscrls_c4_m256", 256)):
        g = load_golden(name); h, gg = protos[M]; K = M // 2 + 1
        X = _X(g["x"], h, M)
        kw = {k: (int(g[k]) if k in ("min_frames", "constraint_option") else float(g[k])) for k in restate.DEFAULT_RLS if k in g.files}
        Y, waH, nu = restate.gsc_rls(X, FS, g["delays"], **kw)
        assert rel_l2(Y[:, :K], g["Y"]) < 1e-11 and rel_l2(waH, g["waH"]) < 1e-10 and nu == int(g["n_updates"])
        Yp, u, nup = restate.gsc_rls_projector(X, FS, g["delays"], **kw)
        assert rel_l2(Yp[:, :K], g["Y"]) < 1e-11 and nup == nu
        B = np.stack([restate.calc_blocking_matrix(restate.calc_array_manifold_f(f, M, FS, g["delays"]), 1) for f in range(K)])
        assert rel_l2(np.einsum("kc,kci->ki", u, np.conj(B)), g["waH"]) < 1e-10      # waH = u conj(B)
        assert rel_l2(restate.synthesis(Y, gg, M, 4, 1)[: len(g["time"])], g["time"]) < 1e-6


def test_reference_python_runs_live_when_present(protos):
    """With /root/reference mounted (build container), run the reference's own Python NLMS / RLS loops now."""
    from oracle import pyref
    if not pyref.available():
        pytest.skip("reference not mounted (GPU box)")
    g = load_golden("gscrls_c4_m256"); h, _ = protos[256]
    X = _X(g["x"], h, 256)
    Y, waH, nu = pyref.run_adaptive("lms", X, FS, g["delays"], 128, min_frames=4)
    Yo, wo, no = restate.gsc_lms(X, FS, g["delays"], min_frames=4)
    assert rel_l2(Yo, Y) < 1e-13 and no == nu
    Y, waH, nu = pyref.run_adaptive("rls", X, FS, g["delays"], 128, min_frames=4, constraint_option=1, alpha2=0.02)
    Yo, wo, no = restate.gsc_rls(X, FS, g["delays"], min_frames=4, constraint_option=1, alpha2=0.02)
    assert rel_l2(Yo, Y) < 1e-12 and no == nu


WPE_A = dict(lower_num=0, upper_num=5, iterations_num=2, load_db=-18.0, band_width=0.0, diagonal_bias=1e-4)
WPE_B = dict(lower_num=2, upper_num=8, iterations_num=3, load_db=-20.0, band_width=3000.0, diagonal_bias=1e-3, start_frame_no=2, end_frame_no=42)
WPE_C = dict(lower_num=1, upper_num=5, iterations_num=2, load_db=-40.0, band_width=0.0, diagonal_bias=1e-4)
WPE_8 = dict(lower_num=1, upper_num=8, iterations_num=2, load_db=-35.0, band_width=0.0, diagonal_bias=1e-4)


def test_wpe_golden(protos):
    """MultiChannelWPEDereverberation (dereverberation.cc:312-733): restatement vs the compiled reference's output, incl. the
    output stage's truncated lag window (lower_num > 0), band_width and estimate_filter(start, end)."""
    g = load_golden("wpe_c4_m256"); h, _ = protos[256]
    X = _X(g["x"], h, 256)
    for tag, kw in (("a", WPE_A), ("b", WPE_B), ("c", WPE_C)):
        Xo, G, used = restate.wpe(X, samplerate=FS, **kw)
        assert rel_l2(Xo[:, :, :129], g["X" + tag]) < 1e-11, tag
        assert np.allclose(Xo[:, :, 129:], np.conj(Xo[:, :, 1:128][:, :, ::-1]))
    assert used == 67 and int(g["used_b"]) == 40
    assert rel_l2(g["Xb"], X[:, :, :129]) > 0.1 and rel_l2(g["Xc"], X[:, :, :129]) > 0.03   # the filters really remove something
    g = load_golden("wpe_c8_m512"); h, _ = protos[512]
    X = _X(g["x"], h, 512)
    Xo, G, used = restate.wpe(X, samplerate=FS, **WPE_8)
    assert rel_l2(Xo[:, :, :257], g["Xa"]) < 1e-11


def test_pseudoinverse_golden():
    g = load_golden("pseudoinverse")
    for A, inv in zip(g["A"], g["inv"]):
        mine, ok = restate.pseudoinverse(A, 1e-8, single=False)
        assert ok and rel_l2(mine, inv) < 1e-4     # LINPACK float SVD vs fp64
        assert rel_l2(inv @ A, np.eye(8)) < 1e-3


def test_delays_known_answers():
    mpos = np.array([[-113.0, 0, 2], [36.0, 0, 2], [76.0, 0, 2], [113.0, 0, 2]])  # unit_test/confs/ds.json
    d = restate.calc_la_delays(mpos, -1.306379)
    assert d[2] == 0.0 and abs(d[0] - (113.0 + 76.0) * np.cos(-1.306379) / 343740.0) < 1e-15
    d2 = restate.calc_delays("linear", mpos, [-1.306379, None, None])
    assert np.array_equal(d, d2)
    dn = restate.calc_nf_delays(mpos, 0.0, 1000.0, 0.0)
    assert dn[2] == 0.0 and dn[0] > 0


def test_spectral_matrix_update_legacy_noconj():
    rng = np.random.default_rng(1)
    x = rng.standard_normal((3, 4)) + 1j * rng.standard_normal((3, 4))
    R = restate.spectral_matrix_update(np.zeros((3, 4, 4), complex), x, 0.95)
    assert np.allclose(R[1], 0.05 * np.outer(x[1], x[1]))  # x x^T, no conjugate (beamformer.cc:131-139)


@pytest.mark.skipif(not os.path.exists(os.path.join(os.path.dirname(GOLDEN), "..", "oracle", "_ref", "libbtkref.so")), reason="oracle/_ref not built")
def test_compiled_reference_matches_its_goldens(protos):
    from oracle import ref
    g = load_golden("gsclms_c8_m512"); h, gg = protos[512]
    res = ref.beamform(g["x"], h, gg, g["delays"], 512, 4, 1, bf_kind=ref.BF_GSC_LMS, lms=dict(min_frames=int(g["min_frames"])))
    assert np.array_equal(res["Y"][:, :257], g["Y"]) and np.array_equal(res["time"], g["time"])
    g = load_golden("mccowan_c4_m256"); h, gg = protos[256]
    res = ref.beamform(g["x"], h, gg, g["delays"], 256, 4, 1, bf_kind=ref.BF_MVDR_SD, mpos=g["mpos"], mvdr_mu=0.01,
                       pf=dict(kind="mccowan", alpha=0.7, type=2, diag_load=0.01))
    assert np.array_equal(res["Y"][:, :129], g["Ya"]) and np.array_equal(res["time"], g["timea"])
    g = load_golden("ds_c2_m256"); h, gg = protos[256]
    res = ref.beamform(g["x"], h, gg, g["delays"], 256, 4, 1, bf_kind=ref.BF_DS)
    assert np.array_equal(res["Y"][:, :129], g["Y"]) and np.array_equal(res["time"], g["time"])


SOS_GOLDENS = (("bmvdr_vad_c8_m512", 512, "bmvdr"), ("bmvdr_tfmask_c4_m256", 256, "bmvdr"), ("gev_vad_c8_m512", 512, "gev"), ("gev_tfmask_c4_m256", 256, "gev"))


@pytest.mark.parametrize("name,M,kind", SOS_GOLDENS)
def test_sos_batch_beamformer_goldens(protos, name, M, kind):
    """SubbandBlindMVDRBeamformer / SubbandGEVBeamformer (pybeamformer.py:1026-1357): restatement vs the reference Python's own
    output (tests/golden/make_golden_sos.py), incl. the label walk that drops an open-ended second segment, fractional TF masks
    with the integer-count truncation, and the GEV eigenvector up to ONE global sign."""
    g = load_golden(name); h, gg = protos[M]; K = M // 2 + 1
    X = _X(g["x"], h, M)
    labels = [tuple(r) for r in g["labels"]] if "labels" in g.files else None
    mt = g["mask_t"] if "mask_t" in g.files else None
    mj = g["mask_j"] if "mask_j" in g.files else None
    Rt, Rn, ct, cn = restate.sos_accumulate(X, FS, M // 2, target_labs=labels, mask_t=mt, mask_j=mj, energy_threshold=float(g["energy_threshold"]))
    assert np.array_equal(ct, g["ct"]) and np.array_equal(cn, g["cn"])
    if kind == "bmvdr":
        w = restate.sos_bmvdr_weights(Rt, Rn, ct, cn, gamma=float(g["gamma"]), ref_micx=int(g["ref_micx"]), offset=float(g["offset"]))
    else:
        w = restate.sos_gev_weights(Rt, Rn, cn, gamma=float(g["gamma"]))
        w2 = restate.sos_gev_weights(Rt, Rn, cn, gamma=float(g["gamma"]), use_jacobi=False)
        assert rel_l2(w2, w) < 1e-9                       # the Jacobi solver the CUDA kernel uses == LAPACK
        w = w * np.sign(np.real(np.vdot(w[0], g["w"][0])))
    assert rel_l2(w, g["w"]) < 1e-9
    Y = restate.sos_apply(X, w)
    assert rel_l2(Y[:, :K], g["Y"]) < 1e-6               # goldens are stored as complex64
    assert rel_l2(restate.synthesis(Y, gg, M, 4, 1)[: len(g["time"])], g["time"]) < 1e-6
    if mt is not None and name.startswith("bmvdr"):
        assert np.any(mt != np.round(mt)) and ct.max() < np.sum(mt > 0, axis=0).max() * 1.7  # fractional masks were exercised


def test_sos_label_walk_quirk():
    """pybeamformer.py:1086-1091: an open-ended segment (end < 0) that does not start at 0 is skipped by the cursor on the first
    earlier frame, so the default target_labs=[(0.1, -1)] never marks a target frame; a closed segment works."""
    e = np.full(50, 100.0)
    wt, wn = restate.sos_label_weights(50, e, FS, 256, [(0.1, -1)], 10.0)
    assert wt.sum() == 0 and wn.sum() == 50
    wt, wn = restate.sos_label_weights(50, e, FS, 256, [(0.0, -1)], 10.0)
    assert wt.sum() == 50
    wt, wn = restate.sos_label_weights(50, e, FS, 256, [(0.1, 0.3), (0.5, 0.6)], 10.0)
    t = np.arange(50) * 256 / FS
    assert wt.sum() == np.sum((t >= 0.1) & (t <= 0.3)) + np.sum((t >= 0.5) & (t <= 0.6))


def test_sos_reference_python_runs_live_when_present(protos):
    from oracle import pyref
    if not pyref.available():
        pytest.skip("reference not mounted (GPU box)")
    g = load_golden("gev_tfmask_c4_m256"); h, _ = protos[256]
    X = _X(g["x"], h, 256)
    res = pyref.run_sos("gev", X, FS, 128, mask_t=g["mask_t"], mask_j=g["mask_j"], energy_threshold=10, gamma=float(g["gamma"]))
    assert rel_l2(np.conj(res["wqH"]), g["w"]) < 1e-12


def test_wpe_single_channel_golden(protos):
    """SingleChannelWPEDereverberationFeature (dereverberation.cc:24-310) = the multi-channel estimator with one channel and no
    diagonal bias: the restatement with C = 1, diagonal_bias = 0 reproduces the compiled reference's single-channel output."""
    g = load_golden("wpe_single_m256"); h, gg = protos[256]
    X = restate.analysis(g["x"][0], h, 256, 4, 1)[:, None, :]
    Xa, _, ua = restate.wpe(X, lower_num=0, upper_num=16, iterations_num=2, load_db=-20.0, band_width=0.0, diagonal_bias=0.0, samplerate=FS)
    Xb, _, ub = restate.wpe(X, lower_num=2, upper_num=12, iterations_num=3, load_db=-25.0, band_width=3000.0, diagonal_bias=0.0, samplerate=FS,
                            start_frame_no=2, end_frame_no=42)
    assert rel_l2(Xa[:, 0, :129], g["Xa"]) < 1e-11 and rel_l2(Xb[:, 0, :129], g["Xb"]) < 1e-11
    assert ua == int(g["used_a"]) and ub == int(g["used_b"]) == 40
    assert rel_l2(restate.synthesis(Xa[:, 0, :], gg, 256, 4, 1), g["time_a"]) < 1e-6


def test_gsc_rls_cpp_golden(protos):
    """The reference's C++ SubbandGSCRLS (beamformer.cc:1447-1699), run through the compiled reference: the B-form restatement
    follows it to 1e-9.  At the class defaults (init_precision_matrix(0.01) on int16-scale spectra) the recursion cancels 14
    digits in every update (Pz - gz Z^H Pz with Z^H Pz Z ~ 1e14 mu), so even in fp64 an algebraically equal form (the
    blocking-matrix-free projector form) lands 1e-4 away: the reason this class has no fp32 CUDA kernel."""
    CASES = (dict(mu=0.9, sigma2=0.01, init_sigma2=0.01), dict(mu=0.97, sigma2=0.0, init_sigma2=1e6, alpha=0.5, qctype=2),
             dict(mu=0.95, sigma2=1e-3, init_sigma2=1.0, alpha=0.3, qctype=1))   # tests/golden/make_golden_rls_cpp.py
    g = load_golden("gscrls_cpp_c4_m256"); h, _ = protos[256]
    X = _X(g["x"], h, 256)
    for i, kw in enumerate(CASES):
        Y, _ = restate.gsc_rls_cpp(X, FS, g["delays"], **kw)
        assert rel_l2(Y[:, :129], g["Y%d" % i]) < 1e-8, i
    Yp, _ = restate.gsc_rls_cpp(X, FS, g["delays"], projector=True, **CASES[0])
    assert 1e-6 < rel_l2(Yp[:, :129], g["Y0"]) < 1e-2          # equal algebra, 14 cancelled digits
    Yp, _ = restate.gsc_rls_cpp(X, FS, g["delays"], projector=True, **CASES[1])
    assert rel_l2(Yp[:, :129], g["Y1"]) < 1e-8                 # a well-scaled start (1/sigma2 = 1e-6) behaves


# ---------------------------------------------------------------------------------------------------------------------
# Known-answer / property checks that need no golden (last row)
def test_blocking_matrix_properties():
    """calc_blocking_matrix_ (beamformer.cc:373-454): orthonormal columns, B B^H = P = I - conj(v) v^T / ||v||^2, B^T v = 0 — for
    the delay-and-sum manifold of every bin and for NC = 2, 3 constraint sets."""
    rng = np.random.default_rng(5)
    C, M = 8, 512
    d = rng.uniform(-2e-4, 2e-4, C)
    wq = restate.calc_mainlobe(M, C, FS, d)
    for f in (1, 7, 100, 255, 256):
        v = wq[f]
        for Nc in (1, 2, 3):
            B = restate.calc_blocking_matrix(v, Nc)
            assert B.shape == (C, C - Nc)
            assert np.allclose(B.conj().T @ B, np.eye(C - Nc), atol=1e-12)
            assert np.abs(B.T @ v).max() < 1e-12
            if Nc == 1:
                P = np.eye(C) - np.outer(np.conj(v), v) / np.real(np.vdot(v, v))
                assert np.allclose(B @ B.conj().T, P, atol=1e-12)


def test_mvdr_is_distortionless_and_ds_returns_a_plane_wave(protos):
    """w^H v = 1 for the MVDR weights of any Hermitian positive-definite R (calc_mvdr_weights, beamformer.cc:2350-2402, with
    d = wq = v / C); delay-and-sum of a pure look-direction plane wave returns the source spectrum (SubbandDS::next)."""
    rng = np.random.default_rng(6)
    C, M = 8, 256; K = M // 2 + 1
    d = rng.uniform(-2e-4, 2e-4, C)
    wq = restate.calc_mainlobe(M, C, FS, d)
    A = rng.standard_normal((K, C, 3 * C)) + 1j * rng.standard_normal((K, C, 3 * C))
    R = A @ np.conj(np.transpose(A, (0, 2, 1))) + 0.1 * np.eye(C)
    w = restate.calc_mvdr_weights(R, wq, single=False)
    v = wq[:K] * C                                              # array manifold (unit modulus)
    assert np.abs(np.einsum("kc,kc->k", np.conj(w[1:K]), v[1:K]) - 1.0).max() < 1e-10
    assert np.allclose(w[0], 1.0)                               # the reference's DC quirk (beamformer.cc:2369-2371)
    s = rng.standard_normal((5, K)) + 1j * rng.standard_normal((5, K))
    X = np.zeros((5, C, M), complex)
    X[:, :, :K] = s[:, None, :] * v.T[None, :, :]
    Y = restate.subband_ds(X, wq)
    assert np.abs(Y[:, 1:K - 1] - s[:, 1:K - 1]).max() < 1e-12


def test_filterbank_round_trip_and_zelinski_gain_range(protos):
    """Analysis -> synthesis of one channel reproduces the (delayed) input to the prototype's design accuracy (interior error
    1.7e-3 with the shipped M = 256 pair); the Zelinski gain stays inside [1e-4, 1] (postfilter.cc:30-41)."""
    h, g = protos[256]; M, m, r = 256, 4, 1; D = M >> r
    rng = np.random.default_rng(7)
    x = (3000 * rng.standard_normal(6000)).astype(np.float32)
    X = restate.analysis(x, h, M, m, r)
    y = restate.synthesis(X, g, M, m, r)
    a, b = m * M, len(x) - m * M          # interior: delay compensation type 2 leaves no delay, the edges carry the filter transients
    assert np.linalg.norm(y[a:b] - x[a:b]) / np.linalg.norm(x[a:b]) < 3e-3
    xs = np.stack([x, np.roll(x, 3), x + 50 * rng.standard_normal(len(x)).astype(np.float32), 0.5 * x])
    Xs = np.stack([restate.analysis(xs[c], h, M, m, r) for c in range(4)], axis=1)
    wq = restate.calc_mainlobe(M, 4, FS, np.zeros(4))
    Yz, W = restate.zelinski_postfilter(restate.subband_ds(Xs, wq), Xs, wq, 0.6, 2, 0)
    assert W.min() >= 1e-4 - 1e-12 and W.max() <= 1.0 + 1e-12


def test_sos_on_the_references_own_fixtures():
    """The reference's own fixtures for this path (confs/{bmvdr,gev}_tfmask.json): the 4-channel Kinect recording, its TF-mask pickles
    and the shipped M = 256 prototypes (first 240 frames; golden_sos_kinect_c4_m256, tests/golden/make_golden_sos.py).  The
    restatement reproduces the reference Python's blind-MVDR and GEV outputs on them."""
    g = load_golden("sos_kinect_c4_m256")
    p = np.load(os.path.join(GOLDEN, "prototype_shipped_M256_m4_r1.npz"))
    M, K = 256, 129
    X = _X(g["x16"].astype(np.float32), p["h"], M)
    mt, mj = g["mask_t"].astype(np.float64), g["mask_j"].astype(np.float64)
    Rt, Rn, ct, cn = restate.sos_accumulate(X, FS, M // 2, mask_t=mt, mask_j=mj, energy_threshold=10.0)
    assert np.array_equal(ct, g["ct"]) and np.array_equal(cn, g["cn"])
    w = restate.sos_bmvdr_weights(Rt, Rn, ct, cn, gamma=1e-6, ref_micx=0, offset=0.0)
    assert rel_l2(w, g["w_bmvdr"]) < 1e-10
    Y = restate.sos_apply(X, w)
    assert rel_l2(Y[:, :K], g["Y_bmvdr"]) < 1e-6
    assert rel_l2(restate.synthesis(Y, p["g"], M, 4, 1), g["time_bmvdr"]) < 1e-6
    w = restate.sos_gev_weights(Rt, Rn, cn, gamma=1e-6)
    w = w * np.sign(np.real(np.vdot(w[0], g["w_gev"][0])))
    assert rel_l2(w, g["w_gev"]) < 1e-9
    assert rel_l2(restate.sos_apply(X, w)[:, :K], g["Y_gev"]) < 1e-6


def test_online_beamforming_on_the_references_own_fixtures():
    """unit_test/test_online_beamforming.py with its default inputs (4-channel Kinect recording, shipped M = 256 prototypes) and its
    own parameter files confs/{ds, ds_and_zelinski, sd, sd_and_zelinski, sd_and_mccowan, sd_and_lefkimmiatis, gsclms, gscrls}.json:
    the restatement against the reference's outputs (golden_online_kinect_c4_m256, tests/golden/make_golden_online_kinect.py)."""
    g = load_golden("online_kinect_c4_m256")
    p = np.load(os.path.join(GOLDEN, "prototype_shipped_M256_m4_r1.npz")); h, gg = p["h"], p["g"]
    M, K, C = 256, 129, 4
    d, mpos = g["delays"], g["mpos"]
    assert np.allclose(d, restate.calc_delays("linear", mpos, [-1.306379, None, None]), rtol=0, atol=1e-18)   # confs/*.json look direction
    s0, n = int(g["s0_static"]), int(g["n_static"])
    X = _X(g["x16"][:, s0:s0 + n].astype(np.float32), h, M)
    wq = restate.calc_mainlobe(M, C, FS, d)

    def check(name, Y, tol_y, tol_t):
        assert rel_l2(Y[:, :K], g["Y_" + name]) < tol_y, name
        t = restate.synthesis(Y, gg, M, 4, 1)[: len(g["time_" + name])]
        assert rel_l2(t, g["time_" + name]) < tol_t, name
        assert abs(float(np.inner(t, t)) / float(g["energy_" + name]) - 1.0) < 10 * tol_t, name    # the script's report: total_energy

    Yds = restate.subband_gsc(X, wq, np.zeros_like(wq))            # 'delay_and_sum' = SubbandGSCBeamformer with zero active weights
    check("ds", Yds, 1e-6, 1e-6)                                   # (goldens are stored as complex64 / float32)
    check("ds_and_zelinski", restate.zelinski_postfilter(Yds, X, wq, 0.7, 2, 0)[0], 1e-6, 1e-6)
    R = restate.diffuse_noise_model(M, mpos, FS)
    R[:, np.eye(C, dtype=bool)] += float(np.float32(0.01))
    wsd = restate.calc_mvdr_weights(R, wq, single=True)            # the reference's float LINPACK SVD
    assert rel_l2(wsd[1:], g["w_sd"][1:]) < 2e-5
    for name in ("sd", "sd_and_zelinski", "sd_and_mccowan", "sd_and_lefkimmiatis"):
        assert np.array_equal(g["w_" + name], g["w_sd"])           # every sd* file uses diagonal_load 0.01
    Ysd = restate.subband_mvdr(X, g["w_sd"])                       # downstream stages on the reference's own weights
    check("sd", Ysd, 1e-6, 1e-6)
    check("sd", restate.subband_mvdr(X, wsd), 3e-5, 3e-5)          # ... and end to end on the restated weights
    check("sd_and_zelinski", restate.zelinski_postfilter(Ysd, X, wq, 0.7, 2, 0)[0], 1e-6, 1e-6)
    check("sd_and_mccowan", restate.mccowan_postfilter(Ysd, X, wq, _coherence(M, mpos, 0.01), 0.7, 2, 0, 0.99)[0], 1e-6, 1e-6)
    check("sd_and_lefkimmiatis", restate.lefkimmiatis_postfilter(Ysd, X, wq, _coherence(M, mpos, 0.1), 0.8, 2, 0, 0.99, 1e-4, 100, single=False)[0], 2e-5, 2e-5)

    Xf = _X(g["x16"].astype(np.float32), h, M)
    Y, waH, nu = restate.gsc_lms(Xf, FS, d)                        # confs/gsclms.json = the class defaults
    assert nu == int(g["n_updates_gsclms"]) and rel_l2(waH, g["waH_gsclms"]) < 1e-10
    check("gsclms", Y, 1e-6, 1e-6)
    Y, waH, nu = restate.gsc_rls(Xf, FS, d)                        # confs/gscrls.json = the class defaults
    assert nu == int(g["n_updates_gscrls"]) and rel_l2(waH, g["waH_gscrls"]) < 1e-8
    check("gscrls", Y, 1e-6, 1e-6)
    # confs/lcmv_and_zelinski.json: target at 0 rad, a null on -1.306379 rad (the reference's calcMainlobeN weights; bin M/2 is the
    # reference's own cascade quirk, restated only on the device side and checked there)
    assert np.allclose(g["lcmv_dT"], restate.calc_delays("linear", mpos, [0.0, None, None]), rtol=0, atol=1e-18)
    assert np.allclose(g["lcmv_dJ"][0], d, rtol=0, atol=1e-18)
    wl = restate.calc_mainlobe_n(M, C, FS, g["lcmv_dT"], g["lcmv_dJ"])
    assert rel_l2(wl[:M // 2], g["lcmv_w"][:M // 2]) < 1e-12
    k = np.arange(1, M // 2)
    vj = np.exp(-2j * np.pi * k[:, None] * FS * g["lcmv_dJ"][0][None, :] / M)
    assert np.abs(np.einsum("kc,kc->k", np.conj(g["lcmv_w"][1:M // 2]), vj)).max() < 1e-9    # the null really sits on the jammer


def test_sos_batch_beamforming_vad_on_the_references_own_fixtures():
    """unit_test/test_sos_batch_beamforming.py on its default inputs (the whole Kinect recording, shipped prototypes) with
    confs/{bmvdr_vad, gev_vad, smimvdr}.json (VAD label [[1.5, 4.0]]): restatement vs the reference's outputs
    (golden_sos_kinect_vad_c4_m256, tests/golden/make_golden_sos.py kinect_vad())."""
    g = load_golden("sos_kinect_vad_c4_m256"); x16 = load_golden("online_kinect_c4_m256")["x16"]
    p = np.load(os.path.join(GOLDEN, "prototype_shipped_M256_m4_r1.npz")); h, gg = p["h"], p["g"]
    M, K, C = 256, 129, 4
    f0, f1 = [int(v) for v in g["frames"]]
    X = _X(x16.astype(np.float32), h, M)
    labels = [tuple(l) for l in g["labels"]]
    Rt, Rn, ct, cn = restate.sos_accumulate(X, FS, M // 2, target_labs=labels, energy_threshold=10.0)
    assert np.array_equal(ct, g["ct"]) and np.array_equal(cn, g["cn"]) and ct[0] == 312 and cn[0] == 298
    w = restate.sos_bmvdr_weights(Rt, Rn, ct, cn, gamma=1e-6, ref_micx=0, offset=0.0)
    assert rel_l2(w, g["w_bmvdr_vad"]) < 1e-9
    Y = restate.sos_apply(X, w)
    assert rel_l2(Y[f0:f1, :K], g["Y_bmvdr_vad"]) < 1e-6 and rel_l2(restate.synthesis(Y, gg, M, 4, 1), g["time_bmvdr_vad"]) < 1e-6
    w = restate.sos_gev_weights(Rt, Rn, cn, gamma=1e-6)
    w = w * np.sign(np.real(np.vdot(w[0], g["w_gev_vad"][0])))          # scipy's eigh leaves one global sign to LAPACK
    assert rel_l2(w, g["w_gev_vad"]) < 1e-8
    Y = restate.sos_apply(X, w)
    assert rel_l2(Y[f0:f1, :K], g["Y_gev_vad"]) < 1e-6 and rel_l2(restate.synthesis(Y, gg, M, 4, 1), g["time_gev_vad"]) < 1e-6
    # SMI-MVDR (confs/smimvdr.json): noise covariance outside the label, mu = 1e-4, look direction of the file
    R, nf = restate.smi_covariance(X, FS, M // 2, tuple(labels), 10.0)
    assert nf == 298 and rel_l2(R, g["cov_smimvdr"]) < 1e-6
    wq = restate.calc_mainlobe(M, C, FS, g["delays"])
    ws = restate.calc_mvdr_weights(R + float(np.float32(g["mu_smimvdr"])) * np.eye(C), wq, single=True)
    assert rel_l2(ws[1:], g["w_smimvdr"][1:]) < 1e-4                    # the reference's float LINPACK SVD
    Y = restate.subband_mvdr(X, g["w_smimvdr"])
    assert rel_l2(Y[f0:f1, :K], g["Y_smimvdr"]) < 1e-6 and rel_l2(restate.synthesis(Y, gg, M, 4, 1), g["time_smimvdr"]) < 1e-6
    Y = restate.subband_mvdr(X, restate.calc_mvdr_weights(R + float(np.float32(g["mu_smimvdr"])) * np.eye(C), wq, single=False))
    assert rel_l2(restate.synthesis(Y, gg, M, 4, 1), g["time_smimvdr"]) < 1e-4   # fp64 solve vs the reference's float SVD: inside the parity budget


def test_wpe_on_the_references_own_fixtures():
    """unit_test/test_subband_dereverberator.py on its default inputs (the whole Kinect recording, shipped prototypes) with
    confs/wpe.json (lags 0..32, 2 iterations, -18 dB, bias 1e-4): multi-channel and single-channel restatement vs the compiled
    reference's outputs (golden_wpe_kinect_c4_m256, tests/golden/make_golden_wpe.py kinect())."""
    import json
    g = load_golden("wpe_kinect_c4_m256"); x16 = load_golden("online_kinect_c4_m256")["x16"]
    p = np.load(os.path.join(GOLDEN, "prototype_shipped_M256_m4_r1.npz")); h, gg = p["h"], p["g"]
    M, K = 256, 129
    conf = json.loads(str(g["conf"]))
    assert conf == dict(lower_num=0, upper_num=32, iterations_num=2, load_db=-18.0, band_width=0.0, diagonal_bias=0.0001)
    f0, f1 = [int(v) for v in g["frames"]]
    X = _X(x16.astype(np.float32), h, M)
    Xm, _, used = restate.wpe(X, samplerate=FS, **conf)
    assert used == int(g["used_multi"]) == 614
    assert rel_l2(Xm[f0:f1, :, :K], g["X_multi"]) < 1e-6                  # golden stored as complex64
    for c in range(4):
        assert rel_l2(restate.synthesis(Xm[:, c, :], gg, M, 4, 1), g["time_multi"][c]) < 1e-6, c
    assert rel_l2(Xm[f0:f1, :, :K], X[f0:f1, :, :K]) > 0.05              # the filters really remove something
    ks = dict(conf); ks["diagonal_bias"] = 0.0                           # the single-channel class has no diagonal bias
    Xs, _, used = restate.wpe(X[:, :1, :], samplerate=FS, **ks)
    assert used == int(g["used_single"]) and rel_l2(Xs[f0:f1, 0, :K], g["X_single"]) < 1e-6
    assert rel_l2(restate.synthesis(Xs[:, 0, :], gg, M, 4, 1), g["time_single"]) < 1e-6
