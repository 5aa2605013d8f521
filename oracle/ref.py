"""oracle/ref.py — ctypes access to oracle/_ref/libbtkref.so (the reference's own C++ hot path, compiled unmodified
against oracle/gsl_shim by oracle/Makefile; driver in oracle/ref_harness.cc).

TEST INFRASTRUCTURE ONLY: importable from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs.  Never from the product package.
"""
import ctypes as ct
import os
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_ref", "libbtkref.so")

BF_DS, BF_GSC, BF_MVDR_SD, BF_SMI_MVDR, BF_GSC_LMS, BF_GSC_RLS_CPP = 0, 1, 2, 3, 4, 5


class RefConfig(ct.Structure):
    _fields_ = [
        ("C", ct.c_int), ("M", ct.c_int), ("m", ct.c_int), ("r", ct.c_int), ("delay_compensation_type", ct.c_int),
        ("samplerate", ct.c_double),
        ("bf_kind", ct.c_int), ("pf_kind", ct.c_int),
        ("pf_alpha", ct.c_double), ("pf_type", ct.c_int), ("pf_min_frames", ct.c_int),
        ("mvdr_mu", ct.c_double), ("sspeed", ct.c_double),
        ("smi_target_start", ct.c_double), ("smi_target_end", ct.c_double), ("smi_energy_threshold", ct.c_double),
        ("lms_beta", ct.c_double), ("lms_gamma", ct.c_double), ("lms_init_diagonal_load", ct.c_double),
        ("lms_regularization_param", ct.c_double), ("lms_energy_floor", ct.c_double), ("lms_sil_thresh", ct.c_double),
        ("lms_max_wa_l2norm", ct.c_double),
        ("lms_min_frames", ct.c_int), ("lms_slowdown_after", ct.c_int),
        ("do_synthesis", ct.c_int),
        ("pf_threshold", ct.c_double), ("pf_min_sv", ct.c_double), ("pf_diag_load", ct.c_double), ("pf_fbin1", ct.c_int),
        ("rls_mu", ct.c_double), ("rls_sigma2", ct.c_double), ("rls_init_sigma2", ct.c_double), ("rls_alpha", ct.c_double), ("rls_qctype", ct.c_int),
    ]


_lib = None


def available():
    return os.path.exists(LIB_PATH)


def lib():
    global _lib
    if _lib is None:
        if not available():
            raise RuntimeError("oracle/_ref/libbtkref.so is missing: run `make -C oracle` where /root/reference exists")
        _lib = ct.CDLL(LIB_PATH)
        _lib.ref_analysis.restype = ct.c_int
        _lib.ref_synthesis.restype = ct.c_int
        _lib.ref_beamform.restype = ct.c_int
        _lib.ref_pseudoinverse.restype = ct.c_int
    return _lib


def _p(a, ty):
    return None if a is None else a.ctypes.data_as(ct.POINTER(ty))


def _num_frames(n, M, m, r, dct):
    from . import restate
    return restate.num_frames(n, M, m, r, dct)


def analysis(x, h, M, m, r, dct=2):
    x = np.ascontiguousarray(x, np.float32)
    h = np.ascontiguousarray(h, np.float64)
    Tcap = _num_frames(len(x), M, m, r, dct) + 8
    out = np.zeros((Tcap, M), np.complex128)
    T = lib().ref_analysis(_p(x, ct.c_float), ct.c_int(len(x)), _p(h, ct.c_double), M, m, r, dct, _p(out, ct.c_double), Tcap)
    assert T >= 0
    return out[:T].copy()


def synthesis(Y, g, M, m, r, dct=2):
    Y = np.ascontiguousarray(Y, np.complex128)
    g = np.ascontiguousarray(g, np.float64)
    T = Y.shape[0]
    D = M >> r
    out = np.zeros(((T + 4) * D,), np.float32)
    nb = lib().ref_synthesis(_p(Y, ct.c_double), T, _p(g, ct.c_double), M, m, r, dct, _p(out, ct.c_float), T + 4)
    assert nb >= 0
    return out[:nb * D].copy()


def wpe(X, lower_num=0, upper_num=32, iterations_num=2, load_db=-20.0, band_width=0.0, diagonal_bias=0.001, samplerate=16000.0,
        start_frame_no=0, end_frame_no=-1):
    """MultiChannelWPEDereverberation + MultiChannelWPEDereverberationFeature (dereverberation.cc:312-733) on snapshots
    X[T][C][M] -> dereverberated X'[T][C][M]; returns (X', frames used for the estimation)."""
    X = np.asarray(X, np.complex128)
    T, C, M = X.shape
    Xin = np.ascontiguousarray(np.transpose(X, (1, 0, 2)))
    Xout = np.zeros_like(Xin)
    L = lib()
    L.ref_wpe.restype = ct.c_int
    used = L.ref_wpe(_p(Xin, ct.c_double), ct.c_int(C), ct.c_int(T), ct.c_int(M), ct.c_int(lower_num), ct.c_int(upper_num), ct.c_int(iterations_num),
                     ct.c_double(load_db), ct.c_double(band_width), ct.c_double(diagonal_bias), ct.c_double(samplerate),
                     ct.c_int(start_frame_no), ct.c_int(end_frame_no), _p(Xout, ct.c_double))
    assert used >= 0, "the reference raised an exception"
    return np.ascontiguousarray(np.transpose(Xout, (1, 0, 2))), used


def wpe_single(X, lower_num=0, upper_num=64, iterations_num=2, load_db=-20.0, band_width=0.0, samplerate=16000.0, start_frame_no=0, end_frame_no=-1):
    """SingleChannelWPEDereverberationFeature (dereverberation.cc:24-310) on one channel's snapshots X[T][M] -> X'[T][M];
    returns (X', frames used for the estimation)."""
    X = np.ascontiguousarray(X, np.complex128)
    T, M = X.shape
    Xout = np.zeros_like(X)
    L = lib()
    L.ref_wpe_single.restype = ct.c_int
    used = L.ref_wpe_single(_p(X, ct.c_double), ct.c_int(T), ct.c_int(M), ct.c_int(lower_num), ct.c_int(upper_num), ct.c_int(iterations_num),
                            ct.c_double(load_db), ct.c_double(band_width), ct.c_double(samplerate), ct.c_int(start_frame_no), ct.c_int(end_frame_no),
                            _p(Xout, ct.c_double))
    assert used >= 0, "the reference raised an exception"
    return Xout, used


def gsc_weights(M, C, samplerate, delays, want_B=True):
    delays = np.ascontiguousarray(delays, np.float64)
    wq = np.zeros((M, C), np.complex128)
    B = np.zeros((M, C, C - 1), np.complex128) if (want_B and C > 1) else None
    lib().ref_gsc_weights(M, C, ct.c_double(samplerate), _p(delays, ct.c_double), _p(wq, ct.c_double), _p(B, ct.c_double))
    return wq, B


def lcmv_weights(M, C, NC, samplerate, delaysT, delaysJ, want_B=True):
    delaysT = np.ascontiguousarray(delaysT, np.float64)
    delaysJ = np.ascontiguousarray(np.atleast_2d(delaysJ), np.float64)
    K = M // 2 + 1
    wq = np.zeros((K, C), np.complex128)
    B = np.zeros((K, C, C - NC), np.complex128) if want_B else None
    lib().ref_lcmv_weights(M, C, NC, ct.c_double(samplerate), _p(delaysT, ct.c_double), _p(delaysJ, ct.c_double), _p(wq, ct.c_double), _p(B, ct.c_double))
    return wq, B


def pseudoinverse(A, thr=1.0e-8):
    A = np.ascontiguousarray(A, np.complex128)
    n = A.shape[0]
    out = np.zeros((n, n), np.complex128)
    ok = lib().ref_pseudoinverse(_p(A, ct.c_double), n, ct.c_double(thr), _p(out, ct.c_double))
    return out, bool(ok)


def mvdr_weights(M, C, samplerate, delays, R=None, mpos=None, sspeed=343740.0, mu=1.0e-4):
    delays = np.ascontiguousarray(delays, np.float64)
    K = M // 2 + 1
    w = np.zeros((K, C), np.complex128)
    Rc = None if R is None else np.ascontiguousarray(R, np.complex128)
    mp = None if mpos is None else np.ascontiguousarray(mpos, np.float64)
    lib().ref_mvdr_weights(M, C, ct.c_double(samplerate), _p(delays, ct.c_double), _p(Rc, ct.c_double), _p(mp, ct.c_double),
                           ct.c_double(sspeed), ct.c_double(mu), _p(w, ct.c_double))
    return w


def beamform(samples, h, g, delays, M, m=4, r=1, dct=2, samplerate=16000.0, bf_kind=BF_DS, wa=None, mpos=None,
             pf=None, mvdr_mu=1.0e-4, sspeed=343740.0, smi_label=(1.0, -1.0), smi_energy_threshold=10.0, lms=None,
             do_synthesis=True, want_subband=True, rls=None):
    """Run the reference pipe on one utterance.  samples float32 [C][n].
    Returns dict(Y=[T][M] complex128, time=float32[nb*D], cov, w, stats)."""
    from . import restate
    samples = np.ascontiguousarray(samples, np.float32)
    C, n = samples.shape
    D = M >> r
    K = M // 2 + 1
    lp = dict(restate.DEFAULT_LMS)
    if lms:
        lp.update(lms)
    cfg = RefConfig(C=C, M=M, m=m, r=r, delay_compensation_type=dct, samplerate=samplerate, bf_kind=bf_kind,
                    pf_kind=0 if pf is None else {"zelinski": 1, "mccowan": 2, "lefkimmiatis": 3}[pf.get("kind", "zelinski")],
                    pf_threshold=0.99 if pf is None else pf.get("threshold", 0.99), pf_min_sv=1e-8 if pf is None else pf.get("min_sv", 1e-8),
                    pf_diag_load=0.0 if pf is None else pf.get("diag_load", 0.01), pf_fbin1=0 if pf is None else pf.get("fbin1", 0),
                    pf_alpha=0.0 if pf is None else pf.get("alpha", 0.6), pf_type=2 if pf is None else pf.get("type", 2),
                    pf_min_frames=0 if pf is None else pf.get("min_frames", 0),
                    mvdr_mu=mvdr_mu, sspeed=sspeed, smi_target_start=smi_label[0], smi_target_end=smi_label[1],
                    smi_energy_threshold=smi_energy_threshold,
                    lms_beta=lp["beta"], lms_gamma=lp["gamma"], lms_init_diagonal_load=lp["init_diagonal_load"],
                    lms_regularization_param=lp["regularization_param"], lms_energy_floor=lp["energy_floor"],
                    lms_sil_thresh=lp["sil_thresh"], lms_max_wa_l2norm=lp["max_wa_l2norm"],
                    lms_min_frames=lp["min_frames"], lms_slowdown_after=lp["slowdown_after"],
                    do_synthesis=1 if do_synthesis else 0)
    if rls:   # SubbandGSCRLS(fftLen, False, myu, sigma2); init_precision_matrix(init_sigma2); set_quadratic_constraint(alpha, qctype)
        cfg.rls_mu = rls.get("mu", 0.9); cfg.rls_sigma2 = rls.get("sigma2", 0.01); cfg.rls_init_sigma2 = rls.get("init_sigma2", 0.01)
        cfg.rls_alpha = rls.get("alpha", -1.0); cfg.rls_qctype = rls.get("qctype", 0)
    Tcap = _num_frames(n, M, m, r, dct) + 8
    Y = np.zeros((Tcap, M), np.complex128) if want_subband else None
    out_time = np.zeros((Tcap * D,), np.float32)
    nblocks = ct.c_int(0)
    cov = np.zeros((K, C, C), np.complex128) if bf_kind == BF_SMI_MVDR else None
    w = np.zeros((K, C), np.complex128)
    stats = np.zeros(3, np.float64)
    h = np.ascontiguousarray(h, np.float64)
    g = np.ascontiguousarray(g, np.float64)
    delays = np.ascontiguousarray(delays, np.float64)
    wa_p = None if wa is None else np.ascontiguousarray(wa, np.float64)
    mp = None if mpos is None else np.ascontiguousarray(mpos, np.float64)
    T = lib().ref_beamform(ct.byref(cfg), _p(samples, ct.c_float), n, _p(h, ct.c_double), _p(g, ct.c_double), _p(delays, ct.c_double),
                           _p(wa_p, ct.c_double), _p(mp, ct.c_double), _p(Y, ct.c_double), Tcap, _p(out_time, ct.c_float), Tcap,
                           ct.byref(nblocks), _p(cov, ct.c_double), _p(w, ct.c_double), _p(stats, ct.c_double))
    res = dict(Y=None if Y is None else Y[:T].copy(), time=out_time[:nblocks.value * D].copy(), cov=cov, stats=stats)
    if bf_kind == BF_GSC_LMS:
        res["w"] = w.reshape(-1)[:K * (C - 1)].reshape(K, C - 1).copy()
    else:
        res["w"] = w
    return res
