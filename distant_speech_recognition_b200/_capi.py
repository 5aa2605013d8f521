"""ctypes binding of include/btkb.h (libbtkb.so: hand-written sm_100a CUDA behind a C-ABI).

There is NO CPU fallback: importing this module fails if the shared library has not been built
(`python -c "import __graft_entry__ as g; g.build()"` or `make -C distant_speech_recognition_b200/csrc`), and
`Pipeline(...)` raises if no CUDA device is usable.
"""
import ctypes as ct
import os
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libbtkb.so")
HEADER_PATH = os.path.join(os.path.dirname(_HERE), "include", "btkb.h")

BF_DS, BF_GSC, BF_MVDR, BF_GSC_LMS, BF_GSC_RLS, BF_GSC_RLS_CPP = 0, 1, 2, 3, 4, 5
SOS_BMVDR, SOS_GEV = 0, 1
PF_NONE, PF_ZELINSKI, PF_MCCOWAN, PF_LEFKIMMIATIS = 0, 1, 2, 3
OK, ERR_INVALID, ERR_NO_DEVICE, ERR_CUDA, ERR_STATE, ERR_ALLOC = 0, -1, -2, -3, -4, -5


class BtkbError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("btkb error %d: %s" % (code, msg))
        self.code = code


class LmsParams(ct.Structure):
    _fields_ = [("beta", ct.c_float), ("gamma", ct.c_float), ("init_diagonal_load", ct.c_float),
                ("regularization_param", ct.c_float), ("energy_floor", ct.c_float), ("sil_thresh", ct.c_float),
                ("max_wa_l2norm", ct.c_float), ("min_frames", ct.c_int), ("slowdown_after", ct.c_int)]


class RlsParams(ct.Structure):
    _fields_ = [("beta", ct.c_float), ("gamma", ct.c_float), ("mu", ct.c_float), ("init_diagonal_load", ct.c_float),
                ("regularization_param", ct.c_float), ("sil_thresh", ct.c_float), ("alpha2", ct.c_float), ("max_wa_l2norm", ct.c_float),
                ("constraint_option", ct.c_int), ("min_frames", ct.c_int)]


class RlsCppParams(ct.Structure):
    _fields_ = [("mu", ct.c_float), ("sigma2", ct.c_float), ("init_sigma2", ct.c_float), ("alpha", ct.c_float), ("qctype", ct.c_int), ("update", ct.c_int)]


class WpeParams(ct.Structure):
    _fields_ = [("enabled", ct.c_int), ("lower_num", ct.c_int), ("upper_num", ct.c_int), ("iterations_num", ct.c_int),
                ("load_db", ct.c_double), ("band_width", ct.c_double), ("diagonal_bias", ct.c_double), ("fp32_normal_equations", ct.c_int)]


class Config(ct.Structure):
    _fields_ = [("device", ct.c_int), ("channels", ct.c_int), ("fft_len", ct.c_int), ("m", ct.c_int), ("r", ct.c_int),
                ("delay_compensation_type", ct.c_int), ("samplerate", ct.c_float), ("beamformer", ct.c_int),
                ("postfilter", ct.c_int), ("pf_alpha", ct.c_float), ("pf_type", ct.c_int), ("pf_min_frames", ct.c_int),
                ("lms", LmsParams), ("max_utterances", ct.c_int), ("max_samples", ct.c_int), ("keep_snapshots", ct.c_int), ("synthesis_gain", ct.c_int), ("normalize_weight", ct.c_int),
                ("pf_threshold", ct.c_float), ("pf_min_sv", ct.c_double), ("pf_fbin1", ct.c_int), ("rls", RlsParams), ("wpe", WpeParams), ("rls_cpp", RlsCppParams)]


def _load():
    if not os.path.exists(LIB_PATH):
        raise ImportError("libbtkb.so is not built (%s). Build it with __graft_entry__.build(); there is no CPU path." % LIB_PATH)
    lib = ct.CDLL(LIB_PATH)
    lib.btkb_last_error.restype = ct.c_char_p
    lib.btkb_destroy.restype = None
    lib.btkb_default_config.restype = None
    return lib


lib = _load()


def _check(rc):
    if rc != 0:
        raise BtkbError(rc, lib.btkb_last_error().decode("utf-8", "replace"))


def device_count():
    return int(lib.btkb_device_count())


def _fp(a):
    return a.ctypes.data_as(ct.POINTER(ct.c_float))


def _dp(a):
    return a.ctypes.data_as(ct.POINTER(ct.c_double))


class Pipeline:
    """One batch pipeline: analysis -> per-bin beamformer (+post-filter) -> synthesis on one GPU."""

    def __init__(self, channels, fft_len=512, m=4, r=1, delay_compensation_type=2, samplerate=16000.0, beamformer=BF_DS,
                 postfilter=PF_NONE, pf_alpha=0.6, pf_type=2, pf_min_frames=0, lms=None, max_utterances=1,
                 max_samples=160000, device=0, normalize_weight=False, pf_threshold=0.99, pf_min_sv=1.0e-8, pf_fbin1=0, rls=None, wpe=None, rls_cpp=None):
        cfg = Config()
        lib.btkb_default_config(ct.byref(cfg))
        cfg.device = device; cfg.channels = channels; cfg.fft_len = fft_len; cfg.m = m; cfg.r = r
        cfg.delay_compensation_type = delay_compensation_type; cfg.samplerate = samplerate
        cfg.beamformer = beamformer; cfg.postfilter = postfilter
        cfg.pf_alpha = pf_alpha; cfg.pf_type = pf_type; cfg.pf_min_frames = pf_min_frames
        cfg.pf_threshold = pf_threshold; cfg.pf_min_sv = pf_min_sv; cfg.pf_fbin1 = pf_fbin1
        if lms:
            for k, v in lms.items():
                setattr(cfg.lms, k, v)
        if rls:
            for k, v in rls.items():
                if k != "slowdown_after":   # a constructor argument the reference's RLS loop never reads
                    setattr(cfg.rls, k, v)
        if rls_cpp:
            for k, v in rls_cpp.items():
                setattr(cfg.rls_cpp, k, v)
        if wpe is not None:
            cfg.wpe.enabled = 1
            for k, v in wpe.items():
                setattr(cfg.wpe, k, v)
        cfg.max_utterances = max_utterances; cfg.max_samples = max_samples; cfg.normalize_weight = 1 if normalize_weight else 0
        self.cfg = cfg
        self.C, self.M, self.K, self.D = channels, fft_len, fft_len // 2 + 1, fft_len >> r
        self._h = ct.c_void_p()
        _check(lib.btkb_create(ct.byref(cfg), ct.byref(self._h)))
        self.U = 0
        self.NC = 1

    def close(self):
        if self._h:
            lib.btkb_destroy(self._h)
            self._h = ct.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- setup
    def set_prototypes(self, h, g=None):
        h = np.ascontiguousarray(h, np.float64)
        gp = None
        if g is not None:
            g = np.ascontiguousarray(g, np.float64)
            gp = _dp(g)
        _check(lib.btkb_set_prototypes(self._h, _dp(h), gp, ct.c_int(len(h))))

    def set_delays(self, delays):
        d = np.ascontiguousarray(np.atleast_2d(delays), np.float64)
        self.U = d.shape[0]
        _check(lib.btkb_set_delays(self._h, ct.c_int(d.shape[0]), _dp(d)))

    def set_delays_lcmv(self, delaysT, delaysJ):
        dT = np.ascontiguousarray(np.atleast_2d(delaysT), np.float64)
        dJ = np.ascontiguousarray(delaysJ, np.float64).reshape(dT.shape[0], -1, self.C)
        self.U = dT.shape[0]; self.NC = dJ.shape[1] + 1
        _check(lib.btkb_set_delays_lcmv(self._h, ct.c_int(self.U), ct.c_int(self.NC), _dp(dT), _dp(dJ)))

    def spectral_matrix_update(self, mu, legacy_noconj=True):
        _check(lib.btkb_spectral_matrix_update(self._h, ct.c_float(mu), ct.c_int(1 if legacy_noconj else 0)))

    def set_weights(self, w):
        w = np.ascontiguousarray(w, np.complex64)
        self.U = w.shape[0]
        _check(lib.btkb_set_weights(self._h, ct.c_int(w.shape[0]), _fp(w)))

    def set_active_weights(self, wa):
        wa = np.ascontiguousarray(wa, np.complex64)
        _check(lib.btkb_set_active_weights(self._h, ct.c_int(wa.shape[0]), _fp(wa)))

    def set_blocking_source(self, from_mvdr_weights):
        _check(lib.btkb_set_blocking_source(self._h, ct.c_int(1 if from_mvdr_weights else 0)))

    def set_noise_covariance(self, R):
        R = np.ascontiguousarray(R, np.complex64)
        self.U = R.shape[0]
        _check(lib.btkb_set_noise_covariance(self._h, ct.c_int(R.shape[0]), _fp(R)))

    def set_diffuse_noise_model(self, U, mpos, sspeed=343740.0):
        mp = np.ascontiguousarray(mpos, np.float64)
        _check(lib.btkb_set_diffuse_noise_model(self._h, ct.c_int(U), _dp(mp), ct.c_float(sspeed)))

    # ---- multi-channel WPE (dereverberation.cc:312-733)
    def run_wpe(self, start_frame_no=0, end_frame_no=-1):
        _check(lib.btkb_run_wpe(self._h, ct.c_int(start_frame_no), ct.c_int(end_frame_no)))

    def apply_wpe(self):
        _check(lib.btkb_apply_wpe(self._h))

    def get_wpe_filter(self):
        P = self.cfg.wpe.upper_num - self.cfg.wpe.lower_num + 1
        out = np.empty((self.U, self.K, self.C, self.C * P), np.complex64)
        _check(lib.btkb_get_wpe_filter(self._h, _fp(out)))
        return out

    def set_wpe_filter(self, G):
        """G complex64 [U][K][C][C*P] as returned by get_wpe_filter (of this or another pipeline)."""
        G = np.ascontiguousarray(G, np.complex64)
        self.U = G.shape[0]
        _check(lib.btkb_set_wpe_filter(self._h, ct.c_int(G.shape[0]), _fp(G)))

    def last_timing_wpe(self):
        ms = ct.c_float(0)
        _check(lib.btkb_last_timing_wpe(self._h, ct.byref(ms)))
        return float(ms.value)

    def set_snapshots(self, X):
        """X complex64 [U][T][C][K]: snapshots computed elsewhere take the place of the analysis output."""
        X = np.ascontiguousarray(X, np.complex64)
        assert X.ndim == 4 and X.shape[2] == self.C and X.shape[3] == self.K
        self.U = X.shape[0]
        _check(lib.btkb_set_snapshots(self._h, ct.c_int(X.shape[0]), ct.c_int(X.shape[1]), _fp(X)))

    def upgrade_blocking_matrix(self):
        """SubbandMVDRGSC::upgrade_blocking_matrix (beamformer.cc:2674-2691)."""
        _check(lib.btkb_upgrade_blocking_matrix(self._h))

    def blocking_matrix_output(self, out_chan):
        """b_i^H x for every frame: complex64 [U][T][K] (SubbandMVDRGSC::blocking_matrix_output, beamformer.cc:2693-2716)."""
        out = np.empty((self.U, self.num_frames(), self.K), np.complex64)
        _check(lib.btkb_blocking_matrix_output(self._h, ct.c_int(out_chan), _fp(out)))
        return out

    def last_wpe_form(self):
        """0: the last estimation solved the lag-domain normal equations (L x L), 1: the frame-domain ones (S x S)."""
        f = ct.c_int(-1)
        _check(lib.btkb_last_wpe_form(self._h, ct.byref(f)))
        return int(f.value)

    # ---- noise coherence of the McCowan / Lefkimmiatis post-filters (postfilter.cc:541-680)
    def pf_set_diffuse_noise_model(self, mpos, samplerate=16000.0, sspeed=343740.0):
        mp = np.ascontiguousarray(mpos, np.float64)
        _check(lib.btkb_pf_set_diffuse_noise_model(self._h, _dp(mp), ct.c_double(samplerate), ct.c_double(sspeed)))

    def pf_set_noise_coherence(self, R):
        R = np.ascontiguousarray(R, np.complex128)
        assert R.shape == (self.K, self.C, self.C)
        _check(lib.btkb_pf_set_noise_coherence(self._h, R.ctypes.data_as(ct.POINTER(ct.c_double))))

    def pf_get_noise_coherence(self):
        R = np.empty((self.K, self.C, self.C), np.complex128)
        _check(lib.btkb_pf_get_noise_coherence(self._h, R.ctypes.data_as(ct.POINTER(ct.c_double))))
        return R

    def pf_set_diagonal_loading(self, mu):
        _check(lib.btkb_pf_set_diagonal_loading(self._h, ct.c_float(mu)))

    def pf_divide_nondiagonal(self, mu):
        _check(lib.btkb_pf_divide_nondiagonal(self._h, ct.c_float(mu)))

    def calc_mvdr_weights(self, mu, dthreshold=None):
        """dthreshold: the reference's singular-value floor (default 1e-8, beamformer.i:414-486): bins below it get the identity."""
        if dthreshold is None:
            _check(lib.btkb_calc_mvdr_weights(self._h, ct.c_float(mu)))
        else:
            _check(lib.btkb_calc_mvdr_weights_ex(self._h, ct.c_float(mu), ct.c_float(dthreshold)))

    # ---- data path
    def submit(self, samples, lengths=None):
        """samples float32 [U][C][n] (host).  Keeps a reference until the next submit (the H2D copy is asynchronous)."""
        s = np.ascontiguousarray(samples, np.float32)
        assert s.ndim == 3 and s.shape[1] == self.C
        self._keep = s
        self.U, self.n = s.shape[0], s.shape[2]
        lp = None
        if lengths is not None:
            self._len = np.ascontiguousarray(lengths, np.int32)
            lp = self._len.ctypes.data_as(ct.POINTER(ct.c_int))
        _check(lib.btkb_submit(self._h, _fp(s), ct.c_int(self.U), ct.c_int(self.n), lp))

    def submit_pointer(self, host_ptr, U, n, lengths=None):
        """Raw host pointer variant (e.g. pinned torch tensor .data_ptr())."""
        self.U, self.n = U, n
        lp = None
        if lengths is not None:
            self._len = np.ascontiguousarray(lengths, np.int32)
            lp = self._len.ctypes.data_as(ct.POINTER(ct.c_int))
        _check(lib.btkb_submit(self._h, ct.cast(ct.c_void_p(host_ptr), ct.POINTER(ct.c_float)), ct.c_int(U), ct.c_int(n), lp))

    def submit_i16_pointer(self, host_ptr, U, n, lengths=None):
        """16-bit PCM host buffer [U][C][n] (e.g. a pinned torch int16 tensor's data_ptr())."""
        self.U, self.n = U, n
        lp = None
        if lengths is not None:
            self._len = np.ascontiguousarray(lengths, np.int32)
            lp = self._len.ctypes.data_as(ct.POINTER(ct.c_int))
        _check(lib.btkb_submit_i16(self._h, ct.cast(ct.c_void_p(host_ptr), ct.POINTER(ct.c_int16)), ct.c_int(U), ct.c_int(n), lp))

    def submit_i16(self, samples, lengths=None):
        s = np.ascontiguousarray(samples, np.int16)
        self._keep = s
        self.submit_i16_pointer(s.ctypes.data, s.shape[0], s.shape[2], lengths)

    def submit_device(self, dev_ptr, U, n, lengths=None):
        self.U, self.n = U, n
        lp = None
        if lengths is not None:
            self._len = np.ascontiguousarray(lengths, np.int32)
            lp = self._len.ctypes.data_as(ct.POINTER(ct.c_int))
        _check(lib.btkb_submit_device(self._h, ct.cast(ct.c_void_p(dev_ptr), ct.POINTER(ct.c_float)), ct.c_int(U), ct.c_int(n), lp))

    def set_subband(self, Y):
        Y = np.ascontiguousarray(Y, np.complex64)
        self.U = Y.shape[0]
        _check(lib.btkb_set_subband(self._h, ct.c_int(Y.shape[0]), ct.c_int(Y.shape[1]), _fp(Y)))

    def run_synthesis(self):
        _check(lib.btkb_run_synthesis(self._h))

    def run(self, synthesis=True):
        _check(lib.btkb_run(self._h, ct.c_int(1 if synthesis else 0)))

    def run_analysis(self):
        _check(lib.btkb_run_analysis(self._h))

    def run_beamformer(self, synthesis=True):
        _check(lib.btkb_run_beamformer(self._h, ct.c_int(1 if synthesis else 0)))

    def accumulate_covariance(self, labels=None, energy_threshold=10.0):
        lp = None
        if labels is not None:
            self._labels = np.ascontiguousarray(labels, np.float64)
            lp = _dp(self._labels)
        _check(lib.btkb_accumulate_covariance(self._h, lp, ct.c_float(energy_threshold)))

    # ---- SOS batch beamformers: blind MVDR / GEV (lib/pybeamformer.py:1026-1357)
    def sos_reset_stats(self):
        _check(lib.btkb_sos_reset_stats(self._h))

    def sos_accumulate_from_label(self, labels, energy_threshold=10.0):
        """labels [U][NL][2] (or [NL][2] for one utterance): target segments in seconds."""
        lab = np.ascontiguousarray(labels, np.float64)
        if lab.ndim == 2:
            lab = lab[None]
        assert lab.ndim == 3 and lab.shape[2] == 2 and lab.shape[0] == self.U
        _check(lib.btkb_sos_accumulate_from_label(self._h, _dp(lab), ct.c_int(lab.shape[1]), ct.c_float(energy_threshold)))

    def sos_accumulate_from_tfmask(self, mask_t, mask_j, energy_threshold=10.0):
        """mask_t, mask_j [U][Tm][K] (or [Tm][K] for one utterance)."""
        mt = np.ascontiguousarray(mask_t, np.float32); mj = np.ascontiguousarray(mask_j, np.float32)
        if mt.ndim == 2:
            mt, mj = mt[None], mj[None]
        assert mt.shape == mj.shape and mt.shape[0] == self.U and mt.shape[2] == self.K
        _check(lib.btkb_sos_accumulate_from_tfmask(self._h, _fp(mt), _fp(mj), ct.c_int(mt.shape[1]), ct.c_float(energy_threshold)))

    def sos_calc_weights(self, kind, gamma=1e-6, ref_micx=0, offset=0.0):
        _check(lib.btkb_sos_calc_weights(self._h, ct.c_int(kind), ct.c_double(gamma), ct.c_int(ref_micx), ct.c_double(offset)))

    def sos_get_stats(self):
        """(Rt, Rn complex128 [U][K][C][C] raw sums, counts float64 [U][K][2])."""
        Rt = np.empty((self.U, self.K, self.C, self.C), np.complex128); Rn = np.empty_like(Rt)
        cnt = np.empty((self.U, self.K, 2), np.float64)
        _check(lib.btkb_sos_get_stats(self._h, Rt.ctypes.data_as(ct.POINTER(ct.c_double)), Rn.ctypes.data_as(ct.POINTER(ct.c_double)), _dp(cnt)))
        return Rt, Rn, cnt

    def synchronize(self):
        _check(lib.btkb_synchronize(self._h))

    # ---- streamed chunks with carried state (btkb.h: btkb_stream_*)
    def stream_begin(self, U):
        self.U = int(U)
        _check(lib.btkb_stream_begin(self._h, ct.c_int(self.U)))

    def stream_submit(self, samples, lengths=None, final=False, synthesis=True):
        """samples float32 [U][C][n]: the next n samples of every utterance (n a multiple of D unless final)."""
        x = np.ascontiguousarray(samples, np.float32)
        assert x.ndim == 3 and x.shape[1] == self.C
        ln = None if lengths is None else np.ascontiguousarray(lengths, np.int32)
        _check(lib.btkb_stream_submit(self._h, _fp(x), ct.c_int(x.shape[2]), None if ln is None else ln.ctypes.data_as(ct.POINTER(ct.c_int)),
                                      ct.c_int(1 if final else 0), ct.c_int(1 if synthesis else 0)))

    def stream_submit_i16(self, samples, lengths=None, final=False, synthesis=True):
        """The same for 16-bit PCM chunks: samples int16 [U][C][n]."""
        x = np.ascontiguousarray(samples, np.int16)
        assert x.ndim == 3 and x.shape[1] == self.C
        ln = None if lengths is None else np.ascontiguousarray(lengths, np.int32)
        _check(lib.btkb_stream_submit_i16(self._h, x.ctypes.data_as(ct.POINTER(ct.c_int16)), ct.c_int(x.shape[2]),
                                          None if ln is None else ln.ctypes.data_as(ct.POINTER(ct.c_int)), ct.c_int(1 if final else 0), ct.c_int(1 if synthesis else 0)))

    def stream_position(self):
        a, b = ct.c_int(0), ct.c_int(0)
        _check(lib.btkb_stream_position(self._h, ct.byref(a), ct.byref(b)))
        return a.value, b.value

    def reset(self):
        _check(lib.btkb_reset(self._h))

    def set_stream(self, cuda_stream):
        """cuda_stream: integer handle of a cudaStream_t (e.g. torch.cuda.Stream().cuda_stream), 0 / None = private stream."""
        _check(lib.btkb_set_stream(self._h, ct.c_void_p(int(cuda_stream) if cuda_stream else None)))

    # ---- results
    @property
    def num_frames(self):
        return int(lib.btkb_num_frames(self._h))

    def num_frames_of(self, u):
        return int(lib.btkb_num_frames_of(self._h, ct.c_int(u)))

    @property
    def num_blocks(self):
        return int(lib.btkb_num_blocks(self._h))

    def fetch_subband(self):
        out = np.empty((self.U, self.num_frames, self.K), np.complex64)
        _check(lib.btkb_fetch_subband(self._h, _fp(out)))
        return out

    def fetch_time(self):
        out = np.empty((self.U, self.num_blocks * self.D), np.float32)
        _check(lib.btkb_fetch_time(self._h, _fp(out)))
        return out

    def fetch_time_into(self, host_ptr):
        _check(lib.btkb_fetch_time(self._h, ct.cast(ct.c_void_p(host_ptr), ct.POINTER(ct.c_float))))

    def fetch_snapshots(self):
        out = np.empty((self.U, self.num_frames, self.C, self.K), np.complex64)
        _check(lib.btkb_fetch_snapshots(self._h, _fp(out)))
        return out

    def fetch_stats(self):
        out = np.zeros((self.U, 3), np.float64)
        _check(lib.btkb_fetch_stats(self._h, _dp(out)))
        return out

    def get_weights(self):
        out = np.empty((self.U, self.K, self.C), np.complex64)
        _check(lib.btkb_get_weights(self._h, _fp(out)))
        return out

    def get_active_weights(self):
        out = np.empty((self.U, self.K, self.C - self.NC), np.complex64)
        _check(lib.btkb_get_active_weights(self._h, _fp(out)))
        return out

    def get_sidelobe_weights(self):
        out = np.empty((self.U, self.K, self.C), np.complex64)
        _check(lib.btkb_get_sidelobe_weights(self._h, _fp(out)))
        return out

    def get_covariance(self):
        out = np.empty((self.U, self.K, self.C, self.C), np.complex64)
        _check(lib.btkb_get_covariance(self._h, _fp(out)))
        return out

    def get_postfilter_weights(self):
        out = np.empty((self.U, self.num_frames, self.K), np.float32)
        _check(lib.btkb_get_postfilter_weights(self._h, _fp(out)))
        return out

    def device_pointers(self):
        """(X, Y, time) device addresses of the resident snapshots [T][C][Gp] complex64, beamformed output [T][Gp] complex64 and
        resynthesised signal [U][nb D] float32, for zero-copy consumers (e.g. torch tensors built from raw pointers)."""
        X, Y, t = ct.c_void_p(), ct.c_void_p(), ct.c_void_p()
        _check(lib.btkb_device_pointers(self._h, ct.byref(X), ct.byref(Y), ct.byref(t)))
        return X.value, Y.value, t.value

    def last_timing(self):
        """dict(total_ms, analysis_ms, perbin_ms, synthesis_ms, launches) of the last run (CUDA events on the pipeline stream)."""
        out = (ct.c_float * 5)()
        _check(lib.btkb_last_timing(self._h, out))
        return dict(total_ms=out[0], analysis_ms=out[1], perbin_ms=out[2], synthesis_ms=out[3], launches=int(out[4]))
