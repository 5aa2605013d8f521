"""btk20.pybeamformer — the reference's pure-Python algorithm layer (btk20_src/lib/pybeamformer.py:41-1023) for the hot
path: delay calculators, GSC / MVDR wrappers, the NLMS GSC and the SMI-MVDR beamformer.  Same class and method names;
the per-bin Python loops of the reference (`__iter__` of SubbandGSCLMSBeamformer, `accu_stats_from_label`) run as CUDA
kernels behind native stream objects."""
import numpy

from .beamformer import SubbandGSCPtr, SubbandMVDRGSCPtr, SubbandGSCLMSPtr, LmsConfig, SubbandGSCRLSNativePtr, RlsConfig, SubbandSOSNativePtr

SSPEED = 343740.0


def calc_la_delays(mpos, azimuth, sspeed=SSPEED, ref_micx=None):
    """lib/pybeamformer.py:41-65."""
    chanN = len(mpos)
    if ref_micx is None:
        ref_micx = chanN // 2
    delays = numpy.array([-mpos[i][0] * numpy.cos(azimuth) / sspeed for i in range(chanN)], numpy.float64)
    return delays - delays[ref_micx]


def calc_pa_delays(mpos, azimuth, polar_angle, sspeed=SSPEED, ref_micx=None):
    """lib/pybeamformer.py:68-94."""
    chanN = len(mpos)
    if ref_micx is None:
        ref_micx = chanN // 2
    delays = numpy.zeros(chanN, numpy.float64)
    for i in range(chanN):
        dx = mpos[i][0] - mpos[ref_micx][0]
        dy = mpos[i][1] - mpos[ref_micx][1]
        delays[i] = -(dx * numpy.cos(azimuth) * numpy.sin(polar_angle) + dy * numpy.sin(azimuth) * numpy.sin(polar_angle)) / sspeed
    return delays


def calc_ca_delays(mpos, azimuth, polar_angle, sspeed=SSPEED):
    """lib/pybeamformer.py:97-120."""
    c_x = -numpy.sin(polar_angle) * numpy.cos(azimuth)
    c_y = -numpy.sin(polar_angle) * numpy.sin(azimuth)
    c_z = -numpy.cos(polar_angle)
    return numpy.array([(c_x * p[0] + c_y * p[1] + c_z * p[2]) / sspeed for p in mpos], numpy.float64)


def calc_nf_delays(mpos, x, y, z, sspeed=SSPEED, ref_micx=None):
    """lib/pybeamformer.py:123-139."""
    chanN = len(mpos)
    if ref_micx is None:
        ref_micx = chanN // 2
    delays = numpy.array([numpy.sqrt((x - p[0]) ** 2 + (y - p[1]) ** 2 + (z - p[2]) ** 2) / sspeed for p in mpos], numpy.float64)
    return delays - delays[ref_micx]


def calc_delays(array_type, mpos, position, sspeed=SSPEED, ref_micx=None):
    """lib/pybeamformer.py:142-153."""
    if array_type == 'linear':
        return calc_la_delays(mpos, position[0], sspeed=sspeed, ref_micx=ref_micx)
    elif array_type == 'planar':
        return calc_pa_delays(mpos, position[0], position[1], sspeed=sspeed, ref_micx=ref_micx)
    elif array_type == 'circular':
        return calc_ca_delays(mpos, position[0], position[1], sspeed=sspeed)
    return calc_nf_delays(mpos, position[0], position[1], position[2], sspeed=sspeed, ref_micx=ref_micx)


class SubbandBeamformer:
    """lib/pybeamformer.py:380-476 — base wrapper around a native beamformer stream."""

    def __init__(self, spec_sources):
        self._spec_sources = spec_sources
        self._chan_num = len(spec_sources)
        self._shiftlen = spec_sources[0].shiftlen()
        self._fftlen = spec_sources[0].size()
        self._fftlen2 = self._fftlen // 2
        for c in range(1, self._chan_num):
            assert self._shiftlen == spec_sources[c].shiftlen(), "%d-th channel: inconsistent shift length" % c
            assert self._fftlen == spec_sources[c].size(), "%d-th channel: inconsistent FFT length" % c
        self._beamformer = None
        self._waH = None
        self._Nc = 1

    def beamformer(self):
        return self._beamformer

    def native_stream(self):
        """The native complex stream behind this Python object (lets PyVectorComplexFeatureStreamPtr skip the per-frame
        C++ -> Python -> C++ crossing of unit_test/test_online_beamforming.py:128)."""
        return self._beamformer

    def spec_sources(self):
        return self._spec_sources

    def __iter__(self):
        if self._beamformer is None:
            raise NotImplementedError("Undefined beamformer object")
        while True:
            try:
                yield numpy.array(self._beamformer.next())
            except StopIteration:
                return

    def reset(self):
        if self._beamformer is None:
            raise NotImplementedError("Undefined beamformer object")
        self._beamformer.reset()

    def next_speaker(self):
        pass

    def chan_num(self):
        return self._chan_num

    def size(self):
        return self._fftlen

    def shiftlen(self):
        return self._shiftlen

    def set_active_weights(self):
        """lib/pybeamformer.py:464-475."""
        assert self._waH is not None, "The active weight vectors have to be set"
        for fbinX in range(self._fftlen2 + 1):
            packed_wa = numpy.zeros(2 * (self._chan_num - self._Nc), numpy.float64)
            packed_wa[0::2] = numpy.real(self._waH[fbinX])
            packed_wa[1::2] = numpy.imag(self._waH[fbinX])
            self._beamformer.set_active_weights_f(fbinX, packed_wa)


class SubbandGSCBeamformer(SubbandBeamformer):
    """lib/pybeamformer.py:478-535 — D&S / LCMV in GSC configuration with static active weights."""

    def __init__(self, spec_sources, Nc=1):
        SubbandBeamformer.__init__(self, spec_sources)
        self._beamformer = SubbandGSCPtr(fftlen=self._fftlen, half_band_shift=False)
        for source in self._spec_sources:
            self._beamformer.set_channel(source)
        self._Nc = Nc
        self._waH = numpy.zeros((self._fftlen, self._chan_num - self._Nc), numpy.complex128)

    def calc_beamformer_weights(self, samplerate, delays, update_active_weights=True):
        self._beamformer.calc_gsc_weights(samplerate, numpy.asarray(delays, numpy.float64))
        if update_active_weights:
            self.set_active_weights()
        self._wq = numpy.array([self._beamformer.get_weights(m) for m in range(self._fftlen2 + 1)], numpy.complex128)


    def calc_beamformer_weights_n(self, samplerate, delays_t, delays_js, update_active_weights=True):
        """lib/pybeamformer.py:516-535 — LCMV: one distortionless constraint plus Nc-1 nulls."""
        assert (self._Nc - 1) == len(delays_js), 'Mismatch between no. constraints and no. jammers'
        self._beamformer.calc_gsc_weights_n(samplerate, numpy.asarray(delays_t, numpy.float64), numpy.asarray(delays_js, numpy.float64), self._Nc)
        if update_active_weights:
            self.set_active_weights()
        self._wqH = numpy.conjugate(numpy.array([self._beamformer.get_weights(m) for m in range(self._fftlen2 + 1)], numpy.complex128))


class SubbandMVDRBeamformer(SubbandBeamformer):
    """lib/pybeamformer.py:538-585 — super-directive / MVDR beamformer."""

    def __init__(self, spec_sources, Nc=1):
        SubbandBeamformer.__init__(self, spec_sources)
        self._beamformer = SubbandMVDRGSCPtr(fftlen=self._fftlen, half_band_shift=False)
        for source in self._spec_sources:
            self._beamformer.set_channel(source)
        self._Nc = Nc
        self._waH = numpy.zeros((self._fftlen, self._chan_num - self._Nc), numpy.complex128)

    def calc_sd_beamformer_weights(self, samplerate, delays, mpos, sspeed=SSPEED, mu=0.01, update_active_weights=True):
        self._beamformer.calc_array_manifold_vectors(samplerate, numpy.asarray(delays, numpy.float64))
        self._beamformer.set_diffuse_noise_model(numpy.asarray(mpos, numpy.float64), samplerate, sspeed)
        self._beamformer.set_all_diagonal_loading(mu)
        self._beamformer.calc_mvdr_weights(samplerate, dthreshold=1.0E-8, calc_inverse_matrix=True)
        if update_active_weights:
            self.set_active_weights()
        self._wqH = numpy.conjugate(numpy.array([self._beamformer.mvdr_weights(m) for m in range(self._fftlen2 + 1)], numpy.complex128))


class SubbandGSCLMSBeamformer(SubbandBeamformer):
    """lib/pybeamformer.py:588-762 — leaky power-normalised LMS GSC.  The reference runs the per-bin update as a Python
    loop inside __iter__; here the same recurrence is the fused per-bin CUDA kernel (csrc/btkb_perbin.cu)."""

    def __init__(self, spec_sources, beta=0.97, gamma=0.01, init_diagonal_load=1.0E+6, regularization_param=1.0E-4, energy_floor=90,
                 sil_thresh=1.0E+8, max_wa_l2norm=100.0, min_frames=128, slowdown_after=4096, Nc=1):
        SubbandBeamformer.__init__(self, spec_sources)
        if Nc != 1:
            raise NotImplementedError("the GPU NLMS implements one linear constraint (Nc = 1)")
        self._Nc = Nc
        cfg = LmsConfig()
        cfg.beta = beta; cfg.gamma = gamma; cfg.init_diagonal_load = init_diagonal_load; cfg.regularization_param = regularization_param
        cfg.energy_floor = energy_floor; cfg.sil_thresh = sil_thresh; cfg.max_wa_l2norm = max_wa_l2norm
        cfg.min_frames = min_frames; cfg.slowdown_after = slowdown_after
        self._beamformer = SubbandGSCLMSPtr(self._fftlen, cfg)
        for source in self._spec_sources:
            self._beamformer.set_channel(source)

    def calc_beamformer_weights(self, samplerate, delays):
        """lib/pybeamformer.py:736-743."""
        self._beamformer.calc_beamformer_weights(samplerate, numpy.asarray(delays, numpy.float64))

    def reset_stats(self):
        """lib/pybeamformer.py:745-757: adaptive state restarts with every reset() of the native stream."""
        self._beamformer.reset()

    def active_weights(self):
        """waH[K][C-Nc] after the run (the reference's self._waH)."""
        return numpy.array(self._beamformer.active_weights(), numpy.complex128)

    def total_updates(self):
        return self._beamformer.total_updates()


class SubbandGSCRLSBeamformer(SubbandBeamformer):
    """lib/pybeamformer.py:765-928 — regularised RLS sidelobe canceller in GSC form.  The reference's per-bin Python loop
    (precision-matrix update, quadratic constraint, norm reset) is the fused CUDA kernel k_perbin_rls (csrc/btkb_perbin.cu)."""

    def __init__(self, spec_sources, beta=0.97, gamma=0.04, mu=0.97, init_diagonal_load=1.0E+6, regularization_param=1.0E-2, sil_thresh=1.0E+8,
                 constraint_option=3, alpha2=10.0, max_wa_l2norm=100.0, min_frames=128, slowdown_after=4096, Nc=1):
        SubbandBeamformer.__init__(self, spec_sources)
        if Nc != 1:
            raise NotImplementedError("the GPU RLS implements one linear constraint (Nc = 1)")
        self._Nc = Nc
        cfg = RlsConfig()
        cfg.beta = beta; cfg.gamma = gamma; cfg.mu = mu; cfg.init_diagonal_load = init_diagonal_load
        cfg.regularization_param = regularization_param; cfg.sil_thresh = sil_thresh; cfg.constraint_option = constraint_option
        cfg.alpha2 = alpha2; cfg.max_wa_l2norm = max_wa_l2norm; cfg.min_frames = min_frames
        self._slowdown_after = slowdown_after   # accepted and, as in the reference's loop, never read (pybeamformer.py:817-901)
        self._beamformer = SubbandGSCRLSNativePtr(self._fftlen, cfg)
        for source in self._spec_sources:
            self._beamformer.set_channel(source)

    def calc_beamformer_weights(self, samplerate, delays):
        """lib/pybeamformer.py:903-911."""
        self._beamformer.calc_beamformer_weights(samplerate, numpy.asarray(delays, numpy.float64))

    def reset_stats(self):
        """lib/pybeamformer.py:913-925."""
        self._beamformer.reset()

    def active_weights(self):
        return numpy.array(self._beamformer.active_weights(), numpy.complex128)

    def total_updates(self):
        return self._beamformer.total_updates()


class SubbandSMIMVDRBeamformer(SubbandMVDRBeamformer):
    """lib/pybeamformer.py:931-1023 — MVDR by sample matrix inversion from VAD-labelled noise frames."""

    def __init__(self, spec_sources, Nc=1):
        SubbandMVDRBeamformer.__init__(self, spec_sources, Nc)
        self._have_stats = False

    def accu_stats_from_label(self, samplerate, target_labs=[(0.1, -1)], energy_threshold=10):
        """lib/pybeamformer.py:948-992 for one target segment per call (the shipped configs use one, confs/smimvdr.json)."""
        if len(target_labs) != 1:
            raise NotImplementedError("one VAD segment per utterance")
        self._beamformer.accumulate_noise_covariance(samplerate, float(target_labs[0][0]), float(target_labs[0][1]), float(energy_threshold))
        self._have_stats = True

    def finalize_stats(self):
        """lib/pybeamformer.py:994-1000 (the division by the frame count happens on the GPU)."""
        assert self._have_stats, "No noise stats accumulated; Use self.accu_stats_from_label()"

    def calc_beamformer_weights(self, samplerate, delays, mu=1e-4, update_active_weights=True):
        """lib/pybeamformer.py:1002-1023."""
        self._beamformer.calc_array_manifold_vectors(samplerate, numpy.asarray(delays, numpy.float64))
        self._beamformer.set_all_diagonal_loading(mu)
        self._beamformer.calc_mvdr_weights(samplerate, dthreshold=1.0E-8, calc_inverse_matrix=True)
        if update_active_weights:
            self.set_active_weights()
        self._wqH = numpy.conjugate(numpy.array([self._beamformer.mvdr_weights(m) for m in range(self._fftlen2 + 1)], numpy.complex128))


class SubbandSOSBatchBeamformer(SubbandBeamformer):
    """lib/pybeamformer.py:1026-1219 — batch beamformer driven by second-order statistics.  The reference accumulates
    x x^H per bin per frame in a Python loop and applies wqH in another; here both are CUDA kernels (csrc/btkb_sos.cu,
    k_perbin) behind one native stream object."""

    def __init__(self, spec_sources):
        SubbandBeamformer.__init__(self, spec_sources)
        self._beamformer = SubbandSOSNativePtr(self._fftlen)
        for source in self._spec_sources:
            self._beamformer.set_channel(source)
        self._isamp = 0
        self._have_stats = False

    def accu_stats_from_label(self, samplerate, target_labs=[(0.1, -1)], energy_threshold=10):
        """lib/pybeamformer.py:1063-1127 (several segments allowed; the segment cursor walks like the reference's `labx`)."""
        self._beamformer.accu_stats_from_label(samplerate, numpy.asarray(target_labs, numpy.float64).reshape(-1, 2), float(energy_threshold))
        self._have_stats = True

    def accu_stats_from_tfmask(self, samplerate, mask_t, mask_j, energy_threshold=10):
        """lib/pybeamformer.py:1129-1183."""
        self._beamformer.accu_stats_from_tfmask(samplerate, numpy.asarray(mask_t, numpy.float32), numpy.asarray(mask_j, numpy.float32), float(energy_threshold))
        self._have_stats = True

    def finalize_stats(self):
        pass

    def reset_stats(self):
        """lib/pybeamformer.py:1209-1213."""
        self._beamformer.reset_stats()
        self._have_stats = False

    def frame_counts(self):
        """(_target_frame_counts, _noise_frame_counts) of the reference, [K] each."""
        c = numpy.array(self._beamformer.frame_counts())
        return c[:, 0], c[:, 1]

    def reset(self):
        self._beamformer.reset()
        self._isamp = 0

    def _export_wqH(self):
        self._wqH = numpy.conjugate(numpy.array([self._beamformer.get_weights(m) for m in range(self._fftlen2 + 1)], numpy.complex128))


class SubbandBlindMVDRBeamformer(SubbandSOSBatchBeamformer):
    """lib/pybeamformer.py:1243-1295 — MVDR without a look direction (MMSE beamformer)."""

    def finalize_stats(self, gamma=1e-6):
        """lib/pybeamformer.py:1283-1295: the normalisation and the diagonal loading run on the GPU inside calc_beamformer_weights."""
        assert self._have_stats, "No target signal stats accumulated; Use self.accu_stats_from_label() or accu_stats_from_tfmask()"
        self._gamma = gamma

    def calc_beamformer_weights(self, ref_micx=0, offset=0.0):
        """lib/pybeamformer.py:1257-1281."""
        if not self._have_stats:
            raise RuntimeError('No target signal SOS')
        assert offset >= 0 and offset <= 1, "The offset value %f is out of [0, 1]" % (offset)
        try:
            self._beamformer.calc_weights(0, getattr(self, "_gamma", 1e-6), ref_micx, offset)
        except Exception as e:
            if "Matrix inversion failed" in str(e):
                raise ArithmeticError(str(e))
            if "stats accumulated" in str(e):
                raise AssertionError(str(e))
            raise
        self._export_wqH()


class SubbandGEVBeamformer(SubbandBlindMVDRBeamformer):
    """lib/pybeamformer.py:1298-1357 — generalised-eigenvector beamformer (the principal eigenvector's phase is fixed up to one
    global sign per utterance, see include/btkb.h BTKB_SOS_GEV)."""

    def calc_beamformer_weights(self):
        if not self._have_stats:
            raise RuntimeError('No target signal SOS')
        try:
            self._beamformer.calc_weights(1, getattr(self, "_gamma", 1e-6), 0, 0.0)
        except Exception as e:
            if "GEV failed" in str(e):
                raise ArithmeticError(str(e))
            if "stats accumulated" in str(e):
                raise AssertionError(str(e))
            raise
        self._export_wqH()
