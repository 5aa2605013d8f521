"""The reference's Python surface (btk20.*) over the C++ host mirror.  CPU part: stream protocol, SampleFeature block
logic (feature/feature.cc:605-649), wav reading, exception mapping (include/jexception.i:20-86).  GPU part (-m gpu): the
front-end script flow of unit_test/test_online_beamforming.py:51-228 / test_sos_batch_beamforming.py:95-233 reproduced
with the same calls and compared with the reference goldens."""
import os
import struct
import wave

import numpy as np
import pytest

from conftest import load_golden, rel_l2

btk20 = pytest.importorskip("distant_speech_recognition_b200.btk20")
from distant_speech_recognition_b200.btk20.feature import SampleFeaturePtr  # noqa: E402
from distant_speech_recognition_b200.btk20.modulated import OverSampledDFTAnalysisBankPtr, OverSampledDFTSynthesisBankPtr, get_window  # noqa: E402
from distant_speech_recognition_b200.btk20.beamformer import SubbandDSPtr, SubbandGSCPtr, SnapShotArrayPtr, calc_all_delays  # noqa: E402
from distant_speech_recognition_b200.btk20.postfilter import ZelinskiPostFilterPtr, McCowanPostFilterPtr, LefkimmiatisPostFilterPtr  # noqa: E402
from distant_speech_recognition_b200.btk20.stream import PyVectorComplexFeatureStreamPtr  # noqa: E402
from distant_speech_recognition_b200.btk20 import pybeamformer  # noqa: E402

FS = 16000


def test_sample_feature_blocks_and_padding():
    s = SampleFeaturePtr(block_len=128, shift_len=128, pad_zeros=True)
    x = np.arange(300, dtype=np.float64)
    s.setSamples(x, FS)
    assert s.samplesN() == 300 and s.size() == 128
    blocks = [np.array(b) for b in s]            # __iter__ = reset(); return self  (feature.i)
    assert len(blocks) == 3
    assert np.array_equal(blocks[0], x[:128]) and np.array_equal(blocks[2][:44], x[256:]) and np.all(blocks[2][44:] == 0)
    assert s.is_end() is False or True
    with pytest.raises(StopIteration):            # jiterator_error -> StopIteration
        s.next()
    s.reset()
    b0 = np.array(s.next(0)); again = np.array(s.next(0))   # frame_no == frame_no_ -> cached frame
    assert np.array_equal(b0, again) and s.frame_no() == 0
    with pytest.raises(IndexError):               # jindex_error: non-consecutive frame number (feature.cc:614-616)
        s.next(5)
    # without zero padding the last partial block ends the stream
    s2 = SampleFeaturePtr(block_len=128, shift_len=128, pad_zeros=False); s2.setSamples(x, FS)
    assert len([1 for _ in s2]) == 2


def test_sample_feature_reads_wav(tmp_path):
    path = os.path.join(tmp_path, "t.wav")
    data = (np.arange(-500, 500) * 7).astype(np.int16)
    with wave.open(path, "wb") as w:
        w.setnchannels(1); w.setsampwidth(2); w.setframerate(FS); w.writeframes(data.tobytes())
    s = SampleFeaturePtr(block_len=256, shift_len=256, pad_zeros=True)
    n = s.read(path, FS)                          # test_online_beamforming.py:83 call shape
    assert n == 1000 and np.array_equal(np.array(s.data()), data.astype(np.float32))   # norm=0: raw int16-scale floats
    with pytest.raises(IOError):
        s.read(os.path.join(tmp_path, "missing.wav"), FS)
    # 32-bit PCM: libsndfile hands the integers out unscaled when normalisation is off (norm = 0, the scripts' call), feature.cc:265-270
    path32 = os.path.join(tmp_path, "t32.wav")
    d32 = (np.arange(-300, 300, dtype=np.int64) * 70001).astype(np.int32)
    with wave.open(path32, "wb") as w:
        w.setnchannels(1); w.setsampwidth(4); w.setframerate(FS); w.writeframes(d32.tobytes())
    s32 = SampleFeaturePtr(block_len=256, shift_len=256, pad_zeros=True)
    assert s32.read(path32, FS) == 600 and np.array_equal(np.array(s32.data()), d32.astype(np.float32))


def test_constructor_checks_and_misc(protos):
    h, g = protos[256]
    s = SampleFeaturePtr(block_len=100, shift_len=100, pad_zeros=True)
    with pytest.raises(Exception):                # jdimension_error: Input block length != D (modulated.cc:337-338)
        OverSampledDFTAnalysisBankPtr(s, prototype=h, M=256, m=4, r=1, delay_compensation_type=2)
    s = SampleFeaturePtr(block_len=128, shift_len=128, pad_zeros=True)
    with pytest.raises(Exception):                # jconsistency_error: prototype size (modulated.cc:239-241)
        OverSampledDFTAnalysisBankPtr(s, prototype=h[:100], M=256, m=4, r=1, delay_compensation_type=2)
    afb = OverSampledDFTAnalysisBankPtr(s, prototype=h, M=256, m=4, r=1, delay_compensation_type=2)
    assert afb.fftlen() == 256 and afb.shiftlen() == 128 and afb.size() == 256
    assert afb.polyphase(3, 2) == h[3 + 256 * 2]
    ds = SubbandDSPtr(fftlen=256, half_band_shift=False)
    ds.set_channel(afb)
    with pytest.raises(Exception):                # delays/channels mismatch (beamformer.cc:504-506)
        ds.calc_array_manifold_vectors(FS, np.zeros(3))
    snap = SnapShotArrayPtr(8, 2)
    snap.set_samples(np.arange(8) + 1j, 0); snap.set_samples(np.arange(8) * 2.0, 1); snap.update()
    assert np.allclose(np.array(snap.snapshot(3)), [3 + 1j, 6])
    assert np.allclose(get_window(2, 5), 0.5 * (1 - np.cos(2 * np.pi * np.arange(5) / 4)))
    d = calc_all_delays(0, 0, 0, np.array([[0.0, 0, 0], [30.0, 40.0, 0], [0, 0, 100.0]]))
    assert np.allclose(d, (np.array([0.0, 50.0, 100.0]) - 50.0) / 343740.0)
    mpos = [[-113.0, 0, 2], [36.0, 0, 2], [76.0, 0, 2], [113.0, 0, 2]]
    from oracle import restate
    assert np.allclose(pybeamformer.calc_delays("linear", mpos, [-1.306379, None, None]), restate.calc_la_delays(mpos, -1.306379))
    assert np.allclose(pybeamformer.calc_nf_delays(mpos, 10.0, 500.0, 3.0), restate.calc_nf_delays(mpos, 10.0, 500.0, 3.0))


def test_dereverberation_constructor_checks(protos):
    """btk20.dereverberation argument checks that need no GPU (dereverberation.cc:365-373, 393-399)."""
    from distant_speech_recognition_b200.btk20.dereverberation import (SingleChannelWPEDereverberationFeaturePtr, MultiChannelWPEDereverberationPtr,
                                                                        MultiChannelWPEDereverberationFeaturePtr)
    h, g = protos[256]
    with pytest.raises(Exception, match="Nyquist"):
        MultiChannelWPEDereverberationPtr(subbands_num=256, channels_num=2, band_width=9000.0, samplerate=FS)
    pre = MultiChannelWPEDereverberationPtr(subbands_num=256, channels_num=2, lower_num=0, upper_num=4)
    assert pre.size() == 256
    sf = SampleFeaturePtr(block_len=128, shift_len=128, pad_zeros=True); sf.setSamples(np.zeros(1000), FS)
    afb = OverSampledDFTAnalysisBankPtr(sf, prototype=h, M=256, m=4, r=1, delay_compensation_type=2)
    pre.set_input(afb); pre.set_input(afb)
    with pytest.raises(MemoryError):                      # jallocation_error "Channel capacity exceeded."
        pre.set_input(afb)
    feat = MultiChannelWPEDereverberationFeaturePtr(pre, channel_no=1)
    assert feat.size() == 256 and feat.shiftlen() == 128 and feat.frame_no() == -1
    with pytest.raises(Exception, match="estimate_filter"):
        feat.next()                                       # jinitialization_error before estimate_filter()
    single = SingleChannelWPEDereverberationFeaturePtr(afb, lower_num=0, upper_num=8)
    with pytest.raises(Exception, match="estimate_filter"):
        single.next()


def _afbs(x, h, M, D):
    afbs = []
    for c in range(x.shape[0]):
        sf = SampleFeaturePtr(block_len=D, shift_len=D, pad_zeros=True)
        sf.setSamples(x[c].astype(np.float64), FS)
        afbs.append(OverSampledDFTAnalysisBankPtr(sf, prototype=h, M=M, m=4, r=1, delay_compensation_type=2))
    return afbs


@pytest.mark.gpu
def test_frontend_flow_gsclms(protos):
    """unit_test/test_online_beamforming.py with "type":"gsclms": afbs -> SubbandGSCLMSBeamformer -> PyVectorComplexFeatureStreamPtr
    -> OverSampledDFTSynthesisBankPtr; for buf in sfb."""
    g = load_golden("gsclms_c8_m512"); h, gg = protos[512]; M, D = 512, 256
    afbs = _afbs(g["x"], h, M, D)
    bf = pybeamformer.SubbandGSCLMSBeamformer(afbs, min_frames=int(g["min_frames"]))
    bf.calc_beamformer_weights(FS, g["delays"])
    spatial_filter = PyVectorComplexFeatureStreamPtr(bf)
    sfb = OverSampledDFTSynthesisBankPtr(spatial_filter, prototype=gg, M=M, m=4, r=1, delay_compensation_type=2)
    out, total_energy = [], 0.0
    for frame_no, buf in enumerate(sfb):
        total_energy += np.inner(buf, buf); out.append(np.array(buf))
    y = np.concatenate(out)
    assert y.shape == g["time"].shape and rel_l2(y, g["time"]) < 1e-4
    assert abs(total_energy - g["stats"][0]) < 1e-4 * g["stats"][0]
    assert bf.total_updates() == g["stats"][2]
    assert rel_l2(bf.active_weights(), g["waH"]) < 1e-3
    # the subband stream itself, with Hermitian fill, and idempotent re-read
    bf.reset()
    Y = np.array([np.array(v) for v in bf])
    assert Y.shape == (g["Y"].shape[0], M) and rel_l2(Y[:, :257], g["Y"]) < 1e-4
    assert np.allclose(Y[:, 300], np.conj(Y[:, M - 300]))
    # second utterance through the same graph: reset() rewinds everything and restarts adaptation
    sfb.reset()
    y2 = np.concatenate([np.array(b) for b in sfb])
    assert np.array_equal(y, y2)


@pytest.mark.gpu
def test_frontend_flow_gscrls(protos):
    """unit_test/test_online_beamforming.py with "type":"gscrls" (confs/gscrls.json): SubbandGSCRLSBeamformer in the same graph."""
    g = load_golden("gscrls_c8_m512"); h, gg = protos[512]; M, D = 512, 256
    afbs = _afbs(g["x"], h, M, D)
    bf = pybeamformer.SubbandGSCRLSBeamformer(afbs, beta=0.97, gamma=0.04, mu=0.97, init_diagonal_load=1.0E+6, regularization_param=1.0E-2,
                                              sil_thresh=1.0E+8, constraint_option=3, alpha2=10.0, max_wa_l2norm=100.0,
                                              min_frames=int(g["min_frames"]), slowdown_after=4096)
    bf.calc_beamformer_weights(FS, g["delays"])
    sfb = OverSampledDFTSynthesisBankPtr(PyVectorComplexFeatureStreamPtr(bf), prototype=gg, M=M, m=4, r=1, delay_compensation_type=2)
    y = np.concatenate([np.array(b) for b in sfb])
    assert y.shape == g["time"].shape and rel_l2(y, g["time"]) < 1e-4
    assert bf.total_updates() == int(g["n_updates"])
    assert rel_l2(bf.active_weights(), g["waH"]) < 1e-3


@pytest.mark.gpu
def test_frontend_flow_gsc_zelinski(protos):
    """D&S/GSC + Zelinski (confs/ds_and_zelinski.json flow): ZelinskiPostFilterPtr(pybf, M, alpha, subtype); set_beamformer."""
    g = load_golden("gsc_zelinski_c8_m512"); h, gg = protos[512]; M, D = 512, 256
    afbs = _afbs(g["x"], h, M, D)
    bf = pybeamformer.SubbandGSCBeamformer(afbs, Nc=1)
    bf._waH[:257] = g["wa"]
    bf.calc_beamformer_weights(FS, g["delays"])
    pybf = PyVectorComplexFeatureStreamPtr(bf)
    pf = ZelinskiPostFilterPtr(pybf, M, 0.7, 2)
    pf.set_beamformer(bf.beamformer())
    sfb = OverSampledDFTSynthesisBankPtr(pf, prototype=gg, M=M, m=4, r=1, delay_compensation_type=2)
    y = np.concatenate([np.array(b) for b in sfb])
    assert rel_l2(y, g["time"]) < 1e-4
    pf.reset()
    Y = np.array([np.array(v) for v in pf])
    assert rel_l2(Y[:, :257], g["Y"]) < 1e-4
    w = np.real(np.array(pf.postfilter_weights()))
    assert w[:257].min() >= 1e-4 - 1e-9 and w.max() <= 1.0
    assert rel_l2(np.array(bf.beamformer().get_weights(17)), g["wq"][17]) < 1e-6


@pytest.mark.gpu
def test_frontend_flow_lcmv_zelinski(protos):
    """confs/lcmv_and_zelinski.json flow (test_online_beamforming.py:92-100,183): SubbandGSCBeamformer(afbs, Nc=2),
    calc_beamformer_weights_n(samplerate, delays_t, delays_js), Zelinski post-filter, synthesis; weights vs the reference's
    golden (calcMainlobe2), output vs the fp64 restatement."""
    from oracle import restate
    from distant_speech_recognition_b200 import synthetic
    g = load_golden("lcmv"); h, gg = protos[512]; M, D, C = 512, 256, 8
    x, _, _, _ = synthetic.make_utterance(31, C, 9000)
    afbs = _afbs(x, h, M, D)
    bf = pybeamformer.SubbandGSCBeamformer(afbs, Nc=2)
    with pytest.raises(AssertionError):
        bf.calc_beamformer_weights_n(FS, g["dT"], [g["dJ1"], g["dJ2"]])      # Nc - 1 jammers expected
    bf.calc_beamformer_weights_n(FS, g["dT"], [g["dJ1"]])
    assert rel_l2(np.conj(bf._wqH), g["w2"]) < 1e-6
    pf = ZelinskiPostFilterPtr(PyVectorComplexFeatureStreamPtr(bf), M, 0.7, 2)
    pf.set_beamformer(bf.beamformer())
    sfb = OverSampledDFTSynthesisBankPtr(pf, prototype=gg, M=M, m=4, r=1, delay_compensation_type=2)
    y = np.concatenate([np.array(b) for b in sfb])
    X = np.stack([restate.analysis(x[c], h, M, 4, 1) for c in range(C)], axis=1)
    w2 = np.concatenate([g["w2"], np.conj(g["w2"][1:256][::-1])])          # full-band quiescent vectors (Hermitian mirror)
    Yo = restate.subband_gsc(X, w2, np.zeros_like(w2))
    ta = restate.calc_mainlobe(M, C, FS, g["dT"])                          # the post-filter aligns with the D&S manifold
    Yz, _ = restate.zelinski_postfilter(Yo, X, ta, 0.7, 2, 0)
    assert rel_l2(y, restate.synthesis(Yz, gg, M, 4, 1)) < 1e-4


@pytest.mark.gpu
def test_frontend_flow_sd_mccowan_and_ds_lefkimmiatis(protos):
    """unit_test/test_online_beamforming.py:132-156,204 with confs/sd_and_mccowan.json and sd_and_lefkimmiatis.json parameters."""
    g = load_golden("mccowan_c4_m256"); h, gg = protos[256]; M, D = 256, 128
    afbs = _afbs(g["x"], h, M, D)
    bf = pybeamformer.SubbandMVDRBeamformer(afbs)
    bf.calc_sd_beamformer_weights(FS, g["delays"], g["mpos"], mu=float(g["mu"]))
    pf = McCowanPostFilterPtr(PyVectorComplexFeatureStreamPtr(bf), M, 0.7, 2)
    with pytest.raises(Exception, match="noise coherence"):      # j_error: "Construct/set first a noise coherence matrix" (postfilter.cc:631-633)
        pf.set_all_diagonal_loading(0.01)
    assert pf.set_diffuse_noise_model(g["mpos"], FS, 343740.0)
    pf.set_all_diagonal_loading(0.01)
    R5 = np.array(pf.noise_spatial_spectral_matrix(5))
    assert R5.shape == (4, 4) and abs(R5[0, 0] - (1.0 + float(np.float32(0.01)))) < 1e-12 and abs(R5[0, 1] - R5[1, 0]) < 1e-15
    pf.set_beamformer(bf.beamformer())
    sfb = OverSampledDFTSynthesisBankPtr(pf, prototype=gg, M=M, m=4, r=1, delay_compensation_type=2)
    y = np.concatenate([np.array(b) for b in sfb])
    assert rel_l2(y, g["timea"]) < 1e-4
    pf.reset()
    Y = np.array([np.array(v) for v in pf])
    assert rel_l2(Y[:, :129], g["Ya"]) < 1e-4
    assert np.all(Y[0, 129:] == 0) and np.allclose(Y[2, 129:], np.conj(Y[2, 1:128][::-1]))   # frame 0: upper half left at zero (postfilter.cc:896-901)
    # per-bin setters round-trip through the device store
    pf.set_noise_spatial_spectral_matrix(7, np.eye(4) * 2.0)
    pf.set_diagonal_looading(7, 0.5); pf.divide_nondiagonal_elements(5, 1.0)
    assert np.allclose(np.array(pf.noise_spatial_spectral_matrix(7)), np.eye(4) * 2.5)
    assert abs(np.array(pf.noise_spatial_spectral_matrix(5))[0, 1] - R5[0, 1] / 2.0) < 1e-15

    g = load_golden("lefkimmiatis_c8_m512"); h, gg = protos[512]; M, D = 512, 256
    afbs = _afbs(g["x"], h, M, D)
    bf = pybeamformer.SubbandGSCBeamformer(afbs, Nc=1)   # 'delay_and_sum' type: zero active weights
    bf.calc_beamformer_weights(FS, g["delays"])
    pf = LefkimmiatisPostFilterPtr(PyVectorComplexFeatureStreamPtr(bf), M, 1e-4, 100, 0.8, 2)
    pf.set_diffuse_noise_model(g["mpos"], FS, 343740.0)
    pf.set_all_diagonal_loading(0.1)
    pf.calc_inverse_noise_spatial_spectral_matrix()
    pf.set_beamformer(bf.beamformer())
    sfb = OverSampledDFTSynthesisBankPtr(pf, prototype=gg, M=M, m=4, r=1, delay_compensation_type=2)
    y = np.concatenate([np.array(b) for b in sfb])
    assert rel_l2(y, g["timea"]) < 1e-4


@pytest.mark.gpu
def test_frontend_flow_smimvdr(protos):
    """unit_test/test_sos_batch_beamforming.py:186-223 with "type":"smimvdr"."""
    g = load_golden("smimvdr_zelinski_c8_m512"); h, gg = protos[512]; M, D = 512, 256
    afbs = _afbs(g["x"], h, M, D)
    bf = pybeamformer.SubbandSMIMVDRBeamformer(afbs, Nc=1)
    bf.accu_stats_from_label(FS, target_labs=[(0.25, 0.75)], energy_threshold=10)
    bf.finalize_stats()
    bf.calc_beamformer_weights(FS, g["delays"], mu=float(g["mu"]))
    pf = ZelinskiPostFilterPtr(PyVectorComplexFeatureStreamPtr(bf), M, 0.7, 2)
    pf.set_beamformer(bf.beamformer())
    sfb = OverSampledDFTSynthesisBankPtr(pf, prototype=gg, M=M, m=4, r=1, delay_compensation_type=2)
    y = np.concatenate([np.array(b) for b in sfb])
    assert rel_l2(y, g["time"]) < 3e-4      # bounded by the reference's float-SVD noise (tests/test_parity_gpu.py)


@pytest.mark.gpu
def test_frontend_flow_bmvdr_and_gev(protos):
    """unit_test/test_sos_batch_beamforming.py:186-233 with "type":"bmvdr" (VAD label) and "type":"gev" (TF masks)."""
    g = load_golden("bmvdr_vad_c8_m512"); h, gg = protos[512]; M, D = 512, 256
    afbs = _afbs(g["x"], h, M, D)
    bf = pybeamformer.SubbandBlindMVDRBeamformer(afbs)
    with pytest.raises(RuntimeError):
        bf.calc_beamformer_weights()
    bf.accu_stats_from_label(FS, target_labs=[tuple(r) for r in g["labels"]], energy_threshold=10)
    bf.finalize_stats(gamma=float(g["gamma"]))
    bf.calc_beamformer_weights(ref_micx=int(g["ref_micx"]), offset=float(g["offset"]))
    ct, cn = bf.frame_counts()
    assert np.array_equal(ct, g["ct"]) and np.array_equal(cn, g["cn"])
    assert rel_l2(np.conj(bf._wqH), g["w"]) < 1e-4
    sfb = OverSampledDFTSynthesisBankPtr(PyVectorComplexFeatureStreamPtr(bf), prototype=gg, M=M, m=4, r=1, delay_compensation_type=2)
    y = np.concatenate([np.array(b) for b in sfb])
    assert rel_l2(y, g["time"]) < 1e-4

    g = load_golden("gev_tfmask_c4_m256"); h, gg = protos[256]; M, D = 256, 128
    afbs = _afbs(g["x"], h, M, D)
    bf = pybeamformer.SubbandGEVBeamformer(afbs)
    bf.accu_stats_from_tfmask(FS, g["mask_t"], g["mask_j"], energy_threshold=10)
    bf.finalize_stats(gamma=float(g["gamma"]))
    bf.calc_beamformer_weights()
    w = np.conj(bf._wqH)
    sgn = np.sign(np.real(np.vdot(w[0], g["w"][0])))     # one global sign is LAPACK-defined in the reference
    assert rel_l2(sgn * w, g["w"]) < 1e-4
    sfb = OverSampledDFTSynthesisBankPtr(PyVectorComplexFeatureStreamPtr(bf), prototype=gg, M=M, m=4, r=1, delay_compensation_type=2)
    y = np.concatenate([np.array(b) for b in sfb])
    assert rel_l2(sgn * y, g["time"]) < 1e-4


@pytest.mark.gpu
def test_frontend_flow_wpe_single_and_multi_channel(protos):
    """unit_test/test_subband_dereverberator.py:53-170: single-channel WPE feature into the synthesis bank; multi-channel
    estimator + one feature stream per channel; the audio is re-read between estimate_filter() and the output pass."""
    from distant_speech_recognition_b200.btk20.dereverberation import (SingleChannelWPEDereverberationFeaturePtr, MultiChannelWPEDereverberationPtr,
                                                                        MultiChannelWPEDereverberationFeaturePtr)
    g = load_golden("wpe_single_m256"); h, gg = protos[256]; M, D = 256, 128
    sf = SampleFeaturePtr(block_len=D, shift_len=D, pad_zeros=True); sf.setSamples(g["x"][0].astype(np.float64), FS)
    afb = OverSampledDFTAnalysisBankPtr(sf, prototype=h, M=M, m=4, r=1, delay_compensation_type=2)
    dereverb = SingleChannelWPEDereverberationFeaturePtr(afb, lower_num=0, upper_num=16, iterations_num=2, load_db=-20.0, band_width=0.0, samplerate=FS)
    sfb = OverSampledDFTSynthesisBankPtr(dereverb, prototype=gg, M=M, m=4, r=1, delay_compensation_type=2)
    with pytest.raises(Exception):
        dereverb.next()                                   # jinitialization_error before estimate_filter()
    dereverb.print_objective_func(50)
    assert dereverb.estimate_filter() == int(g["used_a"])
    sf.setSamples(g["x"][0].astype(np.float64), FS)       # "sample_feat.read(...)" again
    y = np.concatenate([np.array(b) for b in sfb])
    assert rel_l2(y, g["time_a"]) < 1e-4
    X = np.array([np.array(v) for v in dereverb])          # the stream itself, after reset
    assert rel_l2(X[:, :129], g["Xa"]) < 1e-4 and np.allclose(X[:, 129:], np.conj(X[:, 1:128][:, ::-1]))

    g = load_golden("wpe_c4_m256"); x = g["x"]
    pre = MultiChannelWPEDereverberationPtr(subbands_num=M, channels_num=4, lower_num=2, upper_num=8, iterations_num=3, load_db=-20.0, band_width=3000.0,
                                            diagonal_bias=1e-3, samplerate=FS)
    sfs, afbs = [], []
    for c in range(4):
        s_ = SampleFeaturePtr(block_len=D, shift_len=D, pad_zeros=True); s_.setSamples(x[c].astype(np.float64), FS)
        a_ = OverSampledDFTAnalysisBankPtr(s_, prototype=h, M=M, m=4, r=1, delay_compensation_type=2)
        pre.set_input(a_); sfs.append(s_); afbs.append(a_)
    with pytest.raises(MemoryError):
        pre.set_input(afbs[0])                             # jallocation_error "Channel capacity exceeded."
    assert pre.estimate_filter(2, 42) == int(g["used_b"])
    feats = []
    for c in range(4):
        sfs[c].setSamples(x[c].astype(np.float64), FS)
        feats.append(MultiChannelWPEDereverberationFeaturePtr(pre, channel_no=c))
    out = [[] for _ in range(4)]
    while True:
        try:
            for c in range(4):
                out[c].append(np.array(feats[c].next()))
        except StopIteration:
            break
    Xd = np.stack([np.array(o) for o in out], axis=1)      # [T][C][M]
    assert rel_l2(Xd[:, :, :129], g["Xb"]) < 1e-4


@pytest.mark.gpu
def test_frontend_flow_wpe_into_gsclms(protos):
    """configs[4] chain at script level: analysis banks -> MultiChannelWPEDereverberation -> one feature per channel ->
    SubbandGSCLMSBeamformer -> synthesis bank, against the fp64 restatement of the same chain."""
    from oracle import restate
    from distant_speech_recognition_b200 import synthetic
    from distant_speech_recognition_b200.btk20.dereverberation import MultiChannelWPEDereverberationPtr, MultiChannelWPEDereverberationFeaturePtr
    M, D, C, n = 256, 128, 4, 7000
    h, gg = protos[M]
    x, d, _, _ = synthetic.make_utterance(55, C, n)
    wpe = dict(lower_num=1, upper_num=6, iterations_num=2, load_db=-30.0, band_width=0.0, diagonal_bias=1e-4)
    pre = MultiChannelWPEDereverberationPtr(subbands_num=M, channels_num=C, samplerate=FS, **wpe)
    afbs = _afbs(x, h, M, D)
    for a_ in afbs:
        pre.set_input(a_)
    pre.estimate_filter()
    feats = [MultiChannelWPEDereverberationFeaturePtr(pre, channel_no=c) for c in range(C)]
    bf = pybeamformer.SubbandGSCLMSBeamformer(feats, min_frames=5)
    bf.calc_beamformer_weights(FS, d)
    sfb = OverSampledDFTSynthesisBankPtr(PyVectorComplexFeatureStreamPtr(bf), prototype=gg, M=M, m=4, r=1, delay_compensation_type=2)
    y = np.concatenate([np.array(b) for b in sfb])
    X = np.stack([restate.analysis(x[c], h, M, 4, 1) for c in range(C)], axis=1)
    Xw, _, _ = restate.wpe(X, samplerate=FS, **wpe)
    Yo, _, _ = restate.gsc_lms(Xw, FS, d, min_frames=5)
    assert rel_l2(y, restate.synthesis(Yo, gg, M, 4, 1)) < 1e-4
    # the same chain into the C++ SubbandGSC (configs[4]: "SubbandGSC + WPE dereverberation chain"), static weights
    gsc = SubbandGSCPtr(fftlen=M, half_band_shift=False)
    for f_ in feats:
        gsc.set_channel(f_)
    gsc.calc_gsc_weights(FS, d)
    sfb = OverSampledDFTSynthesisBankPtr(gsc, prototype=gg, M=M, m=4, r=1, delay_compensation_type=2)
    y = np.concatenate([np.array(b) for b in sfb])
    wq = restate.calc_mainlobe(M, C, FS, d)
    Yg = restate.subband_gsc(Xw, wq, np.zeros_like(wq))
    assert rel_l2(y, restate.synthesis(Yg, gg, M, 4, 1)) < 1e-4


@pytest.mark.gpu
def test_batch_front_end(protos):
    """btk20.batch.BatchBeamformer (the many-utterances-per-submission form bench.py measures) configured like the reference's JSON
    files: gscrls, bmvdr from VAD labels, delay-and-sum behind WPE, each utterance of the batch against the reference's goldens."""
    from oracle import restate
    from distant_speech_recognition_b200.btk20.batch import BatchBeamformer
    from test_oracle import WPE_8
    h, gg = protos[512]; M = 512
    # --- gscrls (confs/gscrls.json defaults, min_frames as in the golden)
    g = load_golden("gscrls_c8_m512"); x = g["x"]
    bb = BatchBeamformer(8, h, gg, M=M, beamformer={"type": "gscrls", "min_frames": int(g["min_frames"])}, max_utterances=2, max_samples=x.shape[1])
    y, Y, st = bb.process(np.stack([x, x]), np.stack([g["delays"], g["delays"]]))
    for u in range(2):
        assert rel_l2(Y[u], g["Y"]) < 1e-4 and rel_l2(y[u][: len(g["time"])], g["time"]) < 1e-4 and st[u][2] == int(g["n_updates"])
    # --- blind MVDR from VAD labels (confs/bmvdr_vad.json)
    g = load_golden("bmvdr_vad_c8_m512"); x = g["x"]
    bb = BatchBeamformer(8, h, gg, M=M, beamformer={"type": "bmvdr", "ref_micx": int(g["ref_micx"]), "offset": float(g["offset"]), "gamma": float(g["gamma"])},
                         max_utterances=2, max_samples=x.shape[1])
    y, Y, st = bb.process(np.stack([x, x]), vad_labels=np.stack([g["labels"], g["labels"]]))
    for u in range(2):
        assert rel_l2(Y[u], g["Y"]) < 1e-4 and rel_l2(y[u], g["time"]) < 1e-4
    # --- delay-and-sum behind the multi-channel WPE (confs/wpe.json keys)
    g = load_golden("wpe_c8_m512"); x = g["x"]
    _, d, _, _ = __import__("distant_speech_recognition_b200.synthetic", fromlist=["x"]).make_utterance(7, 8, 16)
    bb = BatchBeamformer(8, h, gg, M=M, beamformer={"type": "delay_and_sum"}, wpe=dict(WPE_8), max_utterances=2, max_samples=x.shape[1])
    y, Y, st = bb.process(np.stack([x, x]), np.stack([d, d]))
    Xa = np.concatenate([g["Xa"], np.conj(g["Xa"][:, :, 1:256][:, :, ::-1])], axis=2)       # full-band dereverberated snapshots of the reference
    Yo = restate.subband_ds(Xa, restate.calc_mainlobe(M, 8, FS, d))
    for u in range(2):
        assert rel_l2(Y[u], Yo[:, :257]) < 1e-4


@pytest.mark.gpu
def test_generic_python_stream_into_synthesis_and_analysis_iteration(protos):
    """A pure-Python spatial filter between the banks (the reference's PyFeatureStream use): analysis frames are pulled
    one by one in Python, modified, and fed to the synthesis bank through PyVectorComplexFeatureStreamPtr."""
    from oracle import restate
    h, gg = protos[256]; M, D = 256, 128
    rng = np.random.default_rng(0)
    x = (1000 * rng.standard_normal(5000)).astype(np.float32)
    sf = SampleFeaturePtr(block_len=D, shift_len=D, pad_zeros=True); sf.setSamples(x.astype(np.float64), FS)
    afb = OverSampledDFTAnalysisBankPtr(sf, prototype=h, M=M, m=4, r=1, delay_compensation_type=2)
    X = np.array([np.array(v) for v in afb])
    Xo = restate.analysis(x, h, M, 4, 1)
    assert X.shape == Xo.shape and rel_l2(X, Xo) < 1e-5

    class Halve:
        def __init__(self, src): self.src = src
        def __iter__(self):
            for v in self.src: yield 0.5 * np.array(v)
        def size(self): return self.src.size()
        def reset(self): self.src.reset()
    sfb = OverSampledDFTSynthesisBankPtr(PyVectorComplexFeatureStreamPtr(Halve(afb)), prototype=gg, M=M, m=4, r=1, delay_compensation_type=2)
    y = np.concatenate([np.array(b) for b in sfb])
    yo = restate.synthesis(0.5 * Xo, gg, M, 4, 1)
    assert y.shape == yo.shape and rel_l2(y, yo) < 1e-5


# --------------------------------------------------------------------------- chunked realisation of the stream graph (round 2)
@pytest.mark.gpu
@pytest.mark.parametrize("chunk_blocks", [1, 7, 64])
def test_frontend_flow_chunked_equals_whole_utterance(protos, chunk_blocks):
    """set_chunk_blocks(n): the graph behind next() is realised n blocks at a time through btkb_stream_submit instead of on the whole
    utterance — GSC-NLMS -> synthesis and GSC + Zelinski -> synthesis give the same samples, bit for bit, for every chunk size."""
    g = load_golden("gsclms_c8_m512"); h, gg = protos[512]; M, D = 512, 256

    def lms(chunk):
        afbs = _afbs(g["x"], h, M, D)
        bf = pybeamformer.SubbandGSCLMSBeamformer(afbs, min_frames=int(g["min_frames"]))
        bf.beamformer().set_chunk_blocks(chunk)
        bf.calc_beamformer_weights(FS, g["delays"])
        sfb = OverSampledDFTSynthesisBankPtr(PyVectorComplexFeatureStreamPtr(bf), prototype=gg, M=M, m=4, r=1, delay_compensation_type=2)
        y = np.concatenate([np.array(b) for b in sfb])
        return y, bf.total_updates(), np.array(bf.active_weights())

    y0, n0, w0 = lms(0)
    y1, n1, w1 = lms(chunk_blocks)
    assert rel_l2(y0, g["time"]) < 1e-4
    assert y1.shape == y0.shape and np.array_equal(y0, y1) and n0 == n1 and np.array_equal(w0, w1)

    gz = load_golden("gsc_zelinski_c8_m512")

    def zel(chunk):
        afbs = _afbs(gz["x"], h, M, D)
        bf = pybeamformer.SubbandGSCBeamformer(afbs, Nc=1)
        bf.beamformer().set_chunk_blocks(chunk)
        bf._waH[:257] = gz["wa"]
        bf.calc_beamformer_weights(FS, gz["delays"])
        pf = ZelinskiPostFilterPtr(PyVectorComplexFeatureStreamPtr(bf), M, 0.7, 2)
        pf.set_beamformer(bf.beamformer())
        sfb = OverSampledDFTSynthesisBankPtr(pf, prototype=gg, M=M, m=4, r=1, delay_compensation_type=2)
        y = np.concatenate([np.array(b) for b in sfb])
        pf.reset()
        Y = np.array([np.array(v) for v in pf])
        return y, Y

    ya, Ya = zel(0)
    yb, Yb = zel(chunk_blocks)
    assert rel_l2(ya, gz["time"]) < 1e-4
    assert np.array_equal(ya, yb) and np.array_equal(Ya, Yb)


@pytest.mark.gpu
def test_frontend_flow_look_direction_change_inside_the_frame_loop(protos):
    """unit_test/test_online_beamforming.py:205-225: when the conf lists a second target position the script calls
    beamformer.calc_beamformer_weights() inside `for frame_no, buf in enumerate(sfb)`.  With a chunked realisation the new weights
    take effect with the next chunk (chunk = 1 block: with the next frame, like the reference) and the NLMS state is kept.  Checked
    against the fp64 restatement with carried state; the whole-utterance realisation cannot express this (it would restart)."""
    from distant_speech_recognition_b200 import synthetic
    from oracle import restate
    M, D, C, K = 256, 128, 4, 129
    h, gg = protos[M]
    x, d = synthetic.make_batch(1, C, 60 * D + 31, first=950)
    d1, d2 = d[0], 0.3 * d[0]
    F = 37                                              # the frame after which the look direction changes
    afbs = _afbs(x[0], h, M, D)
    bf = pybeamformer.SubbandGSCLMSBeamformer(afbs, min_frames=9)
    bf.beamformer().set_chunk_blocks(1)
    bf.calc_beamformer_weights(FS, d1)
    Y = []
    for frame_no, v in enumerate(bf):
        Y.append(np.array(v))
        if frame_no == F - 1:
            bf.calc_beamformer_weights(FS, d2)
    Y = np.array(Y)
    Xo = np.stack([restate.analysis(x[0, c], h, M, 4, 1) for c in range(C)], axis=1)
    st = {}
    Yo1, _, _ = restate.gsc_lms(Xo[:F], FS, d1, state=st, min_frames=9)
    Yo2, _, _ = restate.gsc_lms(Xo[F:], FS, d2, state=st, min_frames=9)
    assert Y.shape[0] == Xo.shape[0]
    assert rel_l2(Y[:F, :K], Yo1[:, :K]) < 1e-4 and rel_l2(Y[F:, :K], Yo2[:, :K]) < 1e-4


@pytest.mark.gpu
def test_frontend_flow_cpp_subband_gsc_rls(protos):
    """The reference's C++ class SubbandGSCRLSPtr (beamformer.i:289-339): calc_gsc_weights + init_precision_matrix, iterate; against the
    compiled reference's output for the class defaults and a quadratic constraint; next() before init_precision_matrix raises."""
    from distant_speech_recognition_b200.btk20.beamformer import SubbandGSCRLSPtr
    g = load_golden("gscrls_cpp_c4_m256"); h, gg = protos[256]; M, D = 256, 128
    for i, (ctor, init, qc) in enumerate(((dict(myu=0.9, sigma2=0.01), 0.01, None), (dict(myu=0.97, sigma2=0.0), 1e6, (0.5, 2)))):
        afbs = _afbs(g["x"], h, M, D)
        bf = SubbandGSCRLSPtr(fftlen=M, half_band_shift=False, **ctor)
        for a in afbs:
            bf.set_channel(a)
        bf.calc_gsc_weights(FS, g["delays"])
        if i == 0:
            with pytest.raises(Exception, match="precision matrix"):
                bf.next()
        bf.init_precision_matrix(init)
        if qc:
            bf.set_quadratic_constraint(*qc)
        Y = np.array([np.array(v) for v in bf])
        assert Y.shape == (g["Y%d" % i].shape[0], M) and rel_l2(Y[:, :129], g["Y%d" % i]) < 1e-4


@pytest.mark.gpu
def test_frontend_flow_wpe_filters_of_one_utterance_applied_to_another(protos):
    """MultiChannelWPEDereverberationFeature streams feeding a beamformer must apply the filters of the EARLIER
    estimate_filter() call (dereverberation.cc:441-497, 713-728), not re-estimate them on the audio being processed: estimate on
    utterance A, re-read the sources with utterance B, run analysis -> WPE(apply) -> D&S; against the restatement's
    wpe_estimate(A) + wpe_apply(B)."""
    from oracle import restate
    from distant_speech_recognition_b200 import synthetic
    from distant_speech_recognition_b200.btk20.dereverberation import MultiChannelWPEDereverberationPtr, MultiChannelWPEDereverberationFeaturePtr
    M, D, C, n = 256, 128, 3, 6000
    h, gg = protos[M]
    xa, d, _, _ = synthetic.make_utterance(71, C, n)
    xb = synthetic.make_utterance(72, C, n)[0]
    wpe = dict(lower_num=1, upper_num=5, iterations_num=2, load_db=-25.0, band_width=0.0, diagonal_bias=1e-4)
    sfs, afbs = [], []
    for c in range(C):
        sf = SampleFeaturePtr(block_len=D, shift_len=D, pad_zeros=True); sf.setSamples(xa[c].astype(np.float64), FS)
        sfs.append(sf); afbs.append(OverSampledDFTAnalysisBankPtr(sf, prototype=h, M=M, m=4, r=1, delay_compensation_type=2))
    pre = MultiChannelWPEDereverberationPtr(subbands_num=M, channels_num=C, samplerate=FS, **wpe)
    for a_ in afbs:
        pre.set_input(a_)
    pre.estimate_filter()
    for c in range(C):
        sfs[c].setSamples(xb[c].astype(np.float64), FS)          # new audio behind the same graph
    feats = [MultiChannelWPEDereverberationFeaturePtr(pre, channel_no=c) for c in range(C)]
    ds = SubbandDSPtr(fftlen=M, half_band_shift=False)
    for f_ in feats:
        ds.set_channel(f_)
    ds.calc_array_manifold_vectors(FS, d)
    Y = np.array([np.array(v) for v in ds])
    XA = np.stack([restate.analysis(xa[c], h, M, 4, 1) for c in range(C)], axis=1)
    XB = np.stack([restate.analysis(xb[c], h, M, 4, 1) for c in range(C)], axis=1)
    G = restate.wpe_estimate(XA, wpe["lower_num"], wpe["upper_num"], wpe["iterations_num"], wpe["load_db"], wpe["band_width"], wpe["diagonal_bias"], FS)
    XBd = restate.wpe_apply(XB, G, wpe["lower_num"], wpe["upper_num"], wpe["band_width"], FS)
    Yo = restate.subband_ds(XBd, restate.calc_mainlobe(M, C, FS, d))
    assert rel_l2(Y[:, :129], Yo[:, :129]) < 1e-4
    Yre = restate.subband_ds(restate.wpe(XB, samplerate=FS, **wpe)[0], restate.calc_mainlobe(M, C, FS, d))
    assert rel_l2(Yre[:, :129], Yo[:, :129]) > 1e-3               # re-estimating on B would be visibly different
