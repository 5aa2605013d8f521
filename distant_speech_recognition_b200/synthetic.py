"""Deterministic synthetic multichannel utterances (SURVEY.md §8d / BASELINE.md §2).

fs = 16 kHz, float32 samples at int16 scale (the reference reads wavs un-normalised, feature/feature.cc:265-270).
Utterance u (seed 20260925 + u): target = Gaussian noise low-passed to 4 kHz, sigma 3000, active from t = 1.0 s;
one interferer, sigma 1500, from a second direction; spatially white sensor noise, sigma 100.  Per-channel signals
are built in float64 with exact fractional delays (FFT phase ramp) from the far-field delay conventions of
lib/pybeamformer.py:41-94 (c = 343 740 mm/s) and cast to float32.  NumPy only — this is input data, not the hot path.
"""
import numpy as np

SSPEED = 343740.0
BASE_SEED = 20260925


def linear_array(C, pitch_mm=40.0):
    """C-mic linear array centred on the origin: positions [C][3] in mm."""
    x = (np.arange(C) - (C - 1) / 2.0) * pitch_mm
    return np.stack([x, np.zeros(C), np.zeros(C)], axis=1)


def planar_array(nx, ny, pitch_mm=40.0):
    xs = (np.arange(nx) - (nx - 1) / 2.0) * pitch_mm
    ys = (np.arange(ny) - (ny - 1) / 2.0) * pitch_mm
    return np.array([[x, y, 0.0] for y in ys for x in xs])


def array_for_channels(C):
    """cfg1: 2 mics +-50 mm; cfg2/3/5: 8-mic linear 40 mm; cfg4: 8x8 planar 40 mm."""
    if C == 2:
        return "linear", linear_array(2, 100.0)
    if C == 64:
        return "planar", planar_array(8, 8, 40.0)
    return "linear", linear_array(C, 40.0)


def far_field_delays(array_type, mpos, azimuth, polar=np.pi / 4, ref_micx=None):
    """calc_la_delays / calc_pa_delays (lib/pybeamformer.py:41-94)."""
    mpos = np.asarray(mpos, np.float64)
    C = len(mpos)
    if ref_micx is None:
        ref_micx = C // 2
    if array_type == "linear":
        d = -mpos[:, 0] * np.cos(azimuth) / SSPEED
        return d - d[ref_micx]
    dx = mpos[:, 0] - mpos[ref_micx, 0]
    dy = mpos[:, 1] - mpos[ref_micx, 1]
    return -(dx * np.cos(azimuth) * np.sin(polar) + dy * np.sin(azimuth) * np.sin(polar)) / SSPEED


def _delayed_copies(s, delays, fs):
    """x_c(t) = s(t - tau_c) by an exact phase ramp on a zero-padded FFT."""
    n = len(s)
    nfft = 1 << int(np.ceil(np.log2(n + 256)))
    S = np.fft.rfft(s, nfft)
    f = np.fft.rfftfreq(nfft, 1.0 / fs)
    out = np.fft.irfft(S[None, :] * np.exp(-2j * np.pi * f[None, :] * np.asarray(delays)[:, None]), nfft, axis=1)
    return out[:, :n]


def _lowpass_noise(rng, n, fs, cutoff, sigma):
    w = rng.standard_normal(n)
    W = np.fft.rfft(w)
    f = np.fft.rfftfreq(n, 1.0 / fs)
    W[f > cutoff] = 0.0
    s = np.fft.irfft(W, n)
    return s * (sigma / max(np.std(s), 1e-12))


def make_utterance(u, C, n, fs=16000.0, target_az=np.pi / 3, interferer_az=2 * np.pi / 3, target_start_s=1.0, pcm16=False):
    """Returns (samples float32 [C][n], delays_target float64 [C], mpos [C][3], array_type).
    pcm16=True rounds and clips to the int16 grid (what a 16-bit wav holds): the float32 values are then exactly
    representable as int16, so the float and the 16-bit PCM entry points of the C-ABI see identical inputs."""
    rng = np.random.default_rng(BASE_SEED + int(u))
    array_type, mpos = array_for_channels(C)
    d_t = far_field_delays(array_type, mpos, target_az)
    d_j = far_field_delays(array_type, mpos, interferer_az)
    s = _lowpass_noise(rng, n, fs, 4000.0, 3000.0)
    s[: min(n, int(target_start_s * fs))] = 0.0
    j = _lowpass_noise(rng, n, fs, 4000.0, 1500.0)
    x = _delayed_copies(s, d_t, fs) + _delayed_copies(j, d_j, fs) + 100.0 * rng.standard_normal((C, n))
    if pcm16:
        x = np.clip(np.rint(x), -32768, 32767)
    return x.astype(np.float32), d_t, mpos, array_type


def make_batch(U, C, n, fs=16000.0, first=0, pcm16=False):
    """samples float32 [U][C][n], delays float64 [U][C]."""
    X = np.empty((U, C, n), np.float32)
    dl = np.empty((U, C), np.float64)
    for u in range(U):
        X[u], dl[u], _, _ = make_utterance(first + u, C, n, fs, pcm16=pcm16)
    return X, dl
