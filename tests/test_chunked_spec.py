"""Specification check for chunked submission (DESIGN.md §4 "Streamed chunks with carried state"; built in round 2 as btkb_stream_begin /
btkb_stream_submit, which number frames and blocks absolutely instead of dropping the re-computed ones; GPU identity tests:
tests/test_parity_gpu_r2.py::test_streamed_*).

A live front end hands the pipe an utterance piece by piece.  With the oracle's restatement of the two filter banks
(oracle/restate.py: OverSampledDFTAnalysisBank::next, modulated/modulated.cc:363-469; OverSampledDFTSynthesisBank::next,
modulated.cc:551-612) these tests pin down — bit for bit, in the reference's own fp64 arithmetic — how much history a chunk must be
prefixed with and how many leading frames of the chunk's result are to be dropped, so that K1 (analysis) and K5 (synthesis) can serve
chunks UNCHANGED (a pointer offset on their outputs), and only the per-bin kernel has to load / store its recurrence state:

  analysis:  prefix  H_a = m M - D samples,        drop  m R - 1 - laN  frames,  non-final chunks emit  n/D - laN  frames (no flush frames)
  synthesis: prefix  H_s = max(m R - 1, pd_S + R - 1) subband frames, drop  H_s - pd_S blocks,  a chunk ending at frame F emits blocks up to F - pd_S
             (until the stream holds H_s frames the chunk is served from frame 0: the reference's priming is not a zero history)

The third test names the state the per-bin NLMS kernel (K4) has to load and store at a chunk boundary: u (C complex) and the sub-band
energy per (utterance, bin) chain; E_avg, the step size gamma, the frame counter and the update count per utterance.

CPU only; test infrastructure (imports oracle/)."""
import numpy as np
import pytest

from oracle import restate


def _proto(M, m, seed):
    rng = np.random.default_rng(seed)
    return rng.standard_normal(M * m) / M, rng.standard_normal(M * m) / M


@pytest.mark.parametrize("M,m,r,dct", [(64, 4, 1, 2), (64, 2, 2, 2), (32, 3, 0, 2), (64, 4, 1, 1)])
def test_analysis_chunks_with_sample_history_reproduce_the_whole_run(M, m, r, dct):
    D, R = M >> r, 1 << r
    pd, la = restate.fb_delays(m, r, dct, False)
    h, _ = _proto(M, m, 1)
    rng = np.random.default_rng(2)
    n = 37 * D + 11                                   # ragged end: the last block is zero-padded (feature.cc:605-649)
    x = (3000.0 * rng.standard_normal(n)).astype(np.float32)
    whole = restate.analysis(x, h, M, m, r, dct)
    H, skip = m * M - D, m * R - 1 - la
    cuts = [0, 9 * D, 10 * D, 24 * D, n]              # chunk boundaries are multiples of D except the end of the utterance
    got, emitted = [], 0
    for j in range(len(cuts) - 1):
        a, b, final = cuts[j], cuts[j + 1], j == len(cuts) - 2
        first = a == 0
        xin = x[:b] if first else x[a - min(H, a):b]
        if not first and a < H:                        # fewer than H true samples exist yet: the missing history is the zero history of a fresh bank
            xin = np.concatenate([np.zeros(H - a, np.float32), xin])
        F = restate.analysis(xin, h, M, m, r, dct)
        lo = 0 if first else skip
        hi = F.shape[0] if final else (len(xin) // D - la)   # a non-final chunk must not emit the pd_A flush frames of the end of the stream
        got.append(F[lo:hi]); emitted += hi - lo
        if not final:
            assert emitted == b // D - la
    got = np.concatenate(got)
    assert got.shape == whole.shape and np.array_equal(got.view(np.float64), whole.view(np.float64))


@pytest.mark.parametrize("M,m,r,dct", [(64, 4, 1, 2), (64, 2, 2, 2), (32, 3, 0, 2), (64, 4, 1, 1)])
def test_synthesis_chunks_with_subband_history_reproduce_the_whole_run(M, m, r, dct):
    D, R = M >> r, 1 << r
    pd, _ = restate.fb_delays(m, r, dct, True)
    _, g = _proto(M, m, 3)
    rng = np.random.default_rng(4)
    T = 41
    Y = rng.standard_normal((T, M)) + 1j * rng.standard_normal((T, M))
    whole = restate.synthesis(Y, g, M, m, r, dct).reshape(-1, D)
    Hs = max(m * R - 1, pd + R - 1)                   # frames a block depends on; and at least R - 1 dropped blocks, because a fresh bank takes w_t = 0 for t < 0
    skip = Hs - pd
    cuts = [0, 5, 6, 23, T]
    got = []
    for j in range(len(cuts) - 1):
        a, b = cuts[j], cuts[j + 1]
        if a < Hs:
            # Start of the stream.  The reference primes its buffer with the first pd_S frames without forming their polyphase sums
            # (modulated.cc:580-590), so for R > 1 the first R - 1 blocks are NOT what zero history would give (w_t, t < 0, is taken as
            # zero although the frames it would use exist): until the stream holds H_s frames a chunk is served from frame 0.
            out = restate.synthesis(Y[:b], g, M, m, r, dct).reshape(-1, D)[max(a - pd, 0):]
        else:
            out = restate.synthesis(Y[a - Hs:b], g, M, m, r, dct).reshape(-1, D)[skip:]
        got.append(out)
        assert sum(len(o) for o in got) == max(b - pd, 0)      # blocks emitted so far
    got = np.concatenate(got)
    assert got.shape == whole.shape and np.array_equal(got.view(np.uint32), whole.view(np.uint32))


def test_nlms_chunks_with_carried_state_reproduce_the_whole_run():
    """SubbandGSCLMSBeamformer.__iter__ (lib/pybeamformer.py:659-734) over a stream cut into chunks: with the state of
    restate.gsc_lms carried, subband output, final active weights and update count equal the whole run bit for bit — including
    across the step-size halving (slowdown_after) and the min_frames warm-up, which depend on the ABSOLUTE frame number."""
    rng = np.random.default_rng(7)
    M, C, T = 32, 4, 60
    X = 3000.0 * (rng.standard_normal((T, C, M)) + 1j * rng.standard_normal((T, C, M)))
    X[20:26] = 0.0                                     # a silent stretch: the energy gate closes (no adaptation)
    delays = np.array([0.0, 1.1e-4, 2.3e-4, 3.2e-4])
    kw = dict(min_frames=9, slowdown_after=16, regularization_param=1.0e-4, max_wa_l2norm=1.0e-3)
    Yw, waw, nw = restate.gsc_lms(X, 16000.0, delays, **kw)
    st, Ys = {}, []
    for a, b in ((0, 7), (7, 8), (8, 33), (33, T)):
        Yc, wac, nc = restate.gsc_lms(X[a:b], 16000.0, delays, state=st, **kw)
        Ys.append(Yc)
    Ys = np.concatenate(Ys)
    assert 0 < nw < T and nc == nw and st["isamp"] == T
    assert np.array_equal(Ys.view(np.float64), Yw.view(np.float64)) and np.array_equal(wac.view(np.float64), waw.view(np.float64))
