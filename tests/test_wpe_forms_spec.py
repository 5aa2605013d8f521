"""Specification check for the two forms of the WPE normal equations (product code: btkb_wpe.cu k_wpe_gram_dual,
k_wpe_chol<DUAL>; GPU parity: tests/test_parity_gpu_r2.py::test_wpe_frame_domain_form_* / test_wpe_picks_the_smaller_system_per_batch).

estimate_Gn_ (dereverberation/dereverberation.cc:553-690) solves, per bin and output channel c,

    (A Th_c^-1 A^H + delta_c I) g_c = A Th_c^-1 ybar_c,   delta_c = bias + load (max_i (A Th_c^-1 A^H)_ii + bias)

with A = [lags(s)] (L x S), Th_c = diag(theta_c), ybar_c(s) = conj(x_c(s)): load_R_ (:648-663) adds the SAME amount to every diagonal
entry, which is what makes the push-through identity applicable:

    g_c = A (A^H A + delta_c Th_c)^-1 ybar_c            (an S x S system; K = A^H A does not depend on c or on the iteration)

and K is a sum over channels of sliding-window sums of ONE sequence per diagonal,  K[s'+d][s'] = sum_{l<P} p_d(s'-l),
p_d(t) = sum_c conj(x_c(t+d)) x_c(t).  These tests pin both statements down on the oracle's own restatement (oracle/restate.py
wpe_estimate), in fp64, including lower_num > 0 and an utterance shorter than the filter (S < L, the case the product serves in this
form).  CPU only; test infrastructure (imports oracle/)."""
import numpy as np
import pytest

from oracle import restate


def _lag_matrix(F, k, lower, P):
    """A[i][s'] = lags_i(s' + lower) for the estimation frames s = lower .. T-1 (get_lags_(subbandX, sampleX - lowerN_), :553-617)."""
    T, C, _ = F.shape
    S = max(T - lower, 0)
    return np.stack([restate.wpe_lags(F, k, s, P) for s in range(S)], axis=1) if S else np.zeros((C * P, 0), complex)


@pytest.mark.parametrize("T,C,lower,upper", [(9, 3, 0, 4), (30, 2, 1, 5), (12, 4, 2, 6)])
def test_frame_domain_form_gives_the_filters_of_estimate_Gn(T, C, lower, upper):
    rng = np.random.default_rng(T)
    M, k = 8, 3
    F = 50.0 * (rng.standard_normal((T, C, M)) + 1j * rng.standard_normal((T, C, M)))
    F[2] *= 1e-6                                    # a near-silent frame: theta hits its floor there
    P = upper - lower + 1
    L = C * P
    kw = dict(lower_num=lower, upper_num=upper, load_db=-18.0, band_width=0.0, diagonal_bias=1e-4, samplerate=16000.0)
    A = _lag_matrix(F, k, lower, P)                 # [L][S]
    S = A.shape[1]
    assert (S < L) == (T in (9, 12))                # two of the three cases are "utterance shorter than the filter"
    load = 10.0 ** (kw["load_db"] / 10.0)
    Kmat = A.conj().T @ A
    G_prev = np.zeros((C, M // 2 + 1, L), complex)
    for it in (1, 2):
        G_ref = restate.wpe_estimate(F, iterations_num=it, **kw)          # lag-domain, as the reference builds it
        for c in range(C):
            x = F[lower:, c, k]
            pred = np.conj(G_prev[c, k]) @ A                                # zdotc(G, lags), iteration it - 1's filters
            theta = np.maximum(np.abs(x - pred), 1e-3) ** 2
            Rdiag = (np.abs(A) ** 2 / theta).sum(axis=1)
            delta = kw["diagonal_bias"] + load * (Rdiag.max() + kw["diagonal_bias"])
            z = np.linalg.solve(Kmat + delta * np.diag(theta), np.conj(x))
            g = A @ z
            assert np.linalg.norm(g - G_ref[c, k]) <= 1e-9 * max(np.linalg.norm(G_ref[c, k]), 1e-30), (it, c)
        G_prev = G_ref


def test_frame_domain_gram_is_a_window_sum_per_diagonal():
    rng = np.random.default_rng(3)
    T, C, M, k, lower, P = 20, 3, 8, 2, 1, 6
    F = rng.standard_normal((T, C, M)) + 1j * rng.standard_normal((T, C, M))
    A = _lag_matrix(F, k, lower, P)
    S = A.shape[1]
    Kmat = A.conj().T @ A
    x = F[:, :, k]                                  # x_c(t), t = absolute frame; lags_(c,l)(s') = x_c(s' - l)
    for d in range(S):
        p = np.array([np.sum(np.conj(x[t + d]) * x[t]) for t in range(S - d)])
        for sp in range(S - d):
            w = sum(p[sp - l] for l in range(P) if sp - l >= 0)
            assert abs(w - Kmat[sp + d, sp]) <= 1e-12 * max(abs(Kmat[sp + d, sp]), 1.0)
