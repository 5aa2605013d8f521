"""The per-chain fp64 solve of the SOS beamformers (csrc/btkb_sos_math.cuh: blind-MVDR LU solve, GEV Cholesky reduction +
complex Jacobi) compiled for the CPU with g++ and checked against the oracle's restatement and against the reference's own
output (goldens from lib/pybeamformer.py:1257-1357 through oracle/pyref.py).  This is the same source k_sos_solve compiles for
sm_100a; the harness (tests/host/sos_math_host.cc) is test infrastructure, the product has no CPU path."""
import os
import subprocess
import numpy as np
import pytest

from oracle import restate

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")


@pytest.fixture(scope="module")
def harness(tmp_path_factory):
    exe = str(tmp_path_factory.mktemp("sosmath") / "sos_math_host")
    subprocess.check_call(["g++", "-O2", "-std=c++17", os.path.join(ROOT, "tests", "host", "sos_math_host.cc"), "-o", exe])
    return exe


def solve(exe, kind, Rt, Rn, gamma, ref_micx=0, offset=0.0):
    K, C, _ = Rt.shape
    lines = ["%d %d %d %.17g %d %.17g" % (kind, C, K, gamma, ref_micx, offset)]
    for m in range(K):
        for A in (Rt[m], Rn[m]):
            lines.append(" ".join("%.17g %.17g" % (z.real, z.imag) for z in A.reshape(-1)))
    out = subprocess.run([exe], input="\n".join(lines) + "\n", capture_output=True, text=True, check=True).stdout
    rows = np.array([[float(v) for v in ln.split()] for ln in out.strip().split("\n")])
    ok = rows[:, 0].astype(bool)
    w = rows[:, 1::2] + 1j * rows[:, 2::2]
    return ok, w




def random_stats(K, C, seed):
    rng = np.random.default_rng(seed)
    def herm(n):
        A = rng.standard_normal((K, C, n)) + 1j * rng.standard_normal((K, C, n))
        return A @ np.conj(np.transpose(A, (0, 2, 1)))
    Rt = herm(3) * 50.0          # low-rank-ish target
    Rn = herm(4 * C) + 0.1 * np.eye(C)
    ct = rng.integers(5, 40, K); cn = rng.integers(5, 40, K)
    return Rt, Rn, ct, cn


@pytest.mark.parametrize("C", [2, 4, 8])
def test_bmvdr_solve_matches_restatement(harness, C):
    Rt, Rn, ct, cn = random_stats(33, C, 100 + C)
    w_ref = restate.sos_bmvdr_weights(Rt, Rn, ct, cn, gamma=1e-6, ref_micx=C - 1, offset=0.25)
    ok, w = solve(harness, 0, Rt / ct[:, None, None], Rn / cn[:, None, None], 1e-6, C - 1, 0.25)
    assert ok.all()
    assert np.linalg.norm(w - w_ref) / np.linalg.norm(w_ref) < 1e-10


@pytest.mark.parametrize("C", [2, 4, 8])
def test_gev_eigenvector_matches_restatement_and_scipy(harness, C):
    import scipy.linalg
    Rt, Rn, ct, cn = random_stats(33, C, 200 + C)
    ok, v = solve(harness, 1, Rt, Rn / cn[:, None, None], 1e-6)
    assert ok.all()
    # the harness returns the un-aligned eigenvector per bin (k_sos_align does the bin-to-bin phase): compare per bin up to phase
    for m in range(Rt.shape[0]):
        rn = restate.improve_matrix_condition(Rn[m] / cn[m], 1e-6)
        rn = rn / (np.trace(rn) / C)
        ev, V = scipy.linalg.eigh(Rt[m], rn)
        ref = V[:, -1]
        ph = np.vdot(ref, v[m])
        assert abs(abs(ph) - np.vdot(ref, ref).real) < 1e-8 * np.vdot(ref, ref).real      # same direction and the same B-norm
        assert np.linalg.norm(v[m] * np.conj(ph) / abs(ph) - ref) / np.linalg.norm(ref) < 1e-8
        L = np.linalg.cholesky(rn)
        y0 = np.vdot(L[:, 0], v[m])
        assert abs(y0.imag) < 1e-10 * abs(y0) and y0.real > 0                               # the documented phase convention
        assert abs(np.vdot(v[m], rn @ v[m]).real - 1.0) < 1e-10                              # scipy's normalisation v^H B v = 1


def test_singular_noise_covariance_is_reported(harness):
    C, K = 4, 3
    Rt = np.tile(np.eye(C, dtype=complex), (K, 1, 1))
    Rn = np.zeros((K, C, C), complex)
    ok, _ = solve(harness, 0, Rt, Rn, 0.0)
    assert not ok.any()
    ok, _ = solve(harness, 1, Rt, Rn + np.diag([1.0, 1.0, 1.0, -1.0]), 0.0)
    assert not ok.any()
