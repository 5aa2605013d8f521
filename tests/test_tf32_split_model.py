"""CPU model of the arithmetic k_covariance_tc feeds to the tensor cores (csrc/btkb_cov_tc.cu): TF32 keeps 10 mantissa bits, the kernel
splits every fp32 snapshot into hi (low 13 mantissa bits cleared) and lo = x - hi and accumulates hi hi^T + lo hi^T + hi lo^T.  This
pins the precision claims of DESIGN.md (K2w) without a GPU: a plain TF32 Gram misses the 1e-4 parity budget once the MVDR solve
amplifies it, the three-product split is fp32-class, and the two-product variant T = hi hi^T + 2 lo hi^T recovers the same Gram after
Hermitian symmetrisation (the planned next step)."""
import numpy as np


def tf32(x):
    """What a kind::tf32 operand keeps of an fp32 value (truncation of the low 13 mantissa bits)."""
    return (np.asarray(x, np.float32).view(np.uint32) & np.uint32(0xFFFFE000)).view(np.float32)


def gram(A, B):
    """fp32-accumulated real Gram A B^T of tf32 operands (products of two tf32 numbers are exact in fp32's 24-bit significand)."""
    return (tf32(A).astype(np.float64) @ tf32(B).astype(np.float64).T).astype(np.float32).astype(np.float64)


def complex_cov(S, C):
    """R[c][c'] from the real Gram S of A = [Xr ; Xi] (rows 0..C-1 real parts, C..2C-1 imaginary parts)."""
    return (S[:C, :C] + S[C:, C:]) + 1j * (S[C:, :C] - S[:C, C:])


def test_tf32_split_precision_and_two_product_variant():
    rng = np.random.default_rng(0)
    C, T = 64, 317
    # int16-scale snapshots with a dominant source 20 dB above the sensor noise (an ill-conditioned covariance)
    s = rng.standard_normal(T) + 1j * rng.standard_normal(T)
    v = np.exp(1j * rng.uniform(0, 2 * np.pi, C))
    X = (3e5 * np.outer(v, s) + 3e4 * (rng.standard_normal((C, T)) + 1j * rng.standard_normal((C, T)))).astype(np.complex64)   # 20 dB
    A = np.concatenate([X.real, X.imag]).astype(np.float32)                      # [2C][T]
    R64 = X.astype(np.complex128) @ np.conj(X.astype(np.complex128)).T
    hi = tf32(A); lo = (A - hi).astype(np.float32)
    assert np.array_equal(hi.astype(np.float64) + lo.astype(np.float64), A.astype(np.float64))   # the split is exact

    def err(R):
        return np.linalg.norm(R - R64) / np.linalg.norm(R64)

    e_plain = err(complex_cov(gram(A, A), C))
    e_split = err(complex_cov(gram(hi, hi) + gram(lo, hi) + gram(hi, lo), C))
    Tm = gram(hi, hi) + gram(2.0 * lo, hi)                                        # two products; not symmetric
    Rp = complex_cov(Tm, C)
    e_two = err(0.5 * (Rp + np.conj(Rp).T))                                       # Hermitian symmetrisation of R' 
    assert 5e-5 < e_plain < 2e-3                                                  # ~2^-11 per operand
    assert e_split < 1e-6 and e_two < 1e-6                                        # fp32 class
    # what the MVDR solve makes of it: w = R^-1 d / (d^H R^-1 d) with light loading
    d = v / C
    def mvdr(R):
        Rl = R + 1e-4 * np.trace(R).real / C * np.eye(C)
        t = np.linalg.solve(Rl, d)
        return t / np.vdot(d, t)
    w64 = mvdr(R64)
    ew_plain = np.linalg.norm(mvdr(complex_cov(gram(A, A), C)) - w64) / np.linalg.norm(w64)
    ew_split = np.linalg.norm(mvdr(complex_cov(gram(hi, hi) + gram(lo, hi) + gram(hi, lo), C)) - w64) / np.linalg.norm(w64)
    assert ew_split < 1e-4 < ew_plain
