"""Specification check of the factorisation scheme shared by the two fp64 Cholesky kernels (btkb_wpe.cu k_wpe_chol, btkb_wide.cu
k_mvdr_solve_wide_blk; DESIGN.md §4 K7 and "wide arrays"), restated in NumPy step by step as the kernels do it:

  * right-looking panel Cholesky with 16-column panels: diagonal block, L21 = A21 L11^-H (right-looking inside the row), trailing update
    A22 -= L21 L21^H — with the trailing update split into the four REAL products the tensor-core path issues
    (Sr = Ar Br^T + Ai Bi^T, Si = Ai Br^T - Ar Bi^T);
  * the right-hand side riding along as an extra row (k_wpe_chol: conj(r) as row L of the augmented matrix; the wide solve: the row vector
    d^H), so that the forward substitution is part of the factorisation: after the last panel the row holds (L^-1 rhs)^H;
  * look-ahead: the next diagonal block is taken out of the trailing update first and factored before the rest of the update is applied
    (the order of these two does not change any value — they touch disjoint entries — which is what the test asserts);
  * panel-wise backward substitution L^H g = y.

CPU only; pure NumPy."""
import numpy as np
import pytest

NB = 16


def _factor_diag(D):
    """Column sweep on an nb x nb Hermitian block (lower triangle used): scale by 1/sqrt(pivot), rank-one update of the rest."""
    D = D.copy(); nb = D.shape[0]
    for jj in range(nb):
        inv = 1.0 / np.sqrt(D[jj, jj].real)
        D[jj:, jj] *= inv; D[jj, jj] = D[jj, jj].real
        for r in range(jj + 1, nb):
            for k in range(jj + 1, r + 1):
                D[r, k] -= D[r, jj] * np.conj(D[k, jj])
    return np.tril(D)


def _trsm_row(v, L11):
    """x L11^H = v, right-looking: once x_jj is final the later entries take their updates."""
    v = v.copy(); nb = L11.shape[0]
    for jj in range(nb):
        v[jj] /= L11[jj, jj].real
        v[jj + 1:nb] -= v[jj] * np.conj(L11[jj + 1:nb, jj])
    return v


def _real_products(Lb, Lk):
    """S = Lb Lk^H as the four real products of the mma path."""
    Ar, Ai, Br, Bi = Lb.real, Lb.imag, Lk.real, Lk.imag
    return (Ar @ Br.T + Ai @ Bi.T) + 1j * (Ai @ Br.T - Ar @ Bi.T)


def blocked_solve(A, rhs, look_ahead):
    n = A.shape[0]
    M = np.zeros((n + 1, n + 1), complex)
    M[:n, :n] = np.tril(A); M[n, :n] = np.conj(rhs)            # augmented row: conj(rhs)
    ahead = False
    for j0 in range(0, n, NB):
        nb = min(NB, n - j0); j1 = j0 + nb
        if not ahead:
            M[j0:j1, j0:j1] = _factor_diag(M[j0:j1, j0:j1])
        L11 = M[j0:j1, j0:j1]
        for r in range(j1, n + 1):
            M[r, j0:j1] = _trsm_row(M[r, j0:j1], L11)
        L21 = M[j1:, j0:j1]
        ahead = look_ahead and (n - j1 >= NB)
        blocks = [(bi, bk) for bi in range(0, n + 1 - j1, NB) for bk in range(0, bi + 1, NB)]
        if ahead:                                                # block (0, 0) first, factored at once
            M[j1:j1 + NB, j1:j1 + NB] = np.tril(M[j1:j1 + NB, j1:j1 + NB] - _real_products(L21[:NB], L21[:NB]))
            M[j1:j1 + NB, j1:j1 + NB] = _factor_diag(M[j1:j1 + NB, j1:j1 + NB])
            blocks = blocks[1:]
        for bi, bk in blocks:
            S = _real_products(L21[bi:bi + NB], L21[bk:bk + NB])
            blk = M[j1 + bi:j1 + bi + NB, j1 + bk:j1 + bk + NB]
            blk -= S[:blk.shape[0], :blk.shape[1]]
            if bi == bk:
                M[j1 + bi:j1 + bi + NB, j1 + bk:j1 + bk + NB] = np.tril(blk)
    L = M[:n, :n]
    y = np.conj(M[n, :n])                                        # = L^-1 rhs
    g = y.copy()
    for j0 in range(((n - 1) // NB) * NB, -1, -NB):              # L^H g = y, last panel first
        nb = min(NB, n - j0)
        for jj in range(nb - 1, -1, -1):
            g[j0 + jj] /= L[j0 + jj, j0 + jj].real
            g[j0:j0 + jj] -= np.conj(L[j0 + jj, j0:j0 + jj]) * g[j0 + jj]
        g[:j0] -= np.conj(L[j0:j0 + nb, :j0]).T @ g[j0:j0 + nb]
    return L, y, g


@pytest.mark.parametrize("n", [5, 16, 37, 64, 157])
def test_blocked_scheme_solves_the_system_and_look_ahead_changes_nothing(n):
    rng = np.random.default_rng(n)
    B = rng.standard_normal((n, n + 3)) + 1j * rng.standard_normal((n, n + 3))
    A = B @ B.conj().T + 0.5 * np.eye(n)
    rhs = rng.standard_normal(n) + 1j * rng.standard_normal(n)
    L0, y0, g0 = blocked_solve(A, rhs, look_ahead=False)
    L1, y1, g1 = blocked_solve(A, rhs, look_ahead=True)
    assert np.array_equal(L0, L1) and np.array_equal(g0, g1)   # disjoint entries: the order is free
    Lref = np.linalg.cholesky(A)
    assert np.abs(L0 - Lref).max() < 1e-10 * np.abs(Lref).max()
    assert np.abs(y0 - np.linalg.solve(Lref, rhs)).max() < 1e-10 * np.abs(y0).max()
    assert np.abs(g0 - np.linalg.solve(A, rhs)).max() < 1e-9 * np.abs(g0).max()
