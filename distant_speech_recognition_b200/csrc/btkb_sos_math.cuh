// btkb_sos_math.cuh — per-chain fp64 linear algebra of the SOS batch beamformers (blind MVDR solve, GEV eigenvector),
// written __host__ __device__ so that tests/ can compile the very same code for the CPU and check it against the oracle
// without a GPU (tests/test_sos_math_host.py).  The product only ever calls it from k_sos_solve (btkb_sos.cu).
#pragma once
#include <math.h>
#ifndef BTKB_SOS_BMVDR
#define BTKB_SOS_BMVDR 0
#define BTKB_SOS_GEV 1
#endif
#ifdef __CUDACC__
#define BTKB_HD __host__ __device__ __forceinline__
#else
#define BTKB_HD inline
#endif

namespace btkb {

struct zd { double x, y; };
BTKB_HD zd zmk(double x, double y) { zd r; r.x = x; r.y = y; return r; }
BTKB_HD zd zadd(zd a, zd b) { return zmk(a.x + b.x, a.y + b.y); }
BTKB_HD zd zsub(zd a, zd b) { return zmk(a.x - b.x, a.y - b.y); }
BTKB_HD zd zmul(zd a, zd b) { return zmk(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
BTKB_HD zd zmulc(zd a, zd b) { return zmk(a.x * b.x + a.y * b.y, a.y * b.x - a.x * b.y); }  // a conj(b)
BTKB_HD zd zconj(zd a) { return zmk(a.x, -a.y); }
BTKB_HD zd zscale(zd a, double s) { return zmk(a.x * s, a.y * s); }
BTKB_HD double zabs2(zd a) { return a.x * a.x + a.y * a.y; }
BTKB_HD zd zdiv(zd a, zd b) { const double d = 1.0 / zabs2(b); return zmk((a.x * b.x + a.y * b.y) * d, (a.y * b.x - a.x * b.y) * d); }

// improve_matrix_condition (pybeamformer.py:1231-1240): (x + gamma tr(x)/C I) / (1 + gamma); tr is complex in the reference
template <int C>
BTKB_HD void improve_condition(zd (*A)[C], double gamma) {
  zd tr = zmk(0, 0);
  for (int i = 0; i < C; i++) tr = zadd(tr, A[i][i]);
  const zd sc = zscale(tr, gamma / (double)C);
  const double inv = 1.0 / (1.0 + gamma);
  for (int i = 0; i < C; i++) {
    A[i][i] = zadd(A[i][i], sc);
    for (int j = 0; j < C; j++) A[i][j] = zscale(A[i][j], inv);
  }
}


// GEV stage 1: Rn /= tr(Rn)/C (pybeamformer.py:1326-1328), Cholesky Rn = L L^H (lower triangle of Rn, in place),
// Rt <- L^-1 Rt L^-H (Hermitian).  Returns false when Rn is not positive definite.
template <int C>
BTKB_HD bool gev_reduce(zd (*Rt)[C], zd (*Rn)[C]) {
  zd tr = zmk(0, 0);
  for (int i = 0; i < C; i++) tr = zadd(tr, Rn[i][i]);
  const zd dv = zscale(tr, 1.0 / (double)C);
  for (int i = 0; i < C; i++) for (int j = 0; j < C; j++) Rn[i][j] = zdiv(Rn[i][j], dv);
  for (int j = 0; j < C; j++) {
    double d = Rn[j][j].x;
    for (int q = 0; q < j; q++) d -= zabs2(Rn[j][q]);
    if (!(d > 0.0)) return false;
    const double l = sqrt(d);
    Rn[j][j] = zmk(l, 0.0);
    for (int i = j + 1; i < C; i++) {
      zd s = Rn[i][j];
      for (int q = 0; q < j; q++) s = zsub(s, zmulc(Rn[i][q], Rn[j][q]));
      Rn[i][j] = zscale(s, 1.0 / l);
    }
  }
  // forward-substitute the columns of Rt, conjugate-transpose, and once more: (L^-1 (L^-1 Rt)^H)^H = L^-1 Rt L^-H
  for (int pass = 0; pass < 2; pass++) {
    for (int j = 0; j < C; j++)
      for (int i = 0; i < C; i++) {
        zd s = Rt[i][j];
        for (int q = 0; q < i; q++) s = zsub(s, zmul(Rn[i][q], Rt[q][j]));
        Rt[i][j] = zscale(s, 1.0 / Rn[i][i].x);
      }
    for (int i = 0; i < C; i++) {
      Rt[i][i] = zconj(Rt[i][i]);
      for (int j = i + 1; j < C; j++) { const zd t = zconj(Rt[i][j]), s2 = zconj(Rt[j][i]); Rt[j][i] = t; Rt[i][j] = s2; }
    }
  }
  for (int i = 0; i < C; i++) {   // enforce the Hermitian symmetry the rounding left approximate
    Rt[i][i].y = 0.0;
    for (int j = i + 1; j < C; j++) { const zd m2 = zscale(zadd(Rt[i][j], zconj(Rt[j][i])), 0.5); Rt[i][j] = m2; Rt[j][i] = zconj(m2); }
  }
  return true;
}

// GEV stage 2: cyclic complex Jacobi on the Hermitian A (destroyed): A <- J^H A J, V <- V J; y = unit eigenvector of the
// largest eigenvalue.  Returns that eigenvalue.
template <int C>
BTKB_HD double jacobi_principal(zd (*A)[C], zd* y) {
  zd V[C][C];
  for (int i = 0; i < C; i++) for (int j = 0; j < C; j++) V[i][j] = zmk(i == j ? 1.0 : 0.0, 0.0);
  for (int sweep = 0; sweep < 30; sweep++) {
    double off = 0.0, dia = 0.0;
    for (int i = 0; i < C; i++) { dia += A[i][i].x * A[i][i].x; for (int j = i + 1; j < C; j++) off += 2.0 * zabs2(A[i][j]); }
    if (!(off > 1e-60 * fmax(dia, 1e-300))) break;
    // The (p, q) loops must stay rolled: fully unrolled (what nvcc 12.9 -O3 does for constant C), the device code produced
    // wrong eigenvalues for C >= 4 while -G, the host build and this rolled form agree to 1e-15 (tools/dbg/sos_math_device_check.cu).
#pragma unroll 1
    for (int p = 0; p < C - 1; p++) {
#pragma unroll 1
      for (int q = p + 1; q < C; q++) {
        const zd apq = A[p][q];
        const double gm = sqrt(zabs2(apq));
        const bool skip = (gm <= 1e-20 * (fabs(A[p][p].x) + fabs(A[q][q].x)) || gm < 1e-290);   // a rotation below rounding level
        if (!skip) {
          const zd ph = zscale(apq, 1.0 / gm);
          const double tau = (A[q][q].x - A[p][p].x) / (2.0 * gm);
          const double tt = (tau >= 0.0 ? 1.0 : -1.0) / (fabs(tau) + sqrt(1.0 + tau * tau));
          const double cs = 1.0 / sqrt(1.0 + tt * tt), sn2 = tt * cs;
          const zd sp = zscale(ph, sn2);          // J[p][q] = s ph, J[q][p] = -s conj(ph), J[p][p] = J[q][q] = c
          for (int i = 0; i < C; i++) {           // columns: A <- A J, V <- V J
            const zd aip = A[i][p], aiq = A[i][q];
            A[i][p] = zsub(zscale(aip, cs), zmul(aiq, zconj(sp)));
            A[i][q] = zadd(zmul(aip, sp), zscale(aiq, cs));
            const zd vip = V[i][p], viq = V[i][q];
            V[i][p] = zsub(zscale(vip, cs), zmul(viq, zconj(sp)));
            V[i][q] = zadd(zmul(vip, sp), zscale(viq, cs));
          }
          for (int j = 0; j < C; j++) {           // rows: A <- J^H A
            const zd apj = A[p][j], aqj = A[q][j];
            A[p][j] = zsub(zscale(apj, cs), zmul(sp, aqj));
            A[q][j] = zadd(zmul(zconj(sp), apj), zscale(aqj, cs));
          }
          A[p][p].y = 0.0; A[q][q].y = 0.0;
        }
        A[p][q] = zmk(0, 0); A[q][p] = zmk(0, 0);
      }
    }
  }
  int best = 0;
  for (int i = 1; i < C; i++) if (A[i][i].x > A[best][best].x) best = i;
  double nrm = 0.0;
  for (int i = 0; i < C; i++) nrm += zabs2(V[i][best]);
  nrm = 1.0 / sqrt(nrm);
  for (int i = 0; i < C; i++) y[i] = zscale(V[i][best], nrm);
  return A[best][best].x;
}

// GEV stage 3: v = L^-H y (v^H Rn v = 1, scipy.linalg.eigh's normalisation), phase: LAPACK's reduction (uplo = 'L') leaves
// (L[:,0])^H v real; the sign is taken positive here.
template <int C>
BTKB_HD void gev_back(zd (*L)[C], const zd* y, zd* w) {
  for (int i = C - 1; i >= 0; i--) {
    zd s = y[i];
    for (int q = i + 1; q < C; q++) s = zsub(s, zmul(zconj(L[q][i]), w[q]));
    w[i] = zscale(s, 1.0 / L[i][i].x);
  }
  zd ph = zmk(0, 0);
  for (int j = 0; j < C; j++) ph = zadd(ph, zmul(zconj(L[j][0]), w[j]));
  const double pa = sqrt(zabs2(ph));
  if (pa > 0.0) { const zd rot = zscale(zconj(ph), 1.0 / pa); for (int c = 0; c < C; c++) w[c] = zmul(w[c], rot); }
}

// blind MVDR: no = inv(Rn) Rt by LU with partial pivoting (numpy.linalg.inv = LAPACK getrf/getri); w = no[:, ref] / (offset + tr(no))
template <int C>
BTKB_HD bool bmvdr_solve(zd (*Rt)[C], zd (*Rn)[C], int ref_micx, double offset, zd* w) {
  for (int col = 0; col < C; col++) {
    int piv = col; double best = zabs2(Rn[col][col]);
    for (int r = col + 1; r < C; r++) { const double v = zabs2(Rn[r][col]); if (v > best) { best = v; piv = r; } }
    if (!(best > 0.0)) return false;
    if (piv != col)
      for (int j = 0; j < C; j++) { zd t = Rn[col][j]; Rn[col][j] = Rn[piv][j]; Rn[piv][j] = t; t = Rt[col][j]; Rt[col][j] = Rt[piv][j]; Rt[piv][j] = t; }
    for (int r = col + 1; r < C; r++) {
      const zd f = zdiv(Rn[r][col], Rn[col][col]);
      for (int j = col + 1; j < C; j++) Rn[r][j] = zsub(Rn[r][j], zmul(f, Rn[col][j]));
      for (int j = 0; j < C; j++) Rt[r][j] = zsub(Rt[r][j], zmul(f, Rt[col][j]));
    }
  }
  for (int j = 0; j < C; j++)
    for (int r = C - 1; r >= 0; r--) {
      zd s = Rt[r][j];
      for (int q = r + 1; q < C; q++) s = zsub(s, zmul(Rn[r][q], Rt[q][j]));
      Rt[r][j] = zdiv(s, Rn[r][r]);
    }
  zd tr = zmk(offset, 0.0);
  for (int i = 0; i < C; i++) tr = zadd(tr, Rt[i][i]);
  for (int c = 0; c < C; c++) w[c] = zdiv(Rt[c][ref_micx], tr);   // wqH = conj(no u / (offset + tr(no))), y = wqH . x = w^H x
  return true;
}

// Rt, Rn: normalised statistics of one (utterance, bin) (GEV: Rt unnormalised, :1320-1322), both destroyed.
// Returns false when a factorisation fails; w[C] with y = w^H x otherwise.
template <int C>
BTKB_HD bool sos_solve_chain(zd (*Rt)[C], zd (*Rn)[C], int kind, double gamma, int ref_micx, double offset, zd* w) {
  if (gamma > 0.0) improve_condition<C>(Rn, gamma);
  if (kind == BTKB_SOS_BMVDR) return bmvdr_solve<C>(Rt, Rn, ref_micx, offset, w);
  if (!gev_reduce<C>(Rt, Rn)) return false;
  zd y[C];
  jacobi_principal<C>(Rt, y);
  gev_back<C>(Rn, y, w);
  return true;
}

}  // namespace btkb
