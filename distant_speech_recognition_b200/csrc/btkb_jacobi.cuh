// btkb_jacobi.cuh — smallest singular value of a small complex matrix (C <= 8), fp64, one thread per matrix.
// The reference's pseudoinverse (btk20_src/beamformer/beamformer.cc:232-289) takes a LINPACK SVD and reports failure when any
// singular value is below dThreshold; its callers then fall back to the identity (beamformer.cc:2381-2383, postfilter.cc:973-975).
// Here: cyclic Jacobi on the 2n x 2n real symmetric embedding [[Re H, -Im H], [Im H, Re H]] of a Hermitian H (same spectrum, every
// eigenvalue twice).  A Hermitian A (every matrix the path builds: coherence, covariance, loaded versions of them) goes in as is,
// s_min = min |lambda|; a general A goes in as H = A^H A, s_min = sqrt(min lambda).
#pragma once
#include <cuda_runtime.h>

namespace btkb {

constexpr int JACOBI_CMAX = 8;

__device__ inline double jacobi_min_abs_eig_embedded(double (*S)[2 * JACOBI_CMAX], int n) {
  for (int sweep = 0; sweep < 40; sweep++) {
    double off = 0.0, dia = 0.0;
    for (int p = 0; p < n; p++) { dia += S[p][p] * S[p][p]; for (int q = p + 1; q < n; q++) off += S[p][q] * S[p][q]; }
    if (off <= 1e-32 * dia || off == 0.0) break;
    for (int p = 0; p < n - 1; p++)
      for (int q = p + 1; q < n; q++) {
        const double apq = S[p][q];
        if (fabs(apq) < 1e-300) continue;
        const double theta = (S[q][q] - S[p][p]) / (2.0 * apq);
        const double t = (theta >= 0.0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
        const double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
        for (int r = 0; r < n; r++) {
          if (r == p || r == q) continue;
          const double arp = S[r][p], arq = S[r][q];
          S[r][p] = S[p][r] = c * arp - s * arq;
          S[r][q] = S[q][r] = s * arp + c * arq;
        }
        S[p][p] -= t * apq; S[q][q] += t * apq; S[p][q] = S[q][p] = 0.0;
      }
  }
  double mn = fabs(S[0][0]);
  for (int p = 1; p < n; p++) mn = fmin(mn, fabs(S[p][p]));
  return mn;
}

// A: row-major C x C array of a struct with double members x (re), y (im)
template <class CD>
__device__ inline double min_singular_value(const CD* A, int C) {
  double S[2 * JACOBI_CMAX][2 * JACOBI_CMAX];
  double asym = 0.0, nrm = 0.0;
  for (int i = 0; i < C; i++)
    for (int j = 0; j < C; j++) {
      const double dr = A[i * C + j].x - A[j * C + i].x, di = A[i * C + j].y + A[j * C + i].y;
      asym += dr * dr + di * di; nrm += A[i * C + j].x * A[i * C + j].x + A[i * C + j].y * A[i * C + j].y;
    }
  const bool hermitian = asym <= 1e-24 * nrm;
  for (int i = 0; i < C; i++)
    for (int j = 0; j < C; j++) {
      double re, im;
      if (hermitian) { re = 0.5 * (A[i * C + j].x + A[j * C + i].x); im = 0.5 * (A[i * C + j].y - A[j * C + i].y); }
      else {   // (A^H A)_ij = sum_k conj(A_ki) A_kj
        re = 0.0; im = 0.0;
        for (int k = 0; k < C; k++) { re += A[k * C + i].x * A[k * C + j].x + A[k * C + i].y * A[k * C + j].y; im += A[k * C + i].x * A[k * C + j].y - A[k * C + i].y * A[k * C + j].x; }
      }
      S[i][j] = re; S[i + C][j + C] = re; S[i][j + C] = -im; S[i + C][j] = im;
    }
  const double m = jacobi_min_abs_eig_embedded(S, 2 * C);
  return hermitian ? m : sqrt(m);
}

}  // namespace btkb
