// btkb_postfilter.cu — setup kernels of the McCowan and Lefkimmiatis post-filters (fp64, once per run, tiny grids).
//
// The per-frame work of both filters lives in k_perbin<C, MODE, PF> (btkb_perbin.cu).  What is computed here is everything
// that depends only on the noise coherence R_[fbinX] (one C x C matrix per bin, shared by every utterance of the batch):
//
//   k_pf_diffuse      McCowanPostFilter::set_diffuse_noise_model             postfilter/postfilter.cc:562-627
//   k_pf_diag_load    McCowanPostFilter::set_all_diagonal_loading            postfilter/postfilter.cc:629-642
//   k_pf_divide       McCowanPostFilter::divide_all_nondiagonal_elements     postfilter/postfilter.cc:662-680
//   k_pf_prepare      the R_ij-dependent factors of estimate_average_clean_PSD_ (:783-820) and
//                     estimate_average_noise_PSD_ (:1047-1087), and calc_inverse_noise_spatial_spectral_matrix (:966-978)
//   k_pf_lambda       LefkimmiatisPostFilter::calcLambda                     postfilter/postfilter.cc:980-992
//
// Per pair (i < j) the reference evaluates  (phi_ij - R'_ij (phi_ii + phi_jj)/2) / (1 - R'_ij)  every frame.  R'_ij is a
// constant of the bin, so with q'_ij = 1/(1 - R'_ij) and rho'_i = 1/2 sum_{j != i} R'_ij q'_ij the pair sum becomes
//   sum_{i<j} q'_ij phi_ij  -  sum_i rho'_i phi_ii ,
// and the noise PSD of Lefkimmiatis' filter likewise  sum_i rho''_i phi_ii - sum_{i<j} q''_ij phi_ij  with
// q''_ij = 1/(1 - R''_ij), rho''_i = 1/2 sum_{j != i} q''_ij  (R', R'' = R_ij after each function's own clipping rule).
// PFQ holds, per bin, [q'(NP) | rho'(C) | q''(NP) | rho''(C)] as complex64, laid out [entry][K] so a CTA of k_perbin reads
// consecutive bins coalesced.
#include "btkb_internal.h"
#include <math.h>

namespace btkb {
namespace {

constexpr int CMAX = 8;

struct cd { double x, y; };
__device__ __forceinline__ cd cdm(double x, double y) { cd r; r.x = x; r.y = y; return r; }
__device__ __forceinline__ cd cdadd(cd a, cd b) { return cdm(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ cd cdsub(cd a, cd b) { return cdm(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ cd cdmul(cd a, cd b) { return cdm(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
__device__ __forceinline__ cd cdconj(cd a) { return cdm(a.x, -a.y); }
__device__ __forceinline__ double cdabs2(cd a) { return a.x * a.x + a.y * a.y; }
__device__ __forceinline__ cd cdinv(cd a) { const double d = cdabs2(a); return cdm(a.x / d, -a.y / d); }

__global__ void k_pf_diffuse(const double* mpos, double2* Rpf, int C, int M, int K, double samplerate, double sspeed) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= K) return;
  const double omega_d_c = 2.0 * samplerate * (double)k / ((double)M * sspeed);
  for (int i = 0; i < C; i++)
    for (int j = 0; j < C; j++) {
      double val = 1.0;
      if (i != j) {
        const int a = i > j ? i : j, b = i > j ? j : i;  // the reference fills m > n and mirrors
        const double dx = mpos[a * 3] - mpos[b * 3], dy = mpos[a * 3 + 1] - mpos[b * 3 + 1], dz = mpos[a * 3 + 2] - mpos[b * 3 + 2];
        const double y = M_PI * omega_d_c * sqrt(dx * dx + dy * dy + dz * dz);
        val = (fabs(y) < 1e-12) ? 1.0 : sin(y) / y;  // gsl_sf_sinc(x) = sin(pi x) / (pi x)
      }
      Rpf[((size_t)k * C + i) * C + j] = make_double2(val, 0.0);
    }
}

__global__ void k_pf_diag_load(double2* Rpf, int C, int K, float mu) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= K) return;
  for (int c = 0; c < C; c++) Rpf[((size_t)k * C + c) * C + c].x += (double)mu;  // diagonal_weights_ is a float array
}

__global__ void k_pf_divide(double2* Rpf, int C, int K, float mu) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= K) return;
  const double den = 1.0 + (double)mu;  // (1.0 + mu) promotes the float to double
  for (int i = 0; i < C; i++)
    for (int j = 0; j < C; j++)
      if (i != j) { double2& r = Rpf[((size_t)k * C + i) * C + j]; r.x /= den; r.y /= den; }
}

// smallest singular value of a Hermitian C x C matrix = smallest |eigenvalue|, by cyclic Jacobi on the 2C x 2C real
// symmetric embedding [[Re A, -Im A], [Im A, Re A]] (same spectrum, every eigenvalue twice).
__device__ double min_abs_eig_hermitian(const cd* A, int C) {
  double S[2 * CMAX][2 * CMAX];
  const int n = 2 * C;
  for (int i = 0; i < C; i++)
    for (int j = 0; j < C; j++) {
      // Hermitian part (the coherence matrices of this path are Hermitian; this keeps S symmetric regardless)
      const double re = 0.5 * (A[i * C + j].x + A[j * C + i].x), im = 0.5 * (A[i * C + j].y - A[j * C + i].y);
      S[i][j] = re; S[i + C][j + C] = re; S[i][j + C] = -im; S[i + C][j] = im;
    }
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

__global__ void k_pf_prepare(const double2* Rpf, double2* invR, float2* PFQ, int C, int K, float threshold, double min_sv, int want_inverse) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= K) return;
  const int NP = C * (C - 1) / 2;
  const double thr = (double)threshold;  // threshold_of_Rij_ is a float member (postfilter.h)
  cd A[CMAX * CMAX];
  for (int i = 0; i < C * C; i++) { const double2 r = Rpf[(size_t)k * C * C + i]; A[i] = cdm(r.x, r.y); }
  cd rho1[CMAX], rho2[CMAX];
  for (int c = 0; c < C; c++) { rho1[c] = cdm(0, 0); rho2[c] = cdm(0, 0); }
  int idx = 0;
  for (int i = 0; i < C - 1; i++)
    for (int j = i + 1; j < C; j++) {
      const cd Rij = A[i * C + j];
      cd R1 = Rij;  // estimate_average_clean_PSD_ (:799-801)
      if (R1.x > thr && R1.y <= 0.0) R1 = cdm(thr, 0.0);
      cd R2 = Rij;  // estimate_average_noise_PSD_ (:1065-1070)
      if (R2.x > thr) R2 = cdm(thr, 0.0);
      else if (R2.x == 1.0) R2 = cdm(0.99, 0.0);
      const cd q1 = cdinv(cdsub(cdm(1.0, 0.0), R1));
      const cd q2 = cdinv(cdsub(cdm(1.0, 0.0), R2));
      const cd r1 = cdmul(R1, q1);
      rho1[i] = cdadd(rho1[i], cdm(0.5 * r1.x, 0.5 * r1.y)); rho1[j] = cdadd(rho1[j], cdm(0.5 * r1.x, 0.5 * r1.y));
      rho2[i] = cdadd(rho2[i], cdm(0.5 * q2.x, 0.5 * q2.y)); rho2[j] = cdadd(rho2[j], cdm(0.5 * q2.x, 0.5 * q2.y));
      PFQ[(size_t)idx * K + k] = make_float2((float)q1.x, (float)q1.y);
      PFQ[(size_t)(NP + C + idx) * K + k] = make_float2((float)q2.x, (float)q2.y);
      idx++;
    }
  for (int c = 0; c < C; c++) {
    PFQ[(size_t)(NP + c) * K + k] = make_float2((float)rho1[c].x, (float)rho1[c].y);
    PFQ[(size_t)(2 * NP + C + c) * K + k] = make_float2((float)rho2[c].x, (float)rho2[c].y);
  }
  if (!want_inverse) return;
  // pseudoinverse(R, invR, minSV): any singular value below minSV -> false -> identity (postfilter.cc:973-975, beamformer.cc:232-289)
  cd I[CMAX * CMAX];
  for (int i = 0; i < C; i++) for (int j = 0; j < C; j++) I[i * C + j] = cdm(i == j ? 1.0 : 0.0, 0.0);
  bool ok = !(min_abs_eig_hermitian(A, C) < min_sv);
  if (ok) {  // Gauss-Jordan with partial pivoting
    for (int col = 0; col < C && ok; col++) {
      int piv = col; double best = cdabs2(A[col * C + col]);
      for (int r = col + 1; r < C; r++) { const double v = cdabs2(A[r * C + col]); if (v > best) { best = v; piv = r; } }
      if (!(best > 1e-300)) { ok = false; break; }
      if (piv != col)
        for (int j = 0; j < C; j++) { cd t = A[col * C + j]; A[col * C + j] = A[piv * C + j]; A[piv * C + j] = t; t = I[col * C + j]; I[col * C + j] = I[piv * C + j]; I[piv * C + j] = t; }
      const cd ip = cdinv(A[col * C + col]);
      for (int j = 0; j < C; j++) { A[col * C + j] = cdmul(A[col * C + j], ip); I[col * C + j] = cdmul(I[col * C + j], ip); }
      for (int r = 0; r < C; r++) {
        if (r == col) continue;
        const cd f = A[r * C + col];
        for (int j = 0; j < C; j++) { A[r * C + j] = cdsub(A[r * C + j], cdmul(f, A[col * C + j])); I[r * C + j] = cdsub(I[r * C + j], cdmul(f, I[col * C + j])); }
      }
    }
  }
  for (int i = 0; i < C; i++)
    for (int j = 0; j < C; j++) {
      const cd v = ok ? I[i * C + j] : cdm(i == j ? 1.0 : 0.0, 0.0);
      invR[((size_t)k * C + i) * C + j] = make_double2(v.x, v.y);
    }
}

// Lambda = (invR^H d)^H d per chain (u, k); stored as the real number the filter divides by (Re or |.| by the type bit)
__global__ void k_pf_lambda(const double2* invR, const float2* TA, float* LAM, int U, int C, int K, int Gp, int pf_type) {
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= U * K) return;
  const int k = g % K;
  cd d[CMAX];
  for (int c = 0; c < C; c++) { const float2 t = TA[(size_t)c * Gp + g]; d[c] = cdm(t.x, t.y); }
  cd lam = cdm(0, 0);
  for (int i = 0; i < C; i++) {
    cd th = cdm(0, 0);  // (invR^H d)_i = sum_j conj(invR[j][i]) d_j
    for (int j = 0; j < C; j++) { const double2 r = invR[((size_t)k * C + j) * C + i]; th = cdadd(th, cdmul(cdm(r.x, -r.y), d[j])); }
    lam = cdadd(lam, cdmul(cdconj(th), d[i]));
  }
  LAM[g] = (float)((pf_type & 1) ? lam.x : sqrt(cdabs2(lam)));
}

}  // namespace

cudaError_t launch_pf_diffuse(const double* mpos, double2* Rpf, int C, int M, int K, double samplerate, double sspeed, cudaStream_t st) {
  k_pf_diffuse<<<(K + 63) / 64, 64, 0, st>>>(mpos, Rpf, C, M, K, samplerate, sspeed);
  return cudaGetLastError();
}
cudaError_t launch_pf_diag_load(double2* Rpf, int C, int K, float mu, cudaStream_t st) {
  k_pf_diag_load<<<(K + 63) / 64, 64, 0, st>>>(Rpf, C, K, mu);
  return cudaGetLastError();
}
cudaError_t launch_pf_divide_nondiag(double2* Rpf, int C, int K, float mu, cudaStream_t st) {
  k_pf_divide<<<(K + 63) / 64, 64, 0, st>>>(Rpf, C, K, mu);
  return cudaGetLastError();
}
cudaError_t launch_pf_prepare(const double2* Rpf, double2* invR, float2* PFQ, int C, int K, float threshold, double min_sv, int want_inverse, cudaStream_t st) {
  if (C > CMAX) return cudaErrorInvalidValue;
  k_pf_prepare<<<(K + 31) / 32, 32, 0, st>>>(Rpf, invR, PFQ, C, K, threshold, min_sv, want_inverse);
  return cudaGetLastError();
}
cudaError_t launch_pf_lambda(const double2* invR, const float2* TA, float* LAM, int U, int C, int K, int Gp, int pf_type, cudaStream_t st) {
  if (C > CMAX) return cudaErrorInvalidValue;
  const int n = U * K;
  k_pf_lambda<<<(n + 127) / 128, 128, 0, st>>>(invR, TA, LAM, U, C, K, Gp, pf_type);
  return cudaGetLastError();
}

}  // namespace btkb
