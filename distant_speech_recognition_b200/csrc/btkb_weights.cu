// btkb_weights.cu — setup-time kernels (once per look direction / per utterance, never per frame), sm_100a.
// One thread per (utterance, bin) chain g = u K + k; double precision throughout (the reference computes these in
// double; they are O(U K C^3) and invisible next to the data path).
//
//   k_mainlobe      BeamformerWeights::calcMainlobe            btk20_src/beamformer/beamformer.cc:502-565
//                   == calc_array_manifold_f                   btk20_src/lib/pybeamformer.py:284-306
//   blocking matrix calc_blocking_matrix_ / calc_blocking_matrix   beamformer.cc:373-454 / pybeamformer.py:309-341
//   k_blocking_wl   calcSidelobeCancellerP_f / U_f: wl = B wa   beamformer.cc:729-767
//   k_ua_to_wa      export of the NLMS state: waH = u conj(B)   (inverse of u = waH B^T)
//   k_diffuse       SubbandMVDR::set_diffuse_noise_model        beamformer.cc:2442-2509
//   k_mvdr_solve    set_all_diagonal_loading + calc_mvdr_weights beamformer.cc:2350-2402, 2511-2523
//                   (the reference inverts R with a single-precision LINPACK SVD, beamformer.cc:232-289; here a
//                   double-precision LU solve of R^H t = d — same mathematics, tighter rounding)
//   k_noise_mask    label / energy gating of accu_stats_from_label  pybeamformer.py:963-975
#include "btkb_internal.h"
#include "btkb_jacobi.cuh"
#include "btkb_cd.cuh"

namespace btkb {

constexpr int MAXC = 8;

__global__ void k_mainlobe(WeightsArgs a) {
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= a.U * a.K) return;
  const int u = g / a.K, k = g - u * a.K;
  const double fs = (double)a.samplerate;  // the C++ signature takes float
  for (int c = 0; c < a.C; c++) {
    const double tau = a.delays[(size_t)u * a.C + c];
    double val;
    if (k == 0) val = 0.0;
    else if (k == a.M / 2) val = -M_PI * fs * tau;
    else val = -2.0 * M_PI * (double)k * tau * fs / (double)a.M;
    double s, co;
    sincos(val, &s, &co);
    a.W[(size_t)c * a.Gp + g] = make_float2((float)(co / a.C), (float)(s / a.C));
  }
}

template <int C, int DIR>  // DIR 0: WL = B WA ; DIR 1: WA = UA conj(B)
__global__ void k_blocking(const float2* W, const float2* IN, float2* OUT, int U, int K, int Gp, int NC) {
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= U * K) return;
  cd v[C];
  cd B[C][C - 1];
  for (int c = 0; c < C; c++) { float2 t = W[(size_t)c * Gp + g]; v[c] = cdmake(t.x, t.y); }
  blocking_matrix<C>(v, B, NC);
  const int NA = C - NC;
  if (DIR == 0) {
    cd wa[C - 1];
    for (int i = 0; i < NA; i++) { float2 t = IN[(size_t)i * Gp + g]; wa[i] = cdmake(t.x, t.y); }
    for (int c = 0; c < C; c++) {
      cd s = cdmake(0, 0);
      for (int i = 0; i < NA; i++) s = cdadd(s, cdmul(B[c][i], wa[i]));
      OUT[(size_t)c * Gp + g] = make_float2((float)s.x, (float)s.y);
    }
  } else {
    cd ua[C];
    for (int c = 0; c < C; c++) { float2 t = IN[(size_t)c * Gp + g]; ua[c] = cdmake(t.x, t.y); }
    for (int i = 0; i < NA; i++) {
      cd s = cdmake(0, 0);
      for (int c = 0; c < C; c++) s = cdadd(s, cdmul(ua[c], cdconj(B[c][i])));
      OUT[(size_t)i * Gp + g] = make_float2((float)s.x, (float)s.y);
    }
  }
}

// Look-direction change in the middle of a stream (unit_test/test_online_beamforming.py:205-225 -> calc_beamformer_weights,
// pybeamformer.py:736-743 / 903-910): the reference keeps its adaptive state in the blocking-matrix basis — waH (NLMS, RLS) and the
// precision matrix Pz (RLS) — and simply builds new blocking matrices.  The kernels carry the same state in sensor space,
// u = waH B^T and Pt = conj(B) Pz B^T, so the state is re-expressed: waH = u conj(B_old), Pz = B_old^T Pt conj(B_old), then
// u = waH B_new^T, Pt = conj(B_new) Pz B_new^T.  ST rows: see k_perbin_rls (btkb_perbin.cu).
template <int C>
__global__ void k_adaptive_rebase(const float2* Wold, const float2* Wnew, float2* UA, float* ST, int has_P, int U, int K, int Gp) {
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= U * K) return;
  constexpr int NA = C - 1;
  cd v[C], Bo[C][C - 1], Bn[C][C - 1];
  for (int c = 0; c < C; c++) { float2 t = Wold[(size_t)c * Gp + g]; v[c] = cdmake(t.x, t.y); }
  blocking_matrix<C>(v, Bo, 1);
  for (int c = 0; c < C; c++) { float2 t = Wnew[(size_t)c * Gp + g]; v[c] = cdmake(t.x, t.y); }
  blocking_matrix<C>(v, Bn, 1);
  cd ua[C], wa[C - 1];
  for (int c = 0; c < C; c++) { float2 t = UA[(size_t)c * Gp + g]; ua[c] = cdmake(t.x, t.y); }
  for (int i = 0; i < NA; i++) {
    cd s = cdmake(0, 0);
    for (int c = 0; c < C; c++) s = cdadd(s, cdmul(ua[c], cdconj(Bo[c][i])));
    wa[i] = s;
  }
  for (int c = 0; c < C; c++) {
    cd s = cdmake(0, 0);
    for (int i = 0; i < NA; i++) s = cdadd(s, cdmul(Bn[c][i], wa[i]));
    UA[(size_t)c * Gp + g] = make_float2((float)s.x, (float)s.y);
  }
  if (!has_P) return;
  const size_t gp = (size_t)Gp;
  float* st = ST + g;
  cd P[C][C], T1[C][C - 1], Pz[C - 1][C - 1];
  for (int i = 0; i < C; i++) {
    P[i][i] = cdmake(st[(8 + i) * gp], 0.0);
    for (int j = 0; j < i; j++) {
      const int o = i * (i - 1) / 2 + j;
      P[i][j] = cdmake(st[(8 + C + 2 * o) * gp], st[(9 + C + 2 * o) * gp]);
      P[j][i] = cdconj(P[i][j]);
    }
  }
  for (int i = 0; i < C; i++)            // T1 = Pt conj(B_old)
    for (int a = 0; a < NA; a++) { cd s = cdmake(0, 0); for (int j = 0; j < C; j++) s = cdadd(s, cdmul(P[i][j], cdconj(Bo[j][a]))); T1[i][a] = s; }
  for (int a = 0; a < NA; a++)           // Pz = B_old^T T1
    for (int b = 0; b < NA; b++) { cd s = cdmake(0, 0); for (int i = 0; i < C; i++) s = cdadd(s, cdmul(Bo[i][a], T1[i][b])); Pz[a][b] = s; }
  for (int i = 0; i < C; i++)            // T1 = conj(B_new) Pz
    for (int b = 0; b < NA; b++) { cd s = cdmake(0, 0); for (int a = 0; a < NA; a++) s = cdadd(s, cdmul(cdconj(Bn[i][a]), Pz[a][b])); T1[i][b] = s; }
  for (int i = 0; i < C; i++)            // Pt = T1 B_new^T (Hermitian: lower triangle + real diagonal)
    for (int j = 0; j <= i; j++) {
      cd s = cdmake(0, 0);
      for (int b = 0; b < NA; b++) s = cdadd(s, cdmul(T1[i][b], Bn[j][b]));
      if (j == i) st[(8 + i) * gp] = (float)s.x;
      else { const int o = i * (i - 1) / 2 + j; st[(8 + C + 2 * o) * gp] = (float)s.x; st[(9 + C + 2 * o) * gp] = (float)s.y; }
    }
}
cudaError_t launch_adaptive_rebase(const float2* Wold, const float2* Wnew, float2* UA, float* ST, int has_P, int U, int C, int K, int Gp, cudaStream_t st) {
  const int n = U * K, bs = 64, gs = (n + bs - 1) / bs;
  switch (C) {
    case 2: k_adaptive_rebase<2><<<gs, bs, 0, st>>>(Wold, Wnew, UA, ST, has_P, U, K, Gp); break;
    case 3: k_adaptive_rebase<3><<<gs, bs, 0, st>>>(Wold, Wnew, UA, ST, has_P, U, K, Gp); break;
    case 4: k_adaptive_rebase<4><<<gs, bs, 0, st>>>(Wold, Wnew, UA, ST, has_P, U, K, Gp); break;
    case 5: k_adaptive_rebase<5><<<gs, bs, 0, st>>>(Wold, Wnew, UA, ST, has_P, U, K, Gp); break;
    case 6: k_adaptive_rebase<6><<<gs, bs, 0, st>>>(Wold, Wnew, UA, ST, has_P, U, K, Gp); break;
    case 7: k_adaptive_rebase<7><<<gs, bs, 0, st>>>(Wold, Wnew, UA, ST, has_P, U, K, Gp); break;
    case 8: k_adaptive_rebase<8><<<gs, bs, 0, st>>>(Wold, Wnew, UA, ST, has_P, U, K, Gp); break;
    default: return cudaErrorInvalidValue;
  }
  return cudaGetLastError();
}

template <int DIR>
static cudaError_t launch_blocking(const float2* W, const float2* IN, float2* OUT, int U, int C, int K, int Gp, int NC, cudaStream_t st) {
  const int n = U * K, bs = 128, gs = (n + bs - 1) / bs;
  switch (C) {
    case 2: k_blocking<2, DIR><<<gs, bs, 0, st>>>(W, IN, OUT, U, K, Gp, NC); break;
    case 3: k_blocking<3, DIR><<<gs, bs, 0, st>>>(W, IN, OUT, U, K, Gp, NC); break;
    case 4: k_blocking<4, DIR><<<gs, bs, 0, st>>>(W, IN, OUT, U, K, Gp, NC); break;
    case 5: k_blocking<5, DIR><<<gs, bs, 0, st>>>(W, IN, OUT, U, K, Gp, NC); break;
    case 6: k_blocking<6, DIR><<<gs, bs, 0, st>>>(W, IN, OUT, U, K, Gp, NC); break;
    case 7: k_blocking<7, DIR><<<gs, bs, 0, st>>>(W, IN, OUT, U, K, Gp, NC); break;
    case 8: k_blocking<8, DIR><<<gs, bs, 0, st>>>(W, IN, OUT, U, K, Gp, NC); break;
    default: return cudaErrorInvalidValue;
  }
  return cudaGetLastError();
}
// SubbandMVDRGSC::upgrade_blocking_matrix (beamformer.cc:2674-2691): the vector the blocking matrix is orthogonal to becomes
// wq - wl for the bins 1 .. (bin 0 keeps wq: the reference's loop starts at fbinX = 1); `unit` >= 0 instead writes the unit
// active-weight vector e_unit (rows of OUT = C - NC) used to read one column of B
__global__ void k_upgrade_source(const float2* WQ, const float2* WL, float2* OUT, int rows, int U, int K, int Gp, int unit) {
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= U * K) return;
  const int k = g % K;
  for (int c = 0; c < rows; c++) {
    float2 v;
    if (unit >= 0) v = make_float2(c == unit ? 1.f : 0.f, 0.f);
    else {
      v = WQ[(size_t)c * Gp + g];
      if (k > 0 && WL) { const float2 l = WL[(size_t)c * Gp + g]; v.x -= l.x; v.y -= l.y; }
    }
    OUT[(size_t)c * Gp + g] = v;
  }
}
cudaError_t launch_upgrade_source(const float2* WQ, const float2* WL, float2* OUT, int rows, int U, int K, int Gp, int unit, cudaStream_t st) {
  const int n = U * K, bs = 128;
  k_upgrade_source<<<(n + bs - 1) / bs, bs, 0, st>>>(WQ, WL, OUT, rows, U, K, Gp, unit);
  return cudaGetLastError();
}
cudaError_t launch_blocking_wl(const float2* W, const float2* WA, float2* WL, int U, int C, int K, int Gp, int NC, cudaStream_t st) {
  return launch_blocking<0>(W, WA, WL, U, C, K, Gp, NC, st);
}
cudaError_t launch_ua_to_wa(const float2* UA, const float2* W, float2* WA, int U, int C, int K, int Gp, cudaStream_t st) {
  return launch_blocking<1>(W, UA, WA, U, C, K, Gp, 1, st);
}

cudaError_t launch_mainlobe_weights(const WeightsArgs& a, cudaStream_t st) {
  const int n = a.U * a.K, bs = 128;
  k_mainlobe<<<(n + bs - 1) / bs, bs, 0, st>>>(a);
  return cudaGetLastError();
}

// Gamma_ij = sinc(2 fs k d_ij / (M c)) (GSL sinc(x) = sin(pi x)/(pi x)), diagonal 1.  mpos [C][3] doubles (mm).
__global__ void k_diffuse(const double* mpos, float2* R, int U, int C, int M, int K, int Gp, float samplerate, float sspeed) {
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= U * K) return;
  const int k = g % K;
  const double omega_d_c = 2.0 * (double)samplerate * (double)k / ((double)M * (double)sspeed);
  for (int i = 0; i < C; i++)
    for (int j = 0; j < C; j++) {
      double val = 1.0;
      if (i != j) {
        double dx = mpos[i * 3] - mpos[j * 3], dy = mpos[i * 3 + 1] - mpos[j * 3 + 1], dz = mpos[i * 3 + 2] - mpos[j * 3 + 2];
        double x = omega_d_c * sqrt(dx * dx + dy * dy + dz * dz);
        double y = M_PI * x;
        val = (fabs(y) < 1e-12) ? 1.0 : sin(y) / y;
      }
      R[(size_t)(i * C + j) * Gp + g] = make_float2((float)val, 0.f);
    }
}
cudaError_t launch_diffuse_model(const double* mpos, float2* R, int U, int C, int M, int K, int Gp, float samplerate, float sspeed, cudaStream_t st) {
  const int n = U * K, bs = 128;
  k_diffuse<<<(n + bs - 1) / bs, bs, 0, st>>>(mpos, R, U, C, M, K, Gp, samplerate, sspeed);
  return cudaGetLastError();
}

// w = (R^H)^-1 d / (C d^H R^-1 d), bin 0: all ones (beamformer.cc:2369-2371).  R row-major [i*C+j][g].
template <int C>
__global__ void k_mvdr_solve(const float2* R, const float2* Dm, float2* W, const int* noise_count, int U, int K, int Gp, float mu, int normalize, float dthreshold) {
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= U * K) return;
  const int u = g / K, k = g - u * K;
  if (k == 0) {
    for (int c = 0; c < C; c++) W[(size_t)c * Gp + g] = make_float2(1.f, 0.f);
    return;
  }
  double scale = 1.0;
  if (normalize && noise_count != nullptr && noise_count[u] > 0) scale = 1.0 / (double)noise_count[u];
  // A = R^H (conjugate transpose), with finalize_stats scaling and diagonal loading applied to R first
  cd A[C][C];
  cd b[C], d[C];
  for (int i = 0; i < C; i++)
    for (int j = 0; j < C; j++) {
      float2 t = R[(size_t)(i * C + j) * Gp + g];
      cd rij = cdmake((double)t.x * scale, (double)t.y * scale);
      if (i == j) rij.x += (double)mu;  // set_all_diagonal_loading: float diagonalWeight
      A[j][i] = cdconj(rij);
    }
  for (int c = 0; c < C; c++) { float2 t = Dm[(size_t)c * Gp + g]; d[c] = cdmake(t.x, t.y); b[c] = d[c]; }
  // LU with partial pivoting, in place (multipliers below the diagonal, rows physically swapped, perm[r] = original row now at r)
  int perm[C];
  for (int i = 0; i < C; i++) perm[i] = i;
  bool singular = false;
  for (int col = 0; col < C; col++) {
    int piv = col; double best = cdabs2(A[col][col]);
    for (int r = col + 1; r < C; r++) { double v = cdabs2(A[r][col]); if (v > best) { best = v; piv = r; } }
    if (!(best > 1e-60)) { singular = true; break; }
    if (piv != col) {
      for (int j = 0; j < C; j++) { cd t = A[col][j]; A[col][j] = A[piv][j]; A[piv][j] = t; }
      cd t = b[col]; b[col] = b[piv]; b[piv] = t;
      const int pi = perm[col]; perm[col] = perm[piv]; perm[piv] = pi;
    }
    for (int r = col + 1; r < C; r++) {
      cd f = cddiv(A[r][col], A[col][col]);
      for (int j = col + 1; j < C; j++) A[r][j] = cdsub(A[r][j], cdmul(f, A[col][j]));
      b[r] = cdsub(b[r], cdmul(f, b[col]));
      A[r][col] = f;
    }
  }
  // pseudoinverse(R, invR, dThreshold) reports failure when a singular value is below dThreshold and the caller then uses the
  // identity (beamformer.cc:267-274, 2381-2383): same rule here, on the smallest singular value of the loaded matrix.  The exact value
  // (Jacobi, btkb_jacobi.cuh) is needed only when the threshold lies inside the bracket the LU factors give for free,
  //     1 / ||A^-1||_F  <=  sigma_min  <=  sqrt(C) / ||A^-1||_F
  // (C extra pairs of triangular solves; the Jacobi of every chain made configs[2] ten times slower than round 1: 55 ms against 5.7).
  if (!singular && dthreshold > 0.f) {
    double inv_f2 = 0.0;
    for (int i = 0; i < C; i++) {
      cd y[C];
      for (int r = 0; r < C; r++) y[r] = cdmake(perm[r] == i ? 1.0 : 0.0, 0.0);
      for (int col = 0; col < C; col++)
        for (int r = col + 1; r < C; r++) y[r] = cdsub(y[r], cdmul(A[r][col], y[col]));
      for (int r = C - 1; r >= 0; r--) {
        cd sacc = y[r];
        for (int j = r + 1; j < C; j++) sacc = cdsub(sacc, cdmul(A[r][j], y[j]));
        y[r] = cddiv(sacc, A[r][r]);
        inv_f2 += cdabs2(y[r]);
      }
    }
    const double thr = (double)dthreshold;
    const double lo = (inv_f2 > 0.0) ? 1.0 / sqrt(inv_f2) : 0.0;      // NaN / inf from a numerically singular U land in the exact test too
    const double hi = sqrt((double)C) * lo;
    if (hi < thr) singular = true;
    else if (!(lo >= thr)) {
      cd Ao[C][C];
      for (int i = 0; i < C; i++)
        for (int j = 0; j < C; j++) {
          float2 t = R[(size_t)(i * C + j) * Gp + g];
          cd rij = cdmake((double)t.x * scale, (double)t.y * scale);
          if (i == j) rij.x += (double)mu;
          Ao[j][i] = cdconj(rij);
        }
      singular = min_singular_value(&Ao[0][0], C) < thr;
    }
  }
  cd tvec[C];
  if (!singular) {
    for (int r = C - 1; r >= 0; r--) {
      cd s = b[r];
      for (int j = r + 1; j < C; j++) s = cdsub(s, cdmul(A[r][j], tvec[j]));
      tvec[r] = cddiv(s, A[r][r]);
    }
  } else {
    for (int c = 0; c < C; c++) tvec[c] = d[c];  // identity fallback (beamformer.cc:2381-2383)
  }
  cd lam = cdmake(0, 0);  // Lambda = tmpH^H d
  for (int c = 0; c < C; c++) lam = cdadd(lam, cdmul(cdconj(tvec[c]), d[c]));
  cd norm = cdscale(lam, (double)C);
  for (int c = 0; c < C; c++) { cd wv = cddiv(tvec[c], norm); W[(size_t)c * Gp + g] = make_float2((float)wv.x, (float)wv.y); }
}
cudaError_t launch_mvdr_solve(const float2* R, const float2* D, float2* W, const int* noise_count, int U, int C, int K, int Gp, float mu,
                              int normalize_by_count, float dthreshold, cudaStream_t st) {
  const int n = U * K, bs = 64, gs = (n + bs - 1) / bs;
  switch (C) {
    case 2: k_mvdr_solve<2><<<gs, bs, 0, st>>>(R, D, W, noise_count, U, K, Gp, mu, normalize_by_count, dthreshold); break;
    case 3: k_mvdr_solve<3><<<gs, bs, 0, st>>>(R, D, W, noise_count, U, K, Gp, mu, normalize_by_count, dthreshold); break;
    case 4: k_mvdr_solve<4><<<gs, bs, 0, st>>>(R, D, W, noise_count, U, K, Gp, mu, normalize_by_count, dthreshold); break;
    case 5: k_mvdr_solve<5><<<gs, bs, 0, st>>>(R, D, W, noise_count, U, K, Gp, mu, normalize_by_count, dthreshold); break;
    case 6: k_mvdr_solve<6><<<gs, bs, 0, st>>>(R, D, W, noise_count, U, K, Gp, mu, normalize_by_count, dthreshold); break;
    case 7: k_mvdr_solve<7><<<gs, bs, 0, st>>>(R, D, W, noise_count, U, K, Gp, mu, normalize_by_count, dthreshold); break;
    case 8: k_mvdr_solve<8><<<gs, bs, 0, st>>>(R, D, W, noise_count, U, K, Gp, mu, normalize_by_count, dthreshold); break;
    default: return cudaErrorInvalidValue;
  }
  return cudaGetLastError();
}

// mask[t][u] = frame t of utterance u is a noise frame with energy > thr (pybeamformer.py:963-975); count[u] = their number
__global__ void k_noise_mask(const float* E, const int* lengths, const double* labels, unsigned char* mask, int* count, int U, int T, int D, int laN,
                             int pdA, float samplerate, float thr) {
  const int u = blockIdx.x * blockDim.x + threadIdx.x;
  if (u >= U) return;
  const int Tu = frames_of(lengths[u], D, laN, pdA);
  double elapsed = 0.0;
  const double dt = (double)D / (double)samplerate;
  int labx = 0, n = 0;
  const double s = labels ? labels[2 * u] : 0.0, e = labels ? labels[2 * u + 1] : 0.0;
  for (int t = 0; t < T; t++) {
    bool is_target = false;
    if (labels != nullptr && labx < 1) {
      if (elapsed >= s && (elapsed <= e || e < 0)) is_target = true;
      else if (elapsed > e) labx += 1;
    }
    const bool mk = (t < Tu) && !is_target && (E[(size_t)t * U + u] > thr);
    mask[(size_t)t * U + u] = mk ? 1 : 0;
    n += mk ? 1 : 0;
    elapsed += dt;
  }
  count[u] = n;
}
cudaError_t launch_noise_mask(const float* E, const int* lengths, const double* labels, unsigned char* mask, int* count, int U, int T, int D, int laN,
                              int pdA, float samplerate, float thr, cudaStream_t st) {
  k_noise_mask<<<(U + 63) / 64, 64, 0, st>>>(E, lengths, labels, mask, count, U, T, D, laN, pdA, samplerate, thr);
  return cudaGetLastError();
}

}  // namespace btkb

// ---------------------------------------------------------------------------------------------------------------
// LCMV quiescent weights: BeamformerWeights::calcMainlobeN / calcMainlobe2 (beamformer.cc:573-721) with
// calc_null_beamformer_ (beamformer.cc:299-363): wq = Cm (Cm^H Cm)^-1 g, Cm = [target | jammers] (unit modulus), g = e0.
// NC == 2 inverts with calc_inverse_22mat_ (beamformer.cc:181-221: diagonal +0.01 when |det| < 1e-7); NC > 2 uses the
// reference's pseudoinverse (float SVD there, a double LU inverse here).  The f = M/2 branch of the reference calls
// calc_null_beamformer_ INSIDE its channel loop with the jammer manifolds left over from bin M/2-1
// (beamformer.cc:677-689); that cascade is reproduced literally so the weights match the reference bin for bin.
namespace btkb {

constexpr int MAXNC = 4;

template <int C>
__device__ void null_beamformer(cd* wt, const cd (*Wj)[C], int NC) {
  cd A[MAXNC][MAXNC], inv[MAXNC][MAXNC];
  auto col = [&](int j, int c) -> cd { return j == 0 ? wt[c] : Wj[j - 1][c]; };
  for (int i = 0; i < NC; i++)
    for (int j = 0; j < NC; j++) {
      cd s = cdmake(0, 0);
      for (int c = 0; c < C; c++) s = cdadd(s, cdmul(cdconj(col(i, c)), col(j, c)));
      A[i][j] = s;
    }
  if (NC == 2) {
    cd det = cdsub(cdmul(A[0][0], A[1][1]), cdmul(A[0][1], A[1][0]));
    if (sqrt(cdabs2(det)) < 1.0e-7) {
      A[0][0].x += 0.01; A[1][1].x += 0.01;
      det = cdsub(cdmul(A[0][0], A[1][1]), cdmul(A[0][1], A[1][0]));
    }
    inv[0][0] = cddiv(A[1][1], det); inv[1][1] = cddiv(A[0][0], det);
    inv[0][1] = cdscale(cddiv(A[0][1], det), -1.0); inv[1][0] = cdscale(cddiv(A[1][0], det), -1.0);
  } else {
    // Gauss-Jordan inverse with partial pivoting
    cd aug[MAXNC][2 * MAXNC];
    for (int i = 0; i < NC; i++) for (int j = 0; j < NC; j++) { aug[i][j] = A[i][j]; aug[i][NC + j] = cdmake(i == j ? 1.0 : 0.0, 0.0); }
    for (int cI = 0; cI < NC; cI++) {
      int piv = cI; double best = cdabs2(aug[cI][cI]);
      for (int r = cI + 1; r < NC; r++) if (cdabs2(aug[r][cI]) > best) { best = cdabs2(aug[r][cI]); piv = r; }
      if (piv != cI) for (int j = 0; j < 2 * NC; j++) { cd t = aug[cI][j]; aug[cI][j] = aug[piv][j]; aug[piv][j] = t; }
      cd d = aug[cI][cI];
      for (int j = 0; j < 2 * NC; j++) aug[cI][j] = cddiv(aug[cI][j], d);
      for (int r = 0; r < NC; r++) if (r != cI) { cd f = aug[r][cI]; for (int j = 0; j < 2 * NC; j++) aug[r][j] = cdsub(aug[r][j], cdmul(f, aug[cI][j])); }
    }
    for (int i = 0; i < NC; i++) for (int j = 0; j < NC; j++) inv[i][j] = aug[i][NC + j];
  }
  cd out[C];
  for (int c = 0; c < C; c++) {
    cd s = cdmake(0, 0);
    for (int j = 0; j < NC; j++) s = cdadd(s, cdmul(col(j, c), inv[j][0]));  // Cm (inv g), g = e0
    out[c] = s;
  }
  for (int c = 0; c < C; c++) wt[c] = out[c];
}

template <int C>
__global__ void k_lcmv(const double* delaysT, const double* delaysJ, float2* W, int U, int NC, int M, int K, int Gp, float samplerate) {
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= U * K) return;
  const int u = g / K, k = g - u * K;
  const double fs = (double)samplerate;
  const double* dT = delaysT + (size_t)u * C;
  const double* dJ = delaysJ + (size_t)u * (NC - 1) * C;
  cd wt[C]; cd Wj[MAXNC - 1][C];
  auto polar = [](double a) { double s, c; sincos(a, &s, &c); return cdmake(c, s); };
  if (k == 0) {
    for (int c = 0; c < C; c++) wt[c] = cdmake(1.0 / C, 0.0);
  } else if (k < M / 2) {
    for (int c = 0; c < C; c++) {
      wt[c] = polar(-2.0 * M_PI * (double)k * dT[c] * fs / (double)M);
      for (int n = 0; n < NC - 1; n++) Wj[n][c] = polar(-2.0 * M_PI * (double)k * fs * dJ[n * C + c] / (double)M);
    }
    null_beamformer<C>(wt, Wj, NC);
  } else {
    // f = M/2: start from calcMainlobe's value; pWj still holds the jammer manifolds of bin M/2 - 1
    for (int c = 0; c < C; c++) {
      wt[c] = cdscale(polar(-M_PI * fs * dT[c]), 1.0 / C);
      for (int n = 0; n < NC - 1; n++) Wj[n][c] = polar(-2.0 * M_PI * (double)(M / 2 - 1) * fs * dJ[n * C + c] / (double)M);
    }
    for (int c = 0; c < C; c++) {
      wt[c] = cdscale(wt[c], (double)C);
      for (int n = 0; n < NC - 1; n++) wt[c] = cdscale(polar(-M_PI * fs * dJ[n * C + c]), 1.0 / C);
      null_beamformer<C>(wt, Wj, NC);
    }
  }
  for (int c = 0; c < C; c++) W[(size_t)c * Gp + g] = make_float2((float)wt[c].x, (float)wt[c].y);
}

cudaError_t launch_lcmv_weights(const double* delaysT, const double* delaysJ, float2* W, int U, int C, int NC, int M, int K, int Gp, float samplerate,
                                cudaStream_t st) {
  if (NC < 2 || NC > MAXNC || NC > C) return cudaErrorInvalidValue;
  const int n = U * K, bs = 64, gs = (n + bs - 1) / bs;
  switch (C) {
    case 2: k_lcmv<2><<<gs, bs, 0, st>>>(delaysT, delaysJ, W, U, NC, M, K, Gp, samplerate); break;
    case 3: k_lcmv<3><<<gs, bs, 0, st>>>(delaysT, delaysJ, W, U, NC, M, K, Gp, samplerate); break;
    case 4: k_lcmv<4><<<gs, bs, 0, st>>>(delaysT, delaysJ, W, U, NC, M, K, Gp, samplerate); break;
    case 5: k_lcmv<5><<<gs, bs, 0, st>>>(delaysT, delaysJ, W, U, NC, M, K, Gp, samplerate); break;
    case 6: k_lcmv<6><<<gs, bs, 0, st>>>(delaysT, delaysJ, W, U, NC, M, K, Gp, samplerate); break;
    case 7: k_lcmv<7><<<gs, bs, 0, st>>>(delaysT, delaysJ, W, U, NC, M, K, Gp, samplerate); break;
    case 8: k_lcmv<8><<<gs, bs, 0, st>>>(delaysT, delaysJ, W, U, NC, M, K, Gp, samplerate); break;
    default: return cudaErrorInvalidValue;
  }
  return cudaGetLastError();
}

}  // namespace btkb
