// btkb_internal.h — argument blocks and launch prototypes shared by the kernels and the C-ABI (not installed).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <cstdlib>

namespace btkb {

// ------------------------------------------------------------------ HBM layout
// samples  x  [U][C][n_stride]        float32   (time-major per channel; n_stride multiple of 4)
// snapshots X [T][C][Gp]              complex64, g = u*K + k  (bin fastest, utterances interleaved, frame-major slabs)
// output   Y  [T][Gp]                 complex64
// energy   E  [T][U]                  float32   channel-0 frame energy |x0^H x0| / M
// weights  W  [C][Gp]                 complex64 (wq / wmvdr), WL same shape (wl = B wa)
// time     out[U][nb_stride]          float32

struct AnalysisArgs {
  const float* x; const int* lengths; const float* h;
  float2* X; float* E;
  int U, C, n, n_stride, T, M, m, D, laN, Gp, gain;
  const float2* twtab;   // exp(+2 pi i n / M), n < M (device)
  int tiles_per_cta;     // set by the launcher
  int debug;             // BTKB_ANALYSIS_DEBUG bit mask (profiling experiments only): 1 skip X stores, 2 skip fold, 4 skip FFT passes
  // Index of the sample just past local frame 0's window, in the coordinates of the staged sample rows: (laN + 1) D for a whole
  // utterance; for a streamed chunk (btkb_stream_submit) (laN + t_base + 1) D - s_base, where t_base is the absolute number of the
  // chunk's first frame and s_base the absolute index of the first sample in the row (history prefix of m M - D samples).
  // `lengths` counts the valid samples of the ROW (history included); X / E rows are chunk-local.
  long long w_base;
  // Frames ride the transforms in PAIRS (two real channels x two frames), and floating-point rounding makes a frame's result depend
  // on which frame it is paired with.  A streamed chunk that starts at an odd absolute frame number therefore starts its first tile
  // one frame early (t_skip = 1; that frame is computed and discarded), so that pairs are the whole-utterance run's pairs.
  int t_skip;
  // 16-bit PCM input (btkb_submit_i16): rows [U][C][n16_stride] int16, n16_stride a multiple of 8.  When set, the r = 1 fast path reads it
  // directly (k_analysis_r1<..., I16 = true>): the two channels of a pair are staged as ONE 32-bit word per sample, which halves the
  // shared-memory wavefronts of the polyphase fold; int16 -> fp32 is exact, so the result equals the float path bit for bit.
  const int16_t* x16; int n16_stride;
  int Crow;   // channel rows per frame of X (>= C: wide arrays are padded with all-zero channels to 16 / 32 / 64 rows, btkb_api.cu)
};

struct SynthesisArgs {
  const float2* Y; const int* lengths; const float* g;
  float* out; double* stats;  // stats [U][3]: sum of squares accumulated into [u][0]
  int U, n, T, M, m, r, D, K, Gp, pdS, laN, pdA, nb, nb_stride, gain;
  const float2* twtab;   // exp(+2 pi i n / M), n < M (device)
  int onesided;          // frames tau < onesided carry bins 0..M/2 only (upper half zero): the McCowan / Lefkimmiatis post-filters
                         // leave vector_[M/2+1..M-1] untouched until frame_no_ >= min_frames (postfilter.cc:896-901)
  // streamed chunks (btkb_stream_submit): b_base = absolute number of the chunk's first output block, y_base = absolute frame number
  // of row 0 of Y (the m R - 1 frames of history a block reaches back to precede the chunk's own frames), tu[u] = absolute number of
  // frames utterance u has produced so far (null: whole utterances, derived from `lengths`).  nb / out rows are chunk-local.
  int b_base, y_base; const int* tu;
  int b_skip;   // 1: the chunk starts at an odd absolute block; tiles start one block early so that frame pairs match the whole run's
};

struct LmsArgs {
  float beta, gamma, init_diagonal_load, regularization_param, energy_floor, sil_thresh, max_wa_l2norm;
  int min_frames, slowdown_after;
};

struct RlsArgs {
  float beta, gamma, mu, init_diagonal_load, regularization_param, sil_thresh, alpha2, max_wa_l2norm;
  int constraint_option, min_frames;
};

struct RlsCppArgs { float mu, sigma2, init_sigma2, alpha; int qctype, update; };   // SubbandGSCRLS (C++), btkb_rls_cpp.cu

struct PerBinArgs {
  const float2* X; const float* E; const int* lengths;
  const float2* W;    // [C][Gp] quiescent / mvdr weights (not conjugated)
  const float2* WL;   // [C][Gp] wl = B wa (may be null)
  const float2* TA;   // [C][Gp] time-alignment manifold for the post-filter (D&S weights)
  float2* Y;          // [T][Gp]
  float* PFW;         // [T][Gp] post-filter gains (may be null)
  float2* UA;         // [C][Gp] NLMS state u = waH B^T (in/out, final value written)
  float* stats_updates;  // [U] NLMS update counts (written by the thread owning bin 0)
  // covariance accumulation
  float2* R;          // [C*C][Gp] (row-major i*C+j) accumulated x_i conj(x_j)
  const unsigned char* noise_mask;  // [T][U] 1 = accumulate this frame
  int* noise_count;   // [U]
  int U, C, T, M, K, G, Gp, D, laN, pdA;
  int kind;           // BTKB_BF_*
  int normalize_weight;  // calc_gsc_output's w <- w / (||w|| C) (beamformer.cc:1230-1236)
  int pf_kind; float pf_alpha; int pf_type, pf_min_frames;
  // McCowan / Lefkimmiatis constants (btkb_postfilter.cu): PFQ [NQ][K] per-bin pair factors, LAM [Gp] Lambda = d^H R^-1 d
  const float2* PFQ; const float* LAM; int pf_fbin1;
  LmsArgs lms;
  RlsArgs rls;
  RlsCppArgs rlsc;
  float energy_threshold;
  float2* Scov; int Ts;   // 64-mic tensor-core covariance: series-major workspace S[g][c][Ts] (btkb_cov_tc.cu)
  // streamed chunks (btkb_stream_submit): t_base = absolute number of the chunk's first frame (the recurrences test absolute frame
  // numbers: first frame, min_frames, step-size halving, the post-filters' first two frames), tu[u] = absolute frames of utterance u
  // so far, ST = carried per-chain state [ST_ROWS][Gp] (always stored when non-null, loaded when st_load; u rides in UA).
  int t_base; const int* tu; float* ST; int st_load;
  int Ctrue;  // microphones really present (C counts the zero-padded channel rows of a wide array; the projector step scales by Ctrue)
};
constexpr int ST_ROWS = 72;   // 8 scalars + the largest state: RLS at C = 8 (8 + 56) / Zelinski at C = 8 (56 + 8)

// multi-channel WPE (btkb_wpe.cu)
struct WpeArgs {
  float2* X;             // [T][C][Gp] snapshots: reverberant in, dereverberated out
  const int* lengths;
  float2* S;             // [G][C][Ts] series-major copy of X
  float* TH;             // [G][C][Ts] theta
  float2* Gf;            // [G][C][L] prediction filters (chains outside the estimated band stay zero)
  void* Rw;              // workspace of matrix slots, complex128 (or complex64), `slot` elements each.  Lag-domain form: [chunk][C] slots,
                         // [L+1][Lr]: lower triangle of R_c, row L = conj(r_c).  Frame-domain form: [chunk][C+1] slots, [Sd+1][Lr]: the
                         // C loaded systems of a problem and, in slot C, the Gram matrix K they share
  size_t slot;           // elements per slot (>= (L+1) Lr; >= (Sd+1) Lr when the frame-domain form may be used)
  int form;              // 0: lag-domain normal equations (L x L), 1: frame-domain (S x S, S = estimation frames - lower)
  int Sd;                // frame-domain form: largest S in the batch
  int chunk_frame;       // host side: problems per launch in the frame-domain form (`chunk` of launch_wpe serves the lag-domain form)
  int chol_threads;      // host side: CTA size of k_wpe_chol (0: 256 lag-domain, 128 frame-domain)
  int prefetch;          // k_wpe_chol: L1 prefetch of the trailing entries ahead of their update
  int mma;               // k_wpe_chol (fp64): trailing update on the fp64 tensor cores (mma.sync m8n8k4)
  int* err_flag;         // set to 1 when a Cholesky pivot is not positive
  int U, C, T, Ts, K, G, Gp, D, laN, pdA;
  int lowerN, P, L, Lr, iterations, nbins, est_frames;
  float load_factor, diagonal_bias;
  int apply_only;        // keep the filters Gf of an earlier estimation and run the output stage only
};
size_t wpe_workspace_bytes(int C, size_t slot, int chunk, int fp32);
cudaError_t launch_wpe(const WpeArgs& a, int chunk, int fp32, cudaStream_t st, int* launches);

// SOS batch beamformers (btkb_sos.cu)
struct SosArgs {
  const float2* X; const float* E; const int* lengths;
  const double* labels; int NL;        // [U][NL][2] VAD segments in seconds (null: TF-mask mode)
  const float* mask_t; const float* mask_j;   // [T][Gp] masks in device layout (null: label mode)
  float* wtu;                          // [2][T][U] frame weights: target / noise class x energy gate
  double2* Rs;                         // [2][C*C][Gp] target / noise covariance sums (row-major i*C+j)
  double* cnt;                         // [2][Gp] frame counts
  double2* Wd;                         // [C][Gp] fp64 weights before the GEV phase alignment
  float2* W;                           // [C][Gp] beamformer weights (y = W^H x)
  int* err;                            // error bits (see k_sos_solve)
  int accumulate;                      // add to the statistics already there (accu_stats_* called again, :1113-1127)
  int U, C, T, K, G, Gp, D, laN, pdA;
  float samplerate, thr;
};
cudaError_t launch_sos_scatter_mask(const float* src, float* dst, int U, int Tm, int T, int K, int Gp, cudaStream_t st);
cudaError_t launch_sos_accumulate(const SosArgs& a, cudaStream_t st, int* launches);
cudaError_t launch_sos_solve(const SosArgs& a, int kind, double gamma, int ref_micx, double offset, cudaStream_t st, int* launches);

struct WeightsArgs {
  const double* delays;  // [U][C]
  float2* W;             // [C][Gp]
  int U, C, M, K, Gp; float samplerate;
};

cudaError_t launch_analysis(const AnalysisArgs& a, cudaStream_t st);
// analysis + per-bin NLMS in one kernel, snapshots kept on the SM (btkb_fused.cu; BTKB_FUSED=1)
bool fused_supported(const AnalysisArgs& a, const PerBinArgs& b);
cudaError_t launch_fused_analysis_nlms(const AnalysisArgs& a, const PerBinArgs& b, cudaStream_t st);
cudaError_t launch_synthesis(const SynthesisArgs& a, cudaStream_t st);
cudaError_t launch_perbin(const PerBinArgs& a, cudaStream_t st);
cudaError_t launch_covariance(const PerBinArgs& a, cudaStream_t st);
cudaError_t launch_perbin_rls_cpp(const PerBinArgs& a, const double* delays, float samplerate, cudaStream_t st);   // fp64, btkb_rls_cpp.cu
// wide arrays (C = 16, 32, 64): btkb_wide.cu
cudaError_t launch_perbin_wide(const PerBinArgs& a, cudaStream_t st);
cudaError_t launch_covariance_wide(const PerBinArgs& a, cudaStream_t st);
cudaError_t launch_covariance_tc(const PerBinArgs& a, cudaStream_t st);
size_t covariance_tc_workspace_bytes(int G, int C, int T);   // C = 64: tcgen05 / TMEM (btkb_cov_tc.cu)
cudaError_t launch_mvdr_solve_wide(const float2* R, const float2* D, float2* W, const int* noise_count, int U, int C, int K, int Gp, float mu, int normalize_by_count,
                                   int mode /* 0 LU, 1 warp Cholesky + LU, 2 blocked tensor-core Cholesky + LU, 3 implicit-pivoting LU */,
                                   int* list /* [1 + U K] chains left to the LU by the Cholesky modes */, cudaStream_t st);
cudaError_t launch_mainlobe_weights(const WeightsArgs& a, cudaStream_t st);

// setup kernels (btkb_weights.cu)
cudaError_t launch_blocking_wl(const float2* W, const float2* WA, float2* WL, int U, int C, int K, int Gp, int NC, cudaStream_t st);
cudaError_t launch_upgrade_source(const float2* WQ, const float2* WL, float2* OUT, int rows, int U, int K, int Gp, int unit, cudaStream_t st);
cudaError_t launch_lcmv_weights(const double* delaysT, const double* delaysJ, float2* W, int U, int C, int NC, int M, int K, int Gp, float samplerate, cudaStream_t st);
cudaError_t launch_spectral_recursion(const PerBinArgs& a, float mu, int noconj, cudaStream_t st);
cudaError_t launch_diffuse_model(const double* mpos, float2* R, int U, int C, int M, int K, int Gp, float samplerate, float sspeed, cudaStream_t st);
cudaError_t launch_mvdr_solve(const float2* R, const float2* D, float2* W, const int* noise_count, int U, int C, int K, int Gp, float mu, int normalize_by_count, float dthreshold,
                              cudaStream_t st);
cudaError_t launch_adaptive_rebase(const float2* Wold, const float2* Wnew, float2* UA, float* ST, int has_P, int U, int C, int K, int Gp, cudaStream_t st);
cudaError_t launch_ua_to_wa(const float2* UA, const float2* W, float2* WA, int U, int C, int K, int Gp, cudaStream_t st);
cudaError_t launch_noise_mask(const float* E, const int* lengths, const double* labels, unsigned char* mask, int* count, int U, int T, int D, int laN, int pdA,
                              float samplerate, float thr, cudaStream_t st);

// McCowan / Lefkimmiatis post-filter setup (btkb_postfilter.cu); Rpf / invR are [K][C][C] complex128, C <= 8
cudaError_t launch_pf_diffuse(const double* mpos, double2* Rpf, int C, int M, int K, double samplerate, double sspeed, cudaStream_t st);
cudaError_t launch_pf_diag_load(double2* Rpf, int C, int K, float mu, cudaStream_t st);
cudaError_t launch_pf_divide_nondiag(double2* Rpf, int C, int K, float mu, cudaStream_t st);
cudaError_t launch_pf_prepare(const double2* Rpf, double2* invR, float2* PFQ, int C, int K, float threshold, double min_sv, int want_inverse, cudaStream_t st);
cudaError_t launch_pf_lambda(const double2* invR, const float2* TA, float* LAM, int U, int C, int K, int Gp, int pf_type, cudaStream_t st);
inline int pf_num_consts(int C) { return C * (C - 1) + 2 * C; }  // q'[NP], rho'[C], q''[NP], rho''[C]

// Kernel-variant switches BTKB_ANALYSIS_PACKED / BTKB_SYNTHESIS_PACKED / BTKB_PERBIN_PACKED: 1 (default since round 2) = packed 2 x fp32
// arithmetic (FADD2 / FMUL2 / FFMA2, btkb_f2.cuh), 0 = the scalar kernels.  Bit-identical results, verified on B200 kernel by kernel
// (tests/test_zz_host_surface.py::test_packed_fp32_kernel_equals_the_scalar_kernel); read at every launch so one process can compare.
inline bool env_packed(const char* name) { const char* e = getenv(name); return e ? atoi(e) != 0 : true; }

__host__ __device__ inline int frames_of(int len, int D, int laN, int pdA) { return (len + D - 1) / D - laN + pdA; }

}  // namespace btkb
