// btkb_api.cu — the C-ABI (include/btkb.h) over the sm_100a kernels: pipeline object, device buffers, stream ordering.
// No CPU compute path exists here: every data-path entry point launches CUDA kernels or fails.
#include <cuda_runtime.h>
#include <string>
#include <vector>
#include <cstring>
#include <cstdio>
#include <cmath>
#include <algorithm>

#include "../../include/btkb.h"
#include "btkb_internal.h"

using namespace btkb;

static thread_local std::string g_err;
static int fail(int code, const std::string& msg) { g_err = msg; return code; }
#define CK(call)                                                                                      \
  do {                                                                                                \
    cudaError_t e__ = (call);                                                                         \
    if (e__ != cudaSuccess)                                                                           \
      return fail(BTKB_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e__));                \
  } while (0)

struct btkb_pipeline {
  btkb_config cfg;
  int C, Cp, M, K, D, R, m, laN, pdA, pdS;   // Cp: channel rows of the device arrays (C, or C padded with zero channels to 16 / 32 / 64 for a wide array)
  int Ucap, ncap, n_stride, Tcap, Gpcap;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
  // device buffers
  float* d_x = nullptr; const float* x_cur = nullptr;
  int* d_len = nullptr;
  float *d_h = nullptr, *d_g = nullptr;
  float2 *d_X = nullptr, *d_Y = nullptr, *d_W = nullptr, *d_TA = nullptr, *d_WL = nullptr, *d_WA = nullptr, *d_UA = nullptr, *d_R = nullptr;
  float *d_E = nullptr, *d_time = nullptr, *d_upd = nullptr, *d_PFW = nullptr;
  double *d_delays = nullptr, *d_mpos = nullptr, *d_labels = nullptr, *d_stats = nullptr;
  unsigned char* d_mask = nullptr; int* d_count = nullptr; int* d_todo = nullptr;   // d_todo: [1 + Gpcap] chains the Cholesky pass of the wide MVDR solve leaves to the LU (d_todo[0] = count)
  void* d_scratch = nullptr; size_t scratch_bytes = 0;
  int16_t* d_x16 = nullptr; const int16_t* x16_cur = nullptr; int x16_stride = 0; double* h_delays = nullptr; float2* d_tw = nullptr;  // lazily allocated int16 staging; pinned host staging for delays
  double2 *d_pfR = nullptr, *d_pfInvR = nullptr; float2* d_pfQ = nullptr; float* d_LAM = nullptr;  // McCowan / Lefkimmiatis coherence + constants
  // multi-channel WPE (lazily sized at create when cfg.wpe.enabled)
  float2 *d_wS = nullptr, *d_wG = nullptr; void* d_wR = nullptr; float* d_wTH = nullptr; int* d_werr = nullptr;
  int wpe_P = 0, wpe_L = 0, wpe_Lr = 0, wpe_chunk = 0, wpe_chunk_frame = 0, wpe_chol_threads = 0, wpe_Ts = 0, wpe_nbins = 0, wpe_U = 0, wpe_form = -1, wpe_last_form = -1; size_t wpe_slot = 0; bool have_wpe = false;
  cudaEvent_t wev[2] = {nullptr, nullptr};
  // SOS batch beamformers (lazily allocated by the first btkb_sos_accumulate_*)
  double2 *d_sosR = nullptr, *d_sosWd = nullptr; double* d_sosCnt = nullptr; float *d_sosWtu = nullptr, *d_sosMask = nullptr; double* d_sosLab = nullptr;
  int* d_sosErr = nullptr; int sos_NLcap = 0, sos_U = 0; bool have_sos = false;
  float2* d_covS = nullptr;   // series-major workspace of the 64-mic tensor-core covariance (lazily allocated)
  bool have_pfR = false, pf_applied = false;  // pf_applied: d_Y came out of this pipeline's post-filter (not btkb_set_subband)
  // batch state
  int U = 0, n = 0, T = 0, nb = 0, Gp = 0, wU = 0, NC = 1;
  float2 *d_BS = nullptr, *d_BI = nullptr, *d_Z = nullptr;   // upgrade_blocking_matrix / blocking_matrix_output (allocated on first use)
  bool bm_upgraded = false;
  int bm_source = 0;  // BTKB_BF_MVDR: 0 = blocking matrix from the delay-and-sum manifold (calc_blocking_matrix1), 1 = from wmvdr (calc_blocking_matrix2)
  double* d_delaysJ = nullptr;
  std::vector<int> lengths;
  bool have_h = false, have_g = false, have_ta = false, have_w = false, have_wl = false, have_R = false, R_is_sum = false;
  bool have_X = false, have_Y = false, have_time = false, have_ua = false;
  float timing[5] = {0, 0, 0, 0, 0};
  int launches = 0;
  // streamed chunks (btkb_stream_begin / btkb_stream_submit): sample rows with a history prefix (two buffers, the tail of one becomes
  // the head of the other), Y rows preceded by m R - 1 frames of history, the per-chain recurrence state, absolute counters
  bool streaming = false, stream_final = false, owns_stream = true;
  float* d_xs[2] = {nullptr, nullptr}; int xs_stride = 0, xs_cur = 0, xs_hist = 0;
  float* d_ST = nullptr; int* d_tu = nullptr; float2* Y_out = nullptr;   // Y_out: row 0 of the current batch / chunk inside d_Y
  long long s_samples = 0;   // samples per utterance submitted so far (non-final chunks are whole blocks)
  int s_tnext = 0, s_tlast = 0, s_bnext = 0, s_blast = 0, s_chunks = 0, Hy = 0;   // absolute frame / block number of the next and of the last processed chunk's first frame / block
  std::vector<long long> s_len;   // absolute valid length per utterance
  std::vector<int> s_tu;          // absolute frames per utterance so far
};

static int round_up(int a, int b) { return (a + b - 1) / b * b; }

extern "C" {

const char* btkb_last_error(void) { return g_err.c_str(); }

int btkb_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
  return n;
}

void btkb_default_config(btkb_config* c) {
  memset(c, 0, sizeof(*c));
  c->device = 0; c->channels = 8; c->fft_len = 512; c->m = 4; c->r = 1; c->delay_compensation_type = 2;
  c->samplerate = 16000.f; c->beamformer = BTKB_BF_DS; c->postfilter = BTKB_PF_NONE;
  c->pf_alpha = 0.6f; c->pf_type = 2; c->pf_min_frames = 0;
  c->pf_threshold = 0.99f; c->pf_min_sv = 1.0e-8; c->pf_fbin1 = 0;
  c->lms.beta = 0.97f; c->lms.gamma = 0.01f; c->lms.init_diagonal_load = 1.0e6f; c->lms.regularization_param = 1.0e-4f;
  c->lms.energy_floor = 90.f; c->lms.sil_thresh = 1.0e8f; c->lms.max_wa_l2norm = 100.f; c->lms.min_frames = 128; c->lms.slowdown_after = 4096;
  c->max_utterances = 1; c->max_samples = 160000; c->keep_snapshots = 1; c->synthesis_gain = 1;
  c->wpe.enabled = 0; c->wpe.lower_num = 0; c->wpe.upper_num = 32; c->wpe.iterations_num = 2; c->wpe.load_db = -18.0; c->wpe.band_width = 0.0;
  c->wpe.diagonal_bias = 1.0e-4;
  c->rls.beta = 0.97f; c->rls.gamma = 0.04f; c->rls.mu = 0.97f; c->rls.init_diagonal_load = 1.0e6f; c->rls.regularization_param = 1.0e-2f;
  c->rls.sil_thresh = 1.0e8f; c->rls.alpha2 = 10.f; c->rls.max_wa_l2norm = 100.f; c->rls.constraint_option = 3; c->rls.min_frames = 128;
  c->rls_cpp.mu = 0.9f; c->rls_cpp.sigma2 = 0.01f; c->rls_cpp.init_sigma2 = 0.01f; c->rls_cpp.alpha = -1.0f; c->rls_cpp.qctype = 0; c->rls_cpp.update = 1;
}

static void fb_delays(int m, int r, int dct, bool synthesis, int* pd, int* la) {  // modulated.cc:246-264
  const int R = 1 << r;
  *la = 0;
  if (dct == 1) { *pd = m * R - 1; }
  else if (dct == 2) { if (synthesis) *pd = m * R / 2; else { *pd = m * R - 1; *la = m * R / 2 - 1; } }
  else { *pd = 2 * m - 1; }
}

void btkb_destroy(btkb_pipeline* p) {
  if (!p) return;
  cudaSetDevice(p->cfg.device);
  void* ptrs[] = {p->d_BS, p->d_BI, p->d_Z, p->d_xs[0], p->d_xs[1], p->d_ST, p->d_tu, p->d_x, p->d_len, p->d_h, p->d_g, p->d_X, p->d_Y, p->d_W, p->d_TA, p->d_WL, p->d_WA, p->d_UA, p->d_R, p->d_E, p->d_time, p->d_upd,
                  p->d_PFW, p->d_delays, p->d_mpos, p->d_labels, p->d_stats, p->d_mask, p->d_todo, p->d_count, p->d_scratch, p->d_x16, p->d_delaysJ, p->d_tw,
                  p->d_pfR, p->d_pfInvR, p->d_pfQ, p->d_LAM, p->d_wS, p->d_wG, p->d_wR, p->d_wTH, p->d_werr,
                  p->d_sosR, p->d_sosWd, p->d_sosCnt, p->d_sosWtu, p->d_sosMask, p->d_sosLab, p->d_sosErr, p->d_covS};
  if (p->h_delays) cudaFreeHost(p->h_delays);
  for (void* q : ptrs) if (q) cudaFree(q);
  for (auto& e : p->ev) if (e) cudaEventDestroy(e);
  for (auto& e : p->wev) if (e) cudaEventDestroy(e);
  if (p->stream && p->owns_stream) cudaStreamDestroy(p->stream);
  delete p;
}

int btkb_create(const btkb_config* cfg, btkb_pipeline** out) {
  if (!cfg || !out) return fail(BTKB_ERR_INVALID, "btkb_create: null argument");
  *out = nullptr;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    cudaGetLastError();
    return fail(BTKB_ERR_NO_DEVICE, "btkb_create: no CUDA device is visible; this library has no CPU path");
  }
  if (cfg->device < 0 || cfg->device >= ndev) return fail(BTKB_ERR_INVALID, "btkb_create: bad device ordinal");
  const int M = cfg->fft_len, C = cfg->channels;
  if (M < 256 || M > 2048 || (M & (M - 1))) return fail(BTKB_ERR_INVALID, "btkb_create: fft_len must be a power of two in [256, 2048]");
  if (cfg->m < 1 || cfg->m > 8 || cfg->r < 0 || (M >> cfg->r) < 32) return fail(BTKB_ERR_INVALID, "btkb_create: bad m / r");
  if (C < 1) return fail(BTKB_ERR_INVALID, "btkb_create: channels must be >= 1");
  if (C > 64) return fail(BTKB_ERR_INVALID, "btkb_create: at most 64 channels");
  if (cfg->max_utterances < 1 || cfg->max_samples < 1) return fail(BTKB_ERR_INVALID, "btkb_create: capacities must be positive");
  if (cfg->beamformer < BTKB_BF_DS || cfg->beamformer > BTKB_BF_GSC_RLS_CPP) return fail(BTKB_ERR_INVALID, "btkb_create: unknown beamformer kind");
  if (cfg->beamformer == BTKB_BF_GSC_RLS_CPP && (C < 2 || C > 8 || cfg->postfilter != BTKB_PF_NONE))
    return fail(BTKB_ERR_INVALID, "btkb_create: the C++ SubbandGSCRLS kernel is built for 2..8 channels without a fused post-filter");
  if (cfg->postfilter < BTKB_PF_NONE || cfg->postfilter > BTKB_PF_LEFKIMMIATIS) return fail(BTKB_ERR_INVALID, "btkb_create: unknown post-filter kind");
  if (cfg->beamformer == BTKB_BF_MVDR && C > 8 && !(C == 16 || C == 32 || C == 64))
    return fail(BTKB_ERR_INVALID, "btkb_create: MVDR on a wide array is built for 16, 32 or 64 channels (other wide counts run delay-and-sum / GSC / NLMS on zero-padded channel rows)");
  if (cfg->postfilter >= BTKB_PF_MCCOWAN && (C < 2 || C > 8))
    return fail(BTKB_ERR_INVALID, "btkb_create: the McCowan / Lefkimmiatis post-filters are built for 2..8 channels");
  if (cfg->beamformer == BTKB_BF_GSC_RLS && cfg->postfilter != BTKB_PF_NONE)
    return fail(BTKB_ERR_INVALID, "btkb_create: the reference wires no post-filter behind SubbandGSCRLSBeamformer");
  if (cfg->beamformer == BTKB_BF_GSC_RLS && (C < 2 || C > 8))
    return fail(BTKB_ERR_INVALID, "btkb_create: the RLS sidelobe canceller keeps its C x C precision matrix in registers and is built for 2..8 channels");
  if (cfg->beamformer == BTKB_BF_GSC_LMS && cfg->postfilter != BTKB_PF_NONE)
    return fail(BTKB_ERR_INVALID, "btkb_create: the reference wires no post-filter behind SubbandGSCLMSBeamformer");
  if (cfg->wpe.enabled) {
    if (C < 1 || C > 8) return fail(BTKB_ERR_INVALID, "btkb_create: WPE is built for 1..8 channels");
    if (cfg->wpe.lower_num < 0 || cfg->wpe.upper_num < cfg->wpe.lower_num || cfg->wpe.iterations_num < 0)
      return fail(BTKB_ERR_INVALID, "btkb_create: bad WPE lag range / iteration count");
    if (cfg->wpe.band_width > cfg->samplerate / 2.0)  // dereverberation.cc:369-370
      return fail(BTKB_ERR_INVALID, "Bandwidth is greater than the Nyquist rate.");
  }
  CK(cudaSetDevice(cfg->device));
  btkb_pipeline* p = new btkb_pipeline();
  p->cfg = *cfg;
  p->C = C; p->Cp = (C <= 8) ? C : (C <= 16 ? 16 : (C <= 32 ? 32 : 64)); p->M = M; p->K = M / 2 + 1; p->m = cfg->m; p->R = 1 << cfg->r; p->D = M >> cfg->r;
  fb_delays(cfg->m, cfg->r, cfg->delay_compensation_type, false, &p->pdA, &p->laN);
  int la_dummy; fb_delays(cfg->m, cfg->r, cfg->delay_compensation_type, true, &p->pdS, &la_dummy);
  p->Ucap = cfg->max_utterances; p->ncap = cfg->max_samples; p->n_stride = round_up(cfg->max_samples, 4);
  p->Tcap = (p->ncap + p->D - 1) / p->D + p->pdA;   // >= frames_of(ncap): the final chunk of a stream emits its blocks plus all pd_A flush frames
  p->Gpcap = round_up(p->Ucap * p->K, 128);
  const size_t G = (size_t)p->Gpcap, T = (size_t)p->Tcap, U = (size_t)p->Ucap;
  cudaError_t e = cudaSuccess;
  auto A = [&](void** ptr, size_t bytes) { if (e == cudaSuccess) e = cudaMalloc(ptr, bytes ? bytes : 16); };
  e = cudaStreamCreateWithFlags(&p->stream, cudaStreamNonBlocking);
  for (auto& ev : p->ev) if (e == cudaSuccess) e = cudaEventCreate(&ev);
  A((void**)&p->d_x, U * C * p->n_stride * sizeof(float));
  A((void**)&p->d_len, U * sizeof(int));
  A((void**)&p->d_h, (size_t)p->m * M * sizeof(float));
  A((void**)&p->d_g, (size_t)p->m * M * sizeof(float));
  const size_t Cp = (size_t)p->Cp;
  A((void**)&p->d_X, T * Cp * G * sizeof(float2));
  p->Hy = p->m * p->R + 1;   // rows of Y history in front of a streamed chunk: a block reaches back m R - 1 frames, +1 for the block held back, +1 for the pair-aligned tile origin
  A((void**)&p->d_Y, (T + p->Hy) * G * sizeof(float2));
  A((void**)&p->d_W, Cp * G * sizeof(float2));
  A((void**)&p->d_TA, Cp * G * sizeof(float2));
  A((void**)&p->d_WL, Cp * G * sizeof(float2));
  A((void**)&p->d_WA, Cp * G * sizeof(float2));
  A((void**)&p->d_UA, Cp * G * sizeof(float2));
  if (cfg->beamformer == BTKB_BF_MVDR) A((void**)&p->d_R, (size_t)C * C * G * sizeof(float2));
  A((void**)&p->d_E, T * U * sizeof(float));
  A((void**)&p->d_time, U * (T * p->D) * sizeof(float));
  A((void**)&p->d_upd, U * sizeof(float));
  if (cfg->postfilter != BTKB_PF_NONE) A((void**)&p->d_PFW, T * G * sizeof(float));
  if (cfg->postfilter >= BTKB_PF_MCCOWAN) {
    A((void**)&p->d_pfR, (size_t)p->K * C * C * sizeof(double2));
    A((void**)&p->d_pfInvR, (size_t)p->K * C * C * sizeof(double2));
    A((void**)&p->d_pfQ, (size_t)pf_num_consts(C) * p->K * sizeof(float2));
    A((void**)&p->d_LAM, G * sizeof(float));
  }
  A((void**)&p->d_delays, U * C * sizeof(double));
  A((void**)&p->d_mpos, (size_t)C * 3 * sizeof(double));
  A((void**)&p->d_labels, U * 2 * sizeof(double));
  A((void**)&p->d_stats, U * 3 * sizeof(double));
  A((void**)&p->d_mask, T * U);
  A((void**)&p->d_count, U * sizeof(int));
  A((void**)&p->d_tw, (size_t)M * sizeof(float2));
  if (cfg->wpe.enabled) {
    p->wpe_P = cfg->wpe.upper_num - cfg->wpe.lower_num + 1; p->wpe_L = C * p->wpe_P;
    p->wpe_Ts = round_up(p->Tcap, 2);
    // Which normal equations (btkb_wpe.cu): -1 = per batch, the smaller of the lag-domain (L x L) and the frame-domain (S x S) system.
    // BTKB_WPE_FORM=lag|frame pins one for the life of the pipeline; pinning "frame" sizes the matrix slots for S up to Tcap.
    const char* fe = getenv("BTKB_WPE_FORM");
    p->wpe_form = (fe && !strcmp(fe, "lag")) ? 0 : (fe && !strcmp(fe, "frame")) ? 1 : -1;
    const int nslot = (p->wpe_form == 1) ? std::max(p->wpe_L, std::max(p->Tcap - cfg->wpe.lower_num, 0)) : p->wpe_L;
    p->wpe_Lr = round_up(nslot + 1, 2);  // column L exists: the augmented row carries its own diagonal entry
    p->wpe_slot = (size_t)(nslot + 1) * p->wpe_Lr;
    const int lo = (cfg->wpe.band_width == 0.0) ? M / 2 : (int)((cfg->wpe.band_width / (cfg->samplerate / 2.0)) * (M / 2));  // set_band_width_ (:365-373)
    p->wpe_nbins = std::min(lo, p->K - 1) + 1;
    const char* ce = getenv("BTKB_WPE_CHUNK");
    p->wpe_chunk = ce ? std::max(1, atoi(ce)) : 55;   // 55 x 8 = 440 Cholesky CTAs fit one wave at 3 CTAs/SM x 148 SMs (56 spills 4 CTAs into a second wave: 663 -> 862 ms measured)
    p->wpe_chunk = std::min(p->wpe_chunk, p->Ucap * p->wpe_nbins);
    // frame-domain form: 128-thread CTAs, 4 per SM -> 592 systems in flight; eight such waves per launch (77.9 ms per 8 utterances of configs[4] against 79.9 with four and 83.8 with two)
    const char* cf = getenv("BTKB_WPE_CHUNK_FRAME");
    const size_t esz = cfg->wpe.fp32_normal_equations ? sizeof(float2) : sizeof(double2);
    const int fit = (int)std::max<size_t>(1, ((size_t)6 << 30) / ((size_t)(C + 1) * p->wpe_slot * esz));   // workspace <= 6 GiB
    p->wpe_chunk_frame = std::min(std::min(cf ? std::max(1, atoi(cf)) : std::max(1, 4736 / C), fit), p->Ucap * p->wpe_nbins);
    const char* ct = getenv("BTKB_WPE_CHOL_THREADS");
    p->wpe_chol_threads = ct ? std::min(256, std::max(32, atoi(ct) / 32 * 32)) : 0;
    for (auto& ev : p->wev) if (e == cudaSuccess) e = cudaEventCreate(&ev);
    A((void**)&p->d_wS, (size_t)p->Ucap * p->K * C * p->wpe_Ts * sizeof(float2));
    A((void**)&p->d_wTH, (size_t)p->Ucap * p->K * C * p->wpe_Ts * sizeof(float));
    A((void**)&p->d_wG, (size_t)p->Ucap * p->K * C * p->wpe_L * sizeof(float2));
    A((void**)&p->d_wR, wpe_workspace_bytes(C, p->wpe_slot, std::max(p->wpe_chunk, p->wpe_chunk_frame), cfg->wpe.fp32_normal_equations));
    A((void**)&p->d_werr, sizeof(int));
  }
  if (e == cudaSuccess && p->Cp != C) {   // the padded channel rows stay zero for the life of the pipeline: no kernel writes them
    e = cudaMemset(p->d_X, 0, T * Cp * G * sizeof(float2));
    for (float2* q : {p->d_W, p->d_TA, p->d_WL, p->d_WA, p->d_UA}) if (e == cudaSuccess) e = cudaMemset(q, 0, Cp * G * sizeof(float2));
  }
  if (e != cudaSuccess) {
    std::string msg = std::string("btkb_create: allocation failed: ") + cudaGetErrorString(e);
    btkb_destroy(p);
    cudaGetLastError();
    return fail(BTKB_ERR_ALLOC, msg);
  }
  {  // twiddle table exp(+2 pi i n / M) in double precision, rounded once
    std::vector<float2> tw(M);
    for (int i = 0; i < M; i++) { const double a = 2.0 * M_PI * (double)i / (double)M; tw[i] = make_float2((float)cos(a), (float)sin(a)); }
    if (cudaMemcpy(p->d_tw, tw.data(), tw.size() * sizeof(float2), cudaMemcpyHostToDevice) != cudaSuccess) {
      btkb_destroy(p); cudaGetLastError();
      return fail(BTKB_ERR_CUDA, "btkb_create: twiddle upload failed");
    }
  }
  p->Y_out = p->d_Y;
  *out = p;
  return BTKB_OK;
}

int btkb_set_prototypes(btkb_pipeline* p, const double* h, const double* g, int len) {
  if (!p || !h) return fail(BTKB_ERR_INVALID, "btkb_set_prototypes: null argument");
  if (len != p->m * p->M)  // modulated.cc:239-241 jconsistency_error
    return fail(BTKB_ERR_INVALID, "Prototype sizes do not match (" + std::to_string(len) + " vs. " + std::to_string(p->m * p->M) + ").");
  CK(cudaSetDevice(p->cfg.device));
  std::vector<float> hf(len), gf(len);
  for (int i = 0; i < len; i++) hf[i] = (float)h[i];
  CK(cudaMemcpyAsync(p->d_h, hf.data(), len * sizeof(float), cudaMemcpyHostToDevice, p->stream));
  p->have_h = true;
  if (g) {
    for (int i = 0; i < len; i++) gf[i] = (float)g[i];
    CK(cudaMemcpyAsync(p->d_g, gf.data(), len * sizeof(float), cudaMemcpyHostToDevice, p->stream));
    p->have_g = true;
  }
  CK(cudaStreamSynchronize(p->stream));
  return BTKB_OK;
}

// host [U][K][X] complex64  <->  device [X][Gp]
static void to_device_layout(const float* src, std::vector<float2>& dst, int U, int K, int X, int Gp) {
  dst.assign((size_t)X * Gp, make_float2(0.f, 0.f));
  for (int u = 0; u < U; u++)
    for (int k = 0; k < K; k++)
      for (int c = 0; c < X; c++) {
        const float* s = src + 2 * (((size_t)u * K + k) * X + c);
        dst[(size_t)c * Gp + (size_t)u * K + k] = make_float2(s[0], s[1]);
      }
}
static void from_device_layout(const std::vector<float2>& src, float* dst, int U, int K, int X, int Gp) {
  for (int u = 0; u < U; u++)
    for (int k = 0; k < K; k++)
      for (int c = 0; c < X; c++) {
        float2 v = src[(size_t)c * Gp + (size_t)u * K + k];
        float* d = dst + 2 * (((size_t)u * K + k) * X + c);
        d[0] = v.x; d[1] = v.y;
      }
}

// Weight setters and submissions may arrive in either order and with a different utterance count from one batch to the next (a
// pipeline sized for 256 utterances must also take the trailing partial batch).  Rules:
//   * check_weight_batch only validates; begin_weight_batch commits, and is called after the setter's own argument checks, so a
//     failed call leaves the pipeline untouched;
//   * a setter whose U differs from the U the resident weights were made for starts a NEW weight set: everything derived for the
//     old count (manifold, quiescent / active weights, covariance) is void;
//   * the device layouts of X / Y / weights share the row pitch Gp = roundup(U K, 128), so a resident batch of another size cannot
//     survive a weight setter for a new U (it is dropped: have_X = have_Y = false), and the beamformer refuses to run while the
//     weights and the submitted batch disagree (do_beamformer).
static int check_weight_batch(btkb_pipeline* p, int U, const char* who) {
  if (U < 1 || U > p->Ucap) return fail(BTKB_ERR_INVALID, std::string(who) + ": U out of range");
  return BTKB_OK;
}
static void begin_weight_batch(btkb_pipeline* p, int U) {
  if (p->wU != U) { p->have_w = p->have_wl = p->have_ta = p->have_R = false; p->R_is_sum = false; p->NC = 1; }
  if (p->U != 0 && p->U != U) { p->have_X = p->have_Y = p->have_time = p->have_ua = false; p->U = 0; p->T = 0; p->nb = 0; }
  p->wU = U;
  p->Gp = round_up(U * p->K, 128);
}

int btkb_set_delays(btkb_pipeline* p, int U, const double* delays) {
  if (!p || !delays) return fail(BTKB_ERR_INVALID, "btkb_set_delays: null argument");
  int rc = check_weight_batch(p, U, "btkb_set_delays"); if (rc) return rc;
  if ((p->cfg.beamformer == BTKB_BF_GSC || p->cfg.beamformer == BTKB_BF_GSC_LMS || p->cfg.beamformer == BTKB_BF_GSC_RLS || p->cfg.beamformer == BTKB_BF_GSC_RLS_CPP) && p->C <= 1)  // beamformer.cc:507-510
    return fail(BTKB_ERR_INVALID, "The number of channels must be > 1 but it is " + std::to_string(p->C));
  CK(cudaSetDevice(p->cfg.device));
  begin_weight_batch(p, U);
  // stage through pinned host memory owned by the pipeline: the call stays asynchronous (no stream sync) and `delays`
  // may be reused by the caller immediately
  if (!p->h_delays) CK(cudaMallocHost((void**)&p->h_delays, (size_t)p->Ucap * p->C * sizeof(double)));
  CK(cudaEventSynchronize(p->ev[4]));  // the previous delays upload (if any) has left the staging buffer
  memcpy(p->h_delays, delays, (size_t)U * p->C * sizeof(double));
  CK(cudaMemcpyAsync(p->d_delays, p->h_delays, (size_t)U * p->C * sizeof(double), cudaMemcpyHostToDevice, p->stream));
  CK(cudaEventRecord(p->ev[4], p->stream));
  // a new look direction in the middle of a stream: the adaptive state is kept, re-expressed for the new blocking matrices
  const bool adaptive = p->cfg.beamformer == BTKB_BF_GSC_LMS || p->cfg.beamformer == BTKB_BF_GSC_RLS;
  const bool rebase = adaptive && p->streaming && p->s_chunks > 0 && p->have_ta && p->C >= 2;
  if (rebase) CK(cudaMemcpyAsync(p->d_WL, p->d_TA, (size_t)p->C * p->Gp * sizeof(float2), cudaMemcpyDeviceToDevice, p->stream));   // WL is free in the adaptive modes
  WeightsArgs a{p->d_delays, p->d_TA, U, p->C, p->M, p->K, p->Gp, p->cfg.samplerate};
  CK(launch_mainlobe_weights(a, p->stream));
  p->have_ta = true; p->NC = 1; p->bm_upgraded = false;   // new BeamformerWeights (alloc_bfweight_): an upgraded blocking matrix is gone
  if (p->cfg.beamformer != BTKB_BF_MVDR) {
    CK(cudaMemcpyAsync(p->d_W, p->d_TA, (size_t)p->C * p->Gp * sizeof(float2), cudaMemcpyDeviceToDevice, p->stream));
    p->have_w = true;
  }
  if (rebase) CK(launch_adaptive_rebase(p->d_WL, p->d_W, p->d_UA, p->d_ST, p->cfg.beamformer == BTKB_BF_GSC_RLS ? 1 : 0, U, p->C, p->K, p->Gp, p->stream));
  return BTKB_OK;
}

int btkb_set_delays_lcmv(btkb_pipeline* p, int U, int NC, const double* delaysT, const double* delaysJ) {
  if (!p || !delaysT || !delaysJ) return fail(BTKB_ERR_INVALID, "btkb_set_delays_lcmv: null argument");
  if (NC < 2 || NC > 4 || NC > p->C)  // beamformer.cc:592-594
    return fail(BTKB_ERR_INVALID, "1 < the number of constraints " + std::to_string(NC) + " <= the number of sensors " + std::to_string(p->C) + " (and <= 4 in this build).");
  if (p->cfg.beamformer == BTKB_BF_GSC_LMS || p->cfg.beamformer == BTKB_BF_GSC_RLS)
    return fail(BTKB_ERR_INVALID, "btkb_set_delays_lcmv: the adaptive (NLMS / RLS) kernels implement one constraint");
  int rc = btkb_set_delays(p, U, delaysT);  // calcMainlobe first (also the time-alignment manifold), beamformer.cc:617
  if (rc) return rc;
  CK(cudaStreamSynchronize(p->stream));
  if (!p->d_delaysJ) CK(cudaMalloc((void**)&p->d_delaysJ, (size_t)p->Ucap * 3 * p->C * sizeof(double)));
  CK(cudaMemcpy(p->d_delaysJ, delaysJ, (size_t)U * (NC - 1) * p->C * sizeof(double), cudaMemcpyHostToDevice));
  CK(launch_lcmv_weights(p->d_delays, p->d_delaysJ, p->d_W, U, p->C, NC, p->M, p->K, p->Gp, p->cfg.samplerate, p->stream));
  p->NC = NC; p->have_w = true; p->have_wl = false;
  return BTKB_OK;
}

int btkb_set_weights(btkb_pipeline* p, int U, const float* w) {
  if (!p || !w) return fail(BTKB_ERR_INVALID, "btkb_set_weights: null argument");
  int rc = check_weight_batch(p, U, "btkb_set_weights"); if (rc) return rc;
  CK(cudaSetDevice(p->cfg.device));
  begin_weight_batch(p, U);
  std::vector<float2> tmp;
  to_device_layout(w, tmp, U, p->K, p->C, p->Gp);
  CK(cudaMemcpyAsync(p->d_W, tmp.data(), tmp.size() * sizeof(float2), cudaMemcpyHostToDevice, p->stream));
  CK(cudaStreamSynchronize(p->stream));
  p->have_w = true;
  if (!p->have_ta) {  // setQuiescentVector without delays: the manifold is the weight itself
    CK(cudaMemcpy(p->d_TA, p->d_W, (size_t)p->C * p->Gp * sizeof(float2), cudaMemcpyDeviceToDevice));
    p->have_ta = true;
  }
  return BTKB_OK;
}

static PerBinArgs perbin_args(btkb_pipeline* p);
// the vector the blocking matrix is orthogonal to: wq (the delay-and-sum manifold or, after calc_blocking_matrix2, wmvdr) or,
// after upgrade_blocking_matrix, the stored wq - wl
static const float2* blocking_source(const btkb_pipeline* p) {
  if (p->bm_upgraded) return p->d_BS;
  return (p->cfg.beamformer == BTKB_BF_MVDR && p->bm_source == 0) ? p->d_TA : p->d_W;
}

int btkb_set_active_weights(btkb_pipeline* p, int U, const float* wa) {
  if (!p || !wa) return fail(BTKB_ERR_INVALID, "btkb_set_active_weights: null argument");
  if (!p->have_ta) return fail(BTKB_ERR_STATE, "call calc_gsc_weights_x() once");  // beamformer.cc:1369-1371
  if (p->C < 2) return fail(BTKB_ERR_INVALID, "btkb_set_active_weights: needs at least two channels");
  if (p->C > 8) return fail(BTKB_ERR_INVALID, "btkb_set_active_weights: the blocking-matrix kernels are built for <= 8 channels");
  int rc = check_weight_batch(p, U, "btkb_set_active_weights"); if (rc) return rc;
  if (U != p->wU) return fail(BTKB_ERR_INVALID, "btkb_set_active_weights: the quiescent weights were set for " + std::to_string(p->wU) + " utterances, not " + std::to_string(U));
  // calc_blocking_matrix_(wq_[f], NC, B_[f]) (beamformer.cc:554-562, 693-700): B is built from the quiescent vector
  const float2* bsrc = blocking_source(p);
  if (bsrc == p->d_W && !p->have_w) return fail(BTKB_ERR_STATE, "call calc_mvdr_weights() once");  // calc_blocking_matrix2 returns false without wmvdr (beamformer.cc:2651-2653)
  CK(cudaSetDevice(p->cfg.device));
  std::vector<float2> tmp;
  to_device_layout(wa, tmp, U, p->K, p->C - p->NC, p->Gp);
  CK(cudaMemcpyAsync(p->d_WA, tmp.data(), tmp.size() * sizeof(float2), cudaMemcpyHostToDevice, p->stream));
  CK(launch_blocking_wl(bsrc, p->d_WA, p->d_WL, U, p->C, p->K, p->Gp, p->NC, p->stream));
  CK(cudaStreamSynchronize(p->stream));
  p->have_wl = true;
  return BTKB_OK;
}

int btkb_set_blocking_source(btkb_pipeline* p, int from_mvdr_weights) {
  if (!p) return fail(BTKB_ERR_INVALID, "btkb_set_blocking_source: null pipeline");
  if (p->cfg.beamformer != BTKB_BF_MVDR) return fail(BTKB_ERR_INVALID, "btkb_set_blocking_source: only a BTKB_BF_MVDR pipeline has two candidate quiescent vectors");
  p->bm_source = from_mvdr_weights ? 1 : 0;
  p->bm_upgraded = false;
  p->have_wl = false;  // alloc_bfweight_(1, 1): the active weights set so far are gone (beamformer.cc:2640, 2655)
  return BTKB_OK;
}

int btkb_upgrade_blocking_matrix(btkb_pipeline* p) {
  if (!p) return fail(BTKB_ERR_INVALID, "btkb_upgrade_blocking_matrix: null pipeline");
  if (p->cfg.beamformer != BTKB_BF_MVDR) return fail(BTKB_ERR_INVALID, "btkb_upgrade_blocking_matrix: SubbandMVDRGSC only (BTKB_BF_MVDR)");
  if (!p->have_ta) return fail(BTKB_ERR_STATE, "call calc_array_manifold_vectorsX() once");
  if (p->C < 2 || p->C > 8) return fail(BTKB_ERR_INVALID, "btkb_upgrade_blocking_matrix: the blocking-matrix kernels are built for 2..8 channels");
  const float2* wq = (p->bm_source == 0) ? p->d_TA : p->d_W;   // wq_f: what calc_blocking_matrix1 / 2 stored (beamformer.cc:2638-2672)
  if (wq == p->d_W && !p->have_w) return fail(BTKB_ERR_STATE, "call calc_mvdr_weights() once");
  CK(cudaSetDevice(p->cfg.device));
  if (!p->d_BS) CK(cudaMalloc((void**)&p->d_BS, (size_t)p->Cp * p->Gp * sizeof(float2)));
  CK(launch_upgrade_source(wq, p->have_wl ? p->d_WL : nullptr, p->d_BS, p->C, p->wU, p->K, p->Gp, -1, p->stream));
  CK(cudaStreamSynchronize(p->stream));
  p->bm_upgraded = true;   // wl keeps its value until the next set_active_weights, as in the reference
  return BTKB_OK;
}

int btkb_blocking_matrix_output(btkb_pipeline* p, int outChanX, float* out) {
  if (!p || !out) return fail(BTKB_ERR_INVALID, "btkb_blocking_matrix_output: null argument");
  if (p->cfg.beamformer != BTKB_BF_MVDR) return fail(BTKB_ERR_INVALID, "btkb_blocking_matrix_output: SubbandMVDRGSC only (BTKB_BF_MVDR)");
  if (!p->have_X) return fail(BTKB_ERR_STATE, "btkb_blocking_matrix_output: run the analysis first");
  if (!p->have_ta) return fail(BTKB_ERR_STATE, "call calc_array_manifold_vectorsX() once");
  if (p->C < 2 || p->C > 8) return fail(BTKB_ERR_INVALID, "btkb_blocking_matrix_output: the blocking-matrix kernels are built for 2..8 channels");
  if (outChanX < 0 || outChanX >= p->C - p->NC) return fail(BTKB_ERR_INVALID, "btkb_blocking_matrix_output: outChanX must be in [0, C - NC)");
  if (p->wU != p->U) return fail(BTKB_ERR_INVALID, "btkb_blocking_matrix_output: weights were set for a different number of utterances");
  const float2* bsrc = blocking_source(p);
  if (bsrc == p->d_W && !p->have_w) return fail(BTKB_ERR_STATE, "call calc_mvdr_weights() once");
  CK(cudaSetDevice(p->cfg.device));
  const size_t G = (size_t)p->Gp;
  if (!p->d_BI) CK(cudaMalloc((void**)&p->d_BI, (size_t)2 * p->Cp * G * sizeof(float2)));   // [C] b_i, then [C] the unit active weights
  if (!p->d_Z) CK(cudaMalloc((void**)&p->d_Z, (size_t)p->Tcap * G * sizeof(float2)));
  float2* unit = p->d_BI + (size_t)p->Cp * G;
  CK(launch_upgrade_source(nullptr, nullptr, unit, p->C - p->NC, p->U, p->K, p->Gp, outChanX, p->stream));
  CK(launch_blocking_wl(bsrc, unit, p->d_BI, p->U, p->C, p->K, p->Gp, p->NC, p->stream));   // column outChanX of B
  PerBinArgs a = perbin_args(p);   // b_i^H x for every frame = the delay-and-sum kernel with b_i as its weights
  a.kind = BTKB_BF_DS; a.W = p->d_BI; a.WL = nullptr; a.TA = nullptr; a.Y = p->d_Z; a.PFW = nullptr; a.pf_kind = BTKB_PF_NONE; a.normalize_weight = 0;
  a.ST = nullptr; a.st_load = 0;
  CK(launch_perbin(a, p->stream));
  p->launches += 3;
  std::vector<float2> tmp((size_t)p->T * G);
  CK(cudaMemcpyAsync(tmp.data(), p->d_Z, tmp.size() * sizeof(float2), cudaMemcpyDeviceToHost, p->stream));
  CK(cudaStreamSynchronize(p->stream));
  float2* o = reinterpret_cast<float2*>(out);
  for (int u = 0; u < p->U; u++)
    for (int t = 0; t < p->T; t++)
      memcpy(o + ((size_t)u * p->T + t) * p->K, tmp.data() + (size_t)t * G + (size_t)u * p->K, (size_t)p->K * sizeof(float2));
  return BTKB_OK;
}

int btkb_set_noise_covariance(btkb_pipeline* p, int U, const float* R) {
  if (!p || !R) return fail(BTKB_ERR_INVALID, "btkb_set_noise_covariance: null argument");
  if (!p->d_R) return fail(BTKB_ERR_STATE, "btkb_set_noise_covariance: pipeline was not created with BTKB_BF_MVDR");
  int rc = check_weight_batch(p, U, "btkb_set_noise_covariance"); if (rc) return rc;
  CK(cudaSetDevice(p->cfg.device));
  begin_weight_batch(p, U);
  std::vector<float2> tmp;
  to_device_layout(R, tmp, U, p->K, p->C * p->C, p->Gp);
  CK(cudaMemcpyAsync(p->d_R, tmp.data(), tmp.size() * sizeof(float2), cudaMemcpyHostToDevice, p->stream));
  CK(cudaStreamSynchronize(p->stream));
  p->have_R = true; p->R_is_sum = false;
  return BTKB_OK;
}

int btkb_set_diffuse_noise_model(btkb_pipeline* p, int U, const double* mpos, float sspeed) {
  if (!p || !mpos) return fail(BTKB_ERR_INVALID, "btkb_set_diffuse_noise_model: null argument");
  if (!p->d_R) return fail(BTKB_ERR_STATE, "btkb_set_diffuse_noise_model: pipeline was not created with BTKB_BF_MVDR");
  int rc = check_weight_batch(p, U, "btkb_set_diffuse_noise_model"); if (rc) return rc;
  CK(cudaSetDevice(p->cfg.device));
  begin_weight_batch(p, U);
  CK(cudaMemcpyAsync(p->d_mpos, mpos, (size_t)p->C * 3 * sizeof(double), cudaMemcpyHostToDevice, p->stream));
  CK(launch_diffuse_model(p->d_mpos, p->d_R, U, p->C, p->M, p->K, p->Gp, p->cfg.samplerate, sspeed, p->stream));
  CK(cudaStreamSynchronize(p->stream));
  p->have_R = true; p->R_is_sum = false;
  return BTKB_OK;
}

static int pf_check(btkb_pipeline* p, const char* who) {
  if (!p) return fail(BTKB_ERR_INVALID, std::string(who) + ": null pipeline");
  if (!p->d_pfR) return fail(BTKB_ERR_STATE, std::string(who) + ": pipeline was not created with BTKB_PF_MCCOWAN or BTKB_PF_LEFKIMMIATIS");
  return BTKB_OK;
}

int btkb_pf_set_diffuse_noise_model(btkb_pipeline* p, const double* mpos, double samplerate, double sspeed) {
  int rc = pf_check(p, "btkb_pf_set_diffuse_noise_model"); if (rc) return rc;
  if (!mpos) return fail(BTKB_ERR_INVALID, "btkb_pf_set_diffuse_noise_model: null argument");
  CK(cudaSetDevice(p->cfg.device));
  CK(cudaMemcpyAsync(p->d_mpos, mpos, (size_t)p->C * 3 * sizeof(double), cudaMemcpyHostToDevice, p->stream));
  CK(launch_pf_diffuse(p->d_mpos, p->d_pfR, p->C, p->M, p->K, samplerate, sspeed, p->stream));
  CK(cudaStreamSynchronize(p->stream));
  p->have_pfR = true;
  return BTKB_OK;
}

int btkb_pf_set_noise_coherence(btkb_pipeline* p, const double* R) {
  int rc = pf_check(p, "btkb_pf_set_noise_coherence"); if (rc) return rc;
  if (!R) return fail(BTKB_ERR_INVALID, "btkb_pf_set_noise_coherence: null argument");
  CK(cudaSetDevice(p->cfg.device));
  CK(cudaMemcpyAsync(p->d_pfR, R, (size_t)p->K * p->C * p->C * sizeof(double2), cudaMemcpyHostToDevice, p->stream));
  CK(cudaStreamSynchronize(p->stream));
  p->have_pfR = true;
  return BTKB_OK;
}

int btkb_pf_get_noise_coherence(btkb_pipeline* p, double* R) {
  int rc = pf_check(p, "btkb_pf_get_noise_coherence"); if (rc) return rc;
  if (!R) return fail(BTKB_ERR_INVALID, "btkb_pf_get_noise_coherence: null argument");
  if (!p->have_pfR) return fail(BTKB_ERR_STATE, "Construct/set first a noise coherence matrix");
  CK(cudaSetDevice(p->cfg.device));
  CK(cudaMemcpyAsync(R, p->d_pfR, (size_t)p->K * p->C * p->C * sizeof(double2), cudaMemcpyDeviceToHost, p->stream));
  CK(cudaStreamSynchronize(p->stream));
  return BTKB_OK;
}

int btkb_pf_set_diagonal_loading(btkb_pipeline* p, float mu) {
  int rc = pf_check(p, "btkb_pf_set_diagonal_loading"); if (rc) return rc;
  if (!p->have_pfR) return fail(BTKB_ERR_STATE, "Construct/set first a noise coherence matrix");  // postfilter.cc:631-633
  CK(cudaSetDevice(p->cfg.device));
  CK(launch_pf_diag_load(p->d_pfR, p->C, p->K, mu, p->stream));
  return BTKB_OK;
}

int btkb_pf_divide_nondiagonal(btkb_pipeline* p, float mu) {
  int rc = pf_check(p, "btkb_pf_divide_nondiagonal"); if (rc) return rc;
  if (!p->have_pfR) return fail(BTKB_ERR_STATE, "Construct/set first a noise coherence matrix");
  CK(cudaSetDevice(p->cfg.device));
  CK(launch_pf_divide_nondiag(p->d_pfR, p->C, p->K, mu, p->stream));
  return BTKB_OK;
}

int btkb_calc_mvdr_weights(btkb_pipeline* p, float mu) { return btkb_calc_mvdr_weights_ex(p, mu, 1.0e-8f); }   // dthreshold default of beamformer.i:414-486

int btkb_calc_mvdr_weights_ex(btkb_pipeline* p, float mu, float dthreshold) {
  if (!p) return fail(BTKB_ERR_INVALID, "btkb_calc_mvdr_weights: null argument");
  if (!p->have_R) return fail(BTKB_ERR_STATE, "Set a spatial spectral matrix before calling calc_mvdr_weights()");  // beamformer.cc:2352-2354
  if (!p->have_ta) return fail(BTKB_ERR_STATE, "call calc_array_manifold_vectorsX() once");                      // beamformer.cc:2355-2357
  CK(cudaSetDevice(p->cfg.device));
  if (p->C <= 8) CK(launch_mvdr_solve(p->d_R, p->d_TA, p->d_W, p->d_count, p->wU, p->C, p->K, p->Gp, mu, p->R_is_sum ? 1 : 0, dthreshold, p->stream));
  else {
    if (!p->d_todo) CK(cudaMalloc((void**)&p->d_todo, ((size_t)p->Gpcap + 1) * sizeof(int)));
    // Which solver (btkb_wide.cu: launch_mvdr_solve_wide).  The blocked tensor-core Cholesky is 2.3 x faster than the pivoted LU on Hermitian
    // positive-definite matrices and leaves the others to the LU, so a batch it cannot factor pays for both.  A covariance accumulated on the
    // device from fewer frames than channels is rank deficient (only the loading keeps it definite, and not in fp32 storage): such a batch
    // goes straight to the LU.  BTKB_SOLVE_CHOL=0/1/2 and BTKB_SOLVE_IP=1 pin a solver.
    int mode = 2;
    if (p->R_is_sum) {
      std::vector<int> cnt((size_t)p->wU);
      CK(cudaMemcpyAsync(cnt.data(), p->d_count, cnt.size() * sizeof(int), cudaMemcpyDeviceToHost, p->stream));
      CK(cudaStreamSynchronize(p->stream));
      for (int c : cnt) if (c < 2 * p->C) mode = 0;
    }
    if (const char* ev = getenv("BTKB_SOLVE_CHOL")) mode = std::min(2, std::max(0, atoi(ev)));
    if (const char* ev = getenv("BTKB_SOLVE_IP")) { if (atoi(ev) != 0) mode = 3; }
    CK(launch_mvdr_solve_wide(p->d_R, p->d_TA, p->d_W, p->d_count, p->wU, p->C, p->K, p->Gp, mu, p->R_is_sum ? 1 : 0, mode, p->d_todo, p->stream));
  }
  p->have_w = true;
  return BTKB_OK;
}

static int submit_common(btkb_pipeline* p, int U, int n, const int* lengths) {
  if (U < 1 || U > p->Ucap) return fail(BTKB_ERR_INVALID, "btkb_submit: U exceeds max_utterances");
  if (n < 1 || n > p->ncap) return fail(BTKB_ERR_INVALID, "btkb_submit: n exceeds max_samples");
  // (weights made for another utterance count stay where they are; do_beamformer refuses to combine them with this batch, and the next
  // weight setter for this U replaces them — begin_weight_batch)
  p->U = U; p->n = n;
  p->streaming = false; p->Y_out = p->d_Y; p->x16_cur = nullptr;
  p->lengths.assign(U, n);
  int Tmax = 0;
  for (int u = 0; u < U; u++) {
    if (lengths) { if (lengths[u] < 0 || lengths[u] > n) return fail(BTKB_ERR_INVALID, "btkb_submit: lengths[u] out of range"); p->lengths[u] = lengths[u]; }
    Tmax = std::max(Tmax, frames_of(p->lengths[u], p->D, p->laN, p->pdA));
  }
  p->T = Tmax; p->nb = std::max(Tmax - p->pdS, 0);
  if (p->wU != U) { p->have_w = p->have_wl = p->have_ta = p->have_R = false; p->R_is_sum = false; p->NC = 1; p->wU = 0; }   // their row pitch is another Gp
  p->Gp = round_up(U * p->K, 128);
  p->have_X = p->have_Y = p->have_time = p->have_ua = false;
  CK(cudaMemcpyAsync(p->d_len, p->lengths.data(), U * sizeof(int), cudaMemcpyHostToDevice, p->stream));
  return BTKB_OK;
}

int btkb_submit(btkb_pipeline* p, const float* samples, int U, int n, const int* lengths) {
  if (!p || !samples) return fail(BTKB_ERR_INVALID, "btkb_submit: null argument");
  CK(cudaSetDevice(p->cfg.device));
  int rc = submit_common(p, U, n, lengths); if (rc) return rc;
  // [U][C][n] host -> [U][C][n_stride] device
  CK(cudaMemcpy2DAsync(p->d_x, (size_t)p->n_stride * sizeof(float), samples, (size_t)n * sizeof(float), (size_t)n * sizeof(float), (size_t)U * p->C,
                       cudaMemcpyHostToDevice, p->stream));
  p->x_cur = p->d_x;
  return BTKB_OK;
}

__global__ void k_i16_to_f32(const int16_t* __restrict__ src, float* __restrict__ dst, size_t rows, int n, int n_stride) {
  const size_t total = rows * (size_t)n;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const size_t r = i / n; const int c = (int)(i - r * n);
    dst[r * n_stride + c] = (float)src[i];
  }
}

int btkb_submit_i16(btkb_pipeline* p, const int16_t* samples, int U, int n, const int* lengths) {
  if (!p || !samples) return fail(BTKB_ERR_INVALID, "btkb_submit_i16: null argument");
  CK(cudaSetDevice(p->cfg.device));
  int rc = submit_common(p, U, n, lengths); if (rc) return rc;
  if (!p->d_x16) CK(cudaMalloc((void**)&p->d_x16, (size_t)p->Ucap * p->C * p->n_stride * sizeof(int16_t)));
  const size_t rows = (size_t)U * p->C;
  CK(cudaMemcpyAsync(p->d_x16, samples, rows * n * sizeof(int16_t), cudaMemcpyHostToDevice, p->stream));
  // The m = 4, r = 1 analysis kernel reads 16-bit PCM directly (rows must start on 16-byte boundaries); BTKB_ANALYSIS_I16=0 or any other
  // filter-bank shape converts to float first (k_i16_to_f32), which gives the same snapshots bit for bit (the conversion is exact).
  static const bool direct = [] { const char* e = getenv("BTKB_ANALYSIS_I16"); return !(e && atoi(e) == 0); }();
  if (direct && n % 8 == 0 && p->m == 4 && p->cfg.r == 1) { p->x16_cur = p->d_x16; p->x16_stride = n; p->x_cur = p->d_x; return BTKB_OK; }
  k_i16_to_f32<<<148 * 8, 256, 0, p->stream>>>(p->d_x16, p->d_x, rows, n, p->n_stride);
  CK(cudaGetLastError());
  p->x_cur = p->d_x;
  return BTKB_OK;
}

int btkb_submit_device(btkb_pipeline* p, const float* d_samples, int U, int n, const int* lengths) {
  if (!p || !d_samples) return fail(BTKB_ERR_INVALID, "btkb_submit_device: null argument");
  CK(cudaSetDevice(p->cfg.device));
  int rc = submit_common(p, U, n, lengths); if (rc) return rc;
  if (n % 4 == 0 && n == p->n_stride) { p->x_cur = d_samples; return BTKB_OK; }
  CK(cudaMemcpy2DAsync(p->d_x, (size_t)p->n_stride * sizeof(float), d_samples, (size_t)n * sizeof(float), (size_t)n * sizeof(float), (size_t)U * p->C,
                       cudaMemcpyDeviceToDevice, p->stream));
  p->x_cur = p->d_x;
  return BTKB_OK;
}

static int do_analysis(btkb_pipeline* p) {
  if (!p->have_h) return fail(BTKB_ERR_STATE, "btkb_run: set the analysis prototype first");
  if (p->U == 0) return fail(BTKB_ERR_STATE, "btkb_run: no batch submitted");
  AnalysisArgs a{p->x_cur, p->d_len, p->d_h, p->d_X, p->d_E, p->U, p->C, p->n, (p->x_cur == p->d_x) ? p->n_stride : p->n, p->T, p->M, p->m, p->D, p->laN,
                 p->Gp, 1, p->d_tw, 1, 0, (long long)(p->laN + 1) * p->D, 0, p->x16_cur, p->x16_stride, p->Cp};
  CK(launch_analysis(a, p->stream));
  p->launches++;
  p->have_X = true;
  return BTKB_OK;
}

static PerBinArgs perbin_args(btkb_pipeline* p) {
  PerBinArgs a;
  memset(&a, 0, sizeof(a));
  a.X = p->d_X; a.E = p->d_E; a.lengths = p->d_len; a.W = p->d_W; a.TA = p->have_ta ? p->d_TA : nullptr;
  a.WL = p->have_wl ? p->d_WL : nullptr;
  a.Y = p->Y_out; a.PFW = p->d_PFW; a.UA = p->d_UA; a.stats_updates = p->d_upd;
  a.R = p->d_R; a.noise_mask = p->d_mask; a.noise_count = p->d_count;
  a.U = p->U; a.C = p->Cp; a.Ctrue = p->C; a.T = p->T; a.M = p->M; a.K = p->K; a.G = p->U * p->K; a.Gp = p->Gp; a.D = p->D; a.laN = p->laN; a.pdA = p->pdA;
  a.kind = p->cfg.beamformer; a.normalize_weight = p->cfg.normalize_weight; a.pf_kind = p->cfg.postfilter; a.pf_alpha = p->cfg.pf_alpha; a.pf_type = p->cfg.pf_type; a.pf_min_frames = p->cfg.pf_min_frames;
  const btkb_lms_params& l = p->cfg.lms;
  a.lms = LmsArgs{l.beta, l.gamma, l.init_diagonal_load, l.regularization_param, l.energy_floor, l.sil_thresh, l.max_wa_l2norm, l.min_frames, l.slowdown_after};
  const btkb_rls_params& q = p->cfg.rls;
  a.rls = RlsArgs{q.beta, q.gamma, q.mu, q.init_diagonal_load, q.regularization_param, q.sil_thresh, q.alpha2, q.max_wa_l2norm, q.constraint_option, q.min_frames};
  const btkb_rls_cpp_params& rc = p->cfg.rls_cpp;
  a.rlsc = RlsCppArgs{rc.mu, rc.sigma2, rc.init_sigma2, rc.alpha, rc.qctype, rc.update};
  return a;
}

static int do_wpe(btkb_pipeline* p, int start_frame_no, int end_frame_no, bool apply_only = false) {
  if (!p->cfg.wpe.enabled) return fail(BTKB_ERR_STATE, "btkb_run_wpe: the pipeline was created without cfg.wpe.enabled");
  if (!p->have_X) return fail(BTKB_ERR_STATE, "btkb_run_wpe: run the analysis first");
  const btkb_wpe_params& w = p->cfg.wpe;
  WpeArgs a;
  memset(&a, 0, sizeof(a));
  a.X = p->d_X; a.lengths = p->d_len; a.S = p->d_wS; a.TH = p->d_wTH; a.Gf = p->d_wG; a.Rw = p->d_wR; a.err_flag = p->d_werr;
  a.U = p->U; a.C = p->C; a.T = p->T; a.Ts = round_up(p->T, 2); a.K = p->K; a.G = p->U * p->K; a.Gp = p->Gp; a.D = p->D; a.laN = p->laN; a.pdA = p->pdA;
  a.lowerN = w.lower_num; a.P = p->wpe_P; a.L = p->wpe_L; a.Lr = p->wpe_Lr; a.iterations = w.iterations_num; a.nbins = p->wpe_nbins;
  a.est_frames = (end_frame_no < 0) ? -1 : std::max(0, end_frame_no - std::max(start_frame_no, 0));   // fill_buffer_ (:500-534) never skips input frames
  a.load_factor = (float)pow(10.0, w.load_db / 10.0); a.diagonal_bias = (float)w.diagonal_bias;
  a.apply_only = apply_only ? 1 : 0;
  a.slot = p->wpe_slot; a.chunk_frame = p->wpe_chunk_frame; a.chol_threads = p->wpe_chol_threads;
  { const char* pf = getenv("BTKB_WPE_PREFETCH"); a.prefetch = pf ? atoi(pf) : 1; }
  { const char* mm = getenv("BTKB_WPE_MMA"); a.mma = mm ? atoi(mm) : 1; }
  a.Sd = std::max(((a.est_frames >= 0) ? std::min(a.T, a.est_frames) : a.T) - a.lowerN, 0);
  a.form = (p->wpe_form >= 0) ? p->wpe_form : (a.Sd < a.L ? 1 : 0);   // the frame-domain system has S <= Sd rows, the lag-domain one L
  CK(cudaMemsetAsync(p->d_werr, 0, sizeof(int), p->stream));
  CK(cudaEventRecord(p->wev[0], p->stream));
  CK(launch_wpe(a, p->wpe_chunk, p->cfg.wpe.fp32_normal_equations, p->stream, &p->launches));
  CK(cudaEventRecord(p->wev[1], p->stream));
  int err = 0;
  CK(cudaMemcpyAsync(&err, p->d_werr, sizeof(int), cudaMemcpyDeviceToHost, p->stream));
  CK(cudaStreamSynchronize(p->stream));
  if (err)  // dereverberation.cc:676-678
    return fail(BTKB_ERR_INVALID, "MultiChannelWPEDereverberation: GSL Cholesky decomposition failed.\nSome channels may be too similar. Try to increase 'diagonal_bias' or use 'SingleChannelWPEDereverberationFeature' for each channel");
  p->have_wpe = true;
  if (!apply_only) { p->wpe_U = p->U; p->wpe_last_form = a.form; }
  return BTKB_OK;
}

static int do_beamformer(btkb_pipeline* p) {
  if (!p->have_X) return fail(BTKB_ERR_STATE, "btkb_run_beamformer: run the analysis first");
  if (!p->have_w) {
    if (p->cfg.beamformer == BTKB_BF_MVDR) return fail(BTKB_ERR_STATE, "call calc_mvdr_weights() once");            // beamformer.cc:2544-2546
    return fail(BTKB_ERR_STATE, "call calc_array_manifold_vectorsX() once");                                          // beamformer.cc:1098-1100
  }
  if (p->wU != p->U) return fail(BTKB_ERR_INVALID, "btkb_run_beamformer: weights were set for a different number of utterances");
  const bool narrow = (p->C >= 1 && p->C <= 8), wide = !narrow;   // wide arrays run the lane-split kernel on 16 / 32 / 64 channel rows (zero-padded)
  if (wide && p->Cp != p->C && p->cfg.beamformer == BTKB_BF_MVDR)
    return fail(BTKB_ERR_INVALID, "btkb_run_beamformer: MVDR on a wide array needs 16, 32 or 64 channels (a zero-padded covariance is singular); got " + std::to_string(p->C));
  if (wide && p->cfg.postfilter != BTKB_PF_NONE)
    return fail(BTKB_ERR_INVALID, "btkb_run_beamformer: the post-filters are built for <= 8 channels (their C(C-1)/2 cross-spectral densities must fit the register file)");
  PerBinArgs a = perbin_args(p);
  if (p->cfg.postfilter >= BTKB_PF_MCCOWAN) {
    if (!p->have_pfR) return fail(BTKB_ERR_STATE, "McCowanPostFilter:  construct/set a noise coherence matrix");  // postfilter.cc:828-830
    const bool lef = p->cfg.postfilter == BTKB_PF_LEFKIMMIATIS;
    CK(launch_pf_prepare(p->d_pfR, p->d_pfInvR, p->d_pfQ, p->C, p->K, p->cfg.pf_threshold, p->cfg.pf_min_sv, lef ? 1 : 0, p->stream));
    p->launches++;
    if (lef) {  // Lambda uses arrayManifold() = the delay-and-sum manifold (postfilter.cc:984-987)
      CK(launch_pf_lambda(p->d_pfInvR, p->have_ta ? p->d_TA : p->d_W, p->d_LAM, p->U, p->C, p->K, p->Gp, p->cfg.pf_type, p->stream));
      p->launches++;
    }
    a.PFQ = p->d_pfQ; a.LAM = p->d_LAM; a.pf_fbin1 = p->cfg.pf_fbin1;
  }
  if (p->cfg.beamformer == BTKB_BF_GSC_RLS_CPP) {
    if (!narrow) return fail(BTKB_ERR_INVALID, "btkb_run_beamformer: the C++ SubbandGSCRLS kernel is built for 2..8 channels");
    a.WL = p->d_WL;   // final wl = B wa, readable through btkb_get_sidelobe_weights
    CK(launch_perbin_rls_cpp(a, p->d_delays, p->cfg.samplerate, p->stream));
    p->have_wl = true;
  }
  else if (narrow) CK(launch_perbin(a, p->stream)); else CK(launch_perbin_wide(a, p->stream));
  p->launches++;
  p->have_Y = true; p->pf_applied = true;
  p->have_ua = (p->cfg.beamformer == BTKB_BF_GSC_LMS || p->cfg.beamformer == BTKB_BF_GSC_RLS);
  return BTKB_OK;
}

static int do_synthesis(btkb_pipeline* p) {
  if (!p->have_g) return fail(BTKB_ERR_STATE, "btkb_run: set the synthesis prototype first");
  if (!p->have_Y) return fail(BTKB_ERR_STATE, "btkb_run: no beamformer output to synthesise");
  CK(cudaMemsetAsync(p->d_stats, 0, (size_t)p->U * 3 * sizeof(double), p->stream));
  SynthesisArgs a{p->Y_out, p->d_len, p->d_g, p->d_time, p->d_stats, p->U, p->n, p->T, p->M, p->m, p->cfg.r, p->D, p->K, p->Gp, p->pdS, p->laN, p->pdA,
                  p->nb, p->nb * p->D, p->cfg.synthesis_gain, p->d_tw,
                  (p->cfg.postfilter >= BTKB_PF_MCCOWAN && p->pf_applied) ? p->cfg.pf_min_frames + 1 : 0, 0, 0, nullptr, 0};
  CK(launch_synthesis(a, p->stream));
  p->launches++;
  p->have_time = true;
  return BTKB_OK;
}

int btkb_run_analysis(btkb_pipeline* p) {
  if (!p) return fail(BTKB_ERR_INVALID, "null pipeline");
  CK(cudaSetDevice(p->cfg.device));
  p->launches = 0;
  CK(cudaEventRecord(p->ev[0], p->stream));
  int rc = do_analysis(p); if (rc) return rc;
  CK(cudaEventRecord(p->ev[1], p->stream));
  CK(cudaEventRecord(p->ev[2], p->stream));
  CK(cudaEventRecord(p->ev[3], p->stream));
  return BTKB_OK;
}

int btkb_accumulate_covariance(btkb_pipeline* p, const double* labels, float energy_threshold) {
  if (!p) return fail(BTKB_ERR_INVALID, "null pipeline");
  if (!p->have_X) return fail(BTKB_ERR_STATE, "btkb_accumulate_covariance: run the analysis first");
  if (!p->d_R) return fail(BTKB_ERR_STATE, "btkb_accumulate_covariance: pipeline was not created with BTKB_BF_MVDR");
  if (p->Cp != p->C) return fail(BTKB_ERR_INVALID, "btkb_accumulate_covariance: wide arrays need 16, 32 or 64 channels");
  CK(cudaSetDevice(p->cfg.device));
  if (labels) {
    CK(cudaMemcpyAsync(p->d_labels, labels, (size_t)p->U * 2 * sizeof(double), cudaMemcpyHostToDevice, p->stream));
  }
  CK(launch_noise_mask(p->d_E, p->d_len, labels ? p->d_labels : nullptr, p->d_mask, p->d_count, p->U, p->T, p->D, p->laN, p->pdA, p->cfg.samplerate,
                       energy_threshold, p->stream));
  PerBinArgs a = perbin_args(p);
  // 64 microphones: the Gram is a batched dense contraction and runs on the tensor cores (btkb_cov_tc.cu: tcgen05 + TMEM,
  // 3 x TF32 split); BTKB_COV_TC=0 selects the CUDA-core kernel (kept for A/B checks)
  static const bool use_tc = [] { const char* e = getenv("BTKB_COV_TC"); return !(e && atoi(e) == 0); }();
  if (p->C <= 8) CK(launch_covariance(a, p->stream));
  else if (p->C == 64 && use_tc) {
    const size_t need = covariance_tc_workspace_bytes(p->Ucap * p->K, p->C, p->Tcap);
    if (!p->d_covS) CK(cudaMalloc((void**)&p->d_covS, need));
    a.Scov = p->d_covS;
    CK(launch_covariance_tc(a, p->stream));
  }
  else CK(launch_covariance_wide(a, p->stream));
  if (labels) CK(cudaStreamSynchronize(p->stream));
  p->have_R = true; p->R_is_sum = true;
  p->wU = p->U;
  return BTKB_OK;
}

// ---------------------------------------------------------------------------------------------- SOS batch beamformers
static int ensure_scratch(btkb_pipeline* p, size_t bytes);
static int sos_check(btkb_pipeline* p, const char* who, bool need_X) {
  if (!p) return fail(BTKB_ERR_INVALID, std::string(who) + ": null pipeline");
  if (p->cfg.beamformer != BTKB_BF_DS) return fail(BTKB_ERR_STATE, std::string(who) + ": the SOS beamformers apply their weights like SubbandDS; create the pipeline with BTKB_BF_DS");
  if (p->C < 2 || p->C > 8) return fail(BTKB_ERR_INVALID, std::string(who) + ": built for 2..8 channels");
  if (need_X && !p->have_X) return fail(BTKB_ERR_STATE, std::string(who) + ": run the analysis first");
  if (need_X && p->have_sos && p->sos_U != p->U)
    return fail(BTKB_ERR_INVALID, std::string(who) + ": statistics were accumulated for " + std::to_string(p->sos_U) + " utterances, this batch has " +
                                      std::to_string(p->U) + " (call btkb_sos_reset_stats first)");
  return BTKB_OK;
}
static int sos_alloc(btkb_pipeline* p) {
  if (p->d_sosR) return BTKB_OK;
  const size_t G = (size_t)p->Gpcap;
  CK(cudaMalloc((void**)&p->d_sosR, 2 * (size_t)p->C * p->C * G * sizeof(double2)));
  CK(cudaMalloc((void**)&p->d_sosWd, (size_t)p->C * G * sizeof(double2)));
  CK(cudaMalloc((void**)&p->d_sosCnt, 2 * G * sizeof(double)));
  CK(cudaMalloc((void**)&p->d_sosWtu, 2 * (size_t)p->Tcap * p->Ucap * sizeof(float)));
  CK(cudaMalloc((void**)&p->d_sosErr, sizeof(int)));
  return BTKB_OK;
}
static SosArgs sos_args(btkb_pipeline* p) {
  SosArgs a;
  memset(&a, 0, sizeof(a));
  a.X = p->d_X; a.E = p->d_E; a.lengths = p->d_len; a.wtu = p->d_sosWtu; a.Rs = p->d_sosR; a.cnt = p->d_sosCnt; a.Wd = p->d_sosWd; a.W = p->d_W; a.err = p->d_sosErr;
  a.accumulate = p->have_sos ? 1 : 0;
  a.U = p->U; a.C = p->C; a.T = p->T; a.K = p->K; a.G = p->U * p->K; a.Gp = p->Gp; a.D = p->D; a.laN = p->laN; a.pdA = p->pdA;
  a.samplerate = p->cfg.samplerate;
  return a;
}

int btkb_sos_reset_stats(btkb_pipeline* p) {
  int rc = sos_check(p, "btkb_sos_reset_stats", false); if (rc) return rc;
  p->have_sos = false;
  return BTKB_OK;
}

int btkb_sos_accumulate_from_label(btkb_pipeline* p, const double* labels, int NL, float energy_threshold) {
  int rc = sos_check(p, "btkb_sos_accumulate_from_label", true); if (rc) return rc;
  if (!labels || NL < 1) return fail(BTKB_ERR_INVALID, "btkb_sos_accumulate_from_label: need at least one (start, end) segment per utterance");
  CK(cudaSetDevice(p->cfg.device));
  rc = sos_alloc(p); if (rc) return rc;
  if (NL > p->sos_NLcap) {
    if (p->d_sosLab) { CK(cudaStreamSynchronize(p->stream)); cudaFree(p->d_sosLab); p->d_sosLab = nullptr; }
    CK(cudaMalloc((void**)&p->d_sosLab, (size_t)p->Ucap * NL * 2 * sizeof(double)));
    p->sos_NLcap = NL;
  }
  CK(cudaMemcpyAsync(p->d_sosLab, labels, (size_t)p->U * NL * 2 * sizeof(double), cudaMemcpyHostToDevice, p->stream));
  SosArgs a = sos_args(p);
  a.labels = p->d_sosLab; a.NL = NL; a.thr = energy_threshold;
  CK(launch_sos_accumulate(a, p->stream, &p->launches));
  CK(cudaStreamSynchronize(p->stream));   // `labels` may be pageable host memory
  p->have_sos = true; p->wU = p->U; p->sos_U = p->U;
  return BTKB_OK;
}

int btkb_sos_accumulate_from_tfmask(btkb_pipeline* p, const float* mask_t, const float* mask_j, int Tm, float energy_threshold) {
  int rc = sos_check(p, "btkb_sos_accumulate_from_tfmask", true); if (rc) return rc;
  if (!mask_t || !mask_j) return fail(BTKB_ERR_INVALID, "btkb_sos_accumulate_from_tfmask: null mask");
  if (Tm < p->T) return fail(BTKB_ERR_INVALID, "btkb_sos_accumulate_from_tfmask: the masks have " + std::to_string(Tm) + " frames, the batch has " + std::to_string(p->T) + " (the reference raises IndexError)");
  CK(cudaSetDevice(p->cfg.device));
  rc = sos_alloc(p); if (rc) return rc;
  if (!p->d_sosMask) CK(cudaMalloc((void**)&p->d_sosMask, 2 * (size_t)p->Tcap * p->Gpcap * sizeof(float)));
  const size_t raw = (size_t)p->U * Tm * p->K * sizeof(float);
  rc = ensure_scratch(p, raw); if (rc) return rc;
  float* mt = p->d_sosMask; float* mj = p->d_sosMask + (size_t)p->Tcap * p->Gpcap;
  CK(cudaMemcpyAsync(p->d_scratch, mask_t, raw, cudaMemcpyHostToDevice, p->stream));
  CK(launch_sos_scatter_mask((const float*)p->d_scratch, mt, p->U, Tm, p->T, p->K, p->Gp, p->stream));
  CK(cudaMemcpyAsync(p->d_scratch, mask_j, raw, cudaMemcpyHostToDevice, p->stream));
  CK(launch_sos_scatter_mask((const float*)p->d_scratch, mj, p->U, Tm, p->T, p->K, p->Gp, p->stream));
  p->launches += 2;
  SosArgs a = sos_args(p);
  a.mask_t = mt; a.mask_j = mj; a.thr = energy_threshold;
  CK(launch_sos_accumulate(a, p->stream, &p->launches));
  CK(cudaStreamSynchronize(p->stream));
  p->have_sos = true; p->wU = p->U; p->sos_U = p->U;
  return BTKB_OK;
}

int btkb_sos_calc_weights(btkb_pipeline* p, int kind, double gamma, int ref_micx, double offset) {
  int rc = sos_check(p, "btkb_sos_calc_weights", false); if (rc) return rc;
  if (kind != BTKB_SOS_BMVDR && kind != BTKB_SOS_GEV) return fail(BTKB_ERR_INVALID, "btkb_sos_calc_weights: unknown kind");
  if (!p->have_sos) return fail(BTKB_ERR_STATE, "No target signal SOS");   // pybeamformer.py:1270-1273
  if (ref_micx < 0 || ref_micx >= p->C) return fail(BTKB_ERR_INVALID, "btkb_sos_calc_weights: ref_micx out of range");
  if (!(offset >= 0.0 && offset <= 1.0)) return fail(BTKB_ERR_INVALID, "The offset value " + std::to_string(offset) + " is out of [0, 1]");   // :1274
  CK(cudaSetDevice(p->cfg.device));
  CK(cudaMemsetAsync(p->d_sosErr, 0, sizeof(int), p->stream));
  SosArgs a = sos_args(p);
  CK(launch_sos_solve(a, kind, gamma, ref_micx, offset, p->stream, &p->launches));
  int err = 0;
  CK(cudaMemcpyAsync(&err, p->d_sosErr, sizeof(int), cudaMemcpyDeviceToHost, p->stream));
  CK(cudaStreamSynchronize(p->stream));
  if (err & 1) return fail(BTKB_ERR_STATE, "No target signal stats accumulated; Use self.accu_stats_from_label() or accu_stats_from_tfmask()");   // :1264, 1316
  if (err & 2) return fail(BTKB_ERR_STATE, "No noise stats accumulated; Use self.accu_stats_from_label() or accu_stats_from_tfmask()");
  if (err & 4) return fail(BTKB_ERR_INVALID, kind == BTKB_SOS_BMVDR ? "Matrix inversion failed\nAdd a small value to the diagonal component of the covariance matrix"
                                                                   : "GEV failed\nAdd a small value to the diagonal component of the covariance matrix");
  p->have_w = true;
  if (!p->have_ta) {  // the time-alignment manifold is only used by post-filters; without delays it is the weight itself (as btkb_set_weights)
    CK(cudaMemcpyAsync(p->d_TA, p->d_W, (size_t)p->C * p->Gp * sizeof(float2), cudaMemcpyDeviceToDevice, p->stream));
    p->have_ta = true;
  }
  return BTKB_OK;
}

int btkb_sos_get_stats(btkb_pipeline* p, double* Rt, double* Rn, double* counts) {
  int rc = sos_check(p, "btkb_sos_get_stats", false); if (rc) return rc;
  if (!p->have_sos) return fail(BTKB_ERR_STATE, "btkb_sos_get_stats: no statistics accumulated");
  CK(cudaSetDevice(p->cfg.device));
  const int U = p->U, K = p->K, C = p->C, Gp = p->Gp;
  std::vector<double2> tmp((size_t)C * C * Gp);
  for (int set = 0; set < 2; set++) {
    double* dst = set == 0 ? Rt : Rn;
    if (!dst) continue;
    CK(cudaMemcpyAsync(tmp.data(), p->d_sosR + (size_t)set * C * C * Gp, tmp.size() * sizeof(double2), cudaMemcpyDeviceToHost, p->stream));
    CK(cudaStreamSynchronize(p->stream));
    for (int u = 0; u < U; u++)
      for (int k = 0; k < K; k++)
        for (int e = 0; e < C * C; e++) {
          const double2 v = tmp[(size_t)e * Gp + (size_t)u * K + k];
          double* d = dst + 2 * (((size_t)u * K + k) * C * C + e);
          d[0] = v.x; d[1] = v.y;
        }
  }
  if (counts) {
    std::vector<double> c2((size_t)2 * Gp);
    CK(cudaMemcpyAsync(c2.data(), p->d_sosCnt, c2.size() * sizeof(double), cudaMemcpyDeviceToHost, p->stream));
    CK(cudaStreamSynchronize(p->stream));
    for (int u = 0; u < U; u++)
      for (int k = 0; k < K; k++) {
        counts[2 * ((size_t)u * K + k)] = c2[(size_t)u * K + k];
        counts[2 * ((size_t)u * K + k) + 1] = c2[(size_t)Gp + (size_t)u * K + k];
      }
  }
  return BTKB_OK;
}

int btkb_spectral_matrix_update(btkb_pipeline* p, float mu, int legacy_noconj) {
  if (!p) return fail(BTKB_ERR_INVALID, "null pipeline");
  if (!p->have_X) return fail(BTKB_ERR_STATE, "btkb_spectral_matrix_update: run the analysis first");
  if (p->C > 8) return fail(BTKB_ERR_INVALID, "btkb_spectral_matrix_update: built for <= 8 channels");
  CK(cudaSetDevice(p->cfg.device));
  if (!p->d_R) CK(cudaMalloc((void**)&p->d_R, (size_t)p->C * p->C * p->Gpcap * sizeof(float2)));
  PerBinArgs a = perbin_args(p);
  CK(launch_spectral_recursion(a, mu, legacy_noconj, p->stream));
  p->have_R = true; p->R_is_sum = false; p->wU = p->U;
  return BTKB_OK;
}

int btkb_run_wpe(btkb_pipeline* p, int start_frame_no, int end_frame_no) {
  if (!p) return fail(BTKB_ERR_INVALID, "null pipeline");
  CK(cudaSetDevice(p->cfg.device));
  p->launches = 0;
  return do_wpe(p, start_frame_no, end_frame_no);
}

int btkb_apply_wpe(btkb_pipeline* p) {
  if (!p) return fail(BTKB_ERR_INVALID, "null pipeline");
  if (!p->have_wpe) return fail(BTKB_ERR_STATE, "Call MultiChannelWPEDereverberation::estimate_filter()");   // dereverberation.cc:446-447 (jinitialization_error)
  if (p->U != p->wpe_U) return fail(BTKB_ERR_INVALID, "btkb_apply_wpe: the filters were estimated for a different number of utterances");
  CK(cudaSetDevice(p->cfg.device));
  p->launches = 0;
  return do_wpe(p, 0, -1, true);
}

int btkb_get_wpe_filter(btkb_pipeline* p, float* out) {
  if (!p || !out) return fail(BTKB_ERR_INVALID, "null argument");
  if (!p->have_wpe) return fail(BTKB_ERR_STATE, "Call estimate_filter() first");
  CK(cudaSetDevice(p->cfg.device));
  CK(cudaMemcpyAsync(out, p->d_wG, (size_t)p->U * p->K * p->C * p->wpe_L * sizeof(float2), cudaMemcpyDeviceToHost, p->stream));
  CK(cudaStreamSynchronize(p->stream));
  return BTKB_OK;
}

int btkb_set_wpe_filter(btkb_pipeline* p, int U, const float* G) {
  if (!p || !G) return fail(BTKB_ERR_INVALID, "null argument");
  if (!p->cfg.wpe.enabled) return fail(BTKB_ERR_STATE, "btkb_set_wpe_filter: the pipeline was created without cfg.wpe.enabled");
  if (U < 1 || U > p->Ucap) return fail(BTKB_ERR_INVALID, "btkb_set_wpe_filter: U out of range");
  CK(cudaSetDevice(p->cfg.device));
  CK(cudaMemcpyAsync(p->d_wG, G, (size_t)U * p->K * p->C * p->wpe_L * sizeof(float2), cudaMemcpyHostToDevice, p->stream));
  CK(cudaStreamSynchronize(p->stream));
  p->have_wpe = true; p->wpe_U = U;
  return BTKB_OK;
}

int btkb_last_timing_wpe(btkb_pipeline* p, float* ms) {
  if (!p || !ms) return fail(BTKB_ERR_INVALID, "null argument");
  if (!p->have_wpe) return fail(BTKB_ERR_STATE, "btkb_last_timing_wpe: WPE has not run");
  CK(cudaSetDevice(p->cfg.device));
  CK(cudaEventSynchronize(p->wev[1]));
  CK(cudaEventElapsedTime(ms, p->wev[0], p->wev[1]));
  return BTKB_OK;
}

int btkb_last_wpe_form(btkb_pipeline* p, int* form) {
  if (!p || !form) return fail(BTKB_ERR_INVALID, "null argument");
  if (!p->have_wpe || p->wpe_last_form < 0) return fail(BTKB_ERR_STATE, "btkb_last_wpe_form: no filter has been estimated by this pipeline");
  *form = p->wpe_last_form;
  return BTKB_OK;
}

int btkb_run_beamformer(btkb_pipeline* p, int do_syn) {
  if (!p) return fail(BTKB_ERR_INVALID, "null pipeline");
  CK(cudaSetDevice(p->cfg.device));
  p->launches = 0;
  CK(cudaEventRecord(p->ev[0], p->stream));
  CK(cudaEventRecord(p->ev[1], p->stream));
  int rc = do_beamformer(p); if (rc) return rc;
  CK(cudaEventRecord(p->ev[2], p->stream));
  if (do_syn) { rc = do_synthesis(p); if (rc) return rc; }
  CK(cudaEventRecord(p->ev[3], p->stream));
  return BTKB_OK;
}

int btkb_set_subband(btkb_pipeline* p, int U, int T, const float* Y) {
  if (!p || !Y) return fail(BTKB_ERR_INVALID, "btkb_set_subband: null argument");
  if (U < 1 || U > p->Ucap || T < 0 || T > p->Tcap) return fail(BTKB_ERR_INVALID, "btkb_set_subband: U or T exceeds the pipeline capacity");
  CK(cudaSetDevice(p->cfg.device));
  // a length that yields exactly T frames: T = ceil(len/D) - laN + pdA
  const int nblk = T - p->pdA + p->laN;
  if (nblk < 0) return fail(BTKB_ERR_INVALID, "btkb_set_subband: T is shorter than the filter-bank delay");
  p->U = U; p->n = nblk * p->D; p->lengths.assign(U, nblk * p->D);
  p->T = T; p->nb = std::max(T - p->pdS, 0); p->Gp = round_up(U * p->K, 128);
  p->streaming = false; p->Y_out = p->d_Y;
  CK(cudaMemcpyAsync(p->d_len, p->lengths.data(), U * sizeof(int), cudaMemcpyHostToDevice, p->stream));
  std::vector<float2> tmp((size_t)T * p->Gp, make_float2(0.f, 0.f));
  for (int u = 0; u < U; u++)
    for (int t = 0; t < T; t++)
      memcpy(&tmp[(size_t)t * p->Gp + (size_t)u * p->K], Y + 2 * (((size_t)u * T + t) * p->K), sizeof(float2) * p->K);
  CK(cudaMemcpyAsync(p->d_Y, tmp.data(), tmp.size() * sizeof(float2), cudaMemcpyHostToDevice, p->stream));
  CK(cudaStreamSynchronize(p->stream));
  p->have_Y = true; p->have_X = false; p->have_time = false; p->have_ua = false; p->pf_applied = false;
  return BTKB_OK;
}

int btkb_set_snapshots(btkb_pipeline* p, int U, int T, const float* X) {
  if (!p || !X) return fail(BTKB_ERR_INVALID, "btkb_set_snapshots: null argument");
  if (U < 1 || U > p->Ucap || T < 0 || T > p->Tcap) return fail(BTKB_ERR_INVALID, "btkb_set_snapshots: U or T exceeds the pipeline capacity");
  if (p->cfg.wpe.enabled) return fail(BTKB_ERR_INVALID, "btkb_set_snapshots: not offered for a pipeline with cfg.wpe.enabled");
  CK(cudaSetDevice(p->cfg.device));
  const int nblk = T - p->pdA + p->laN;   // a length that yields exactly T frames: T = ceil(len/D) - laN + pdA
  if (nblk < 0) return fail(BTKB_ERR_INVALID, "btkb_set_snapshots: T is shorter than the filter-bank delay");
  if (p->wU != 0 && p->wU != U) { p->have_w = p->have_wl = p->have_ta = p->have_R = false; p->R_is_sum = false; p->NC = 1; p->wU = 0; }
  p->U = U; p->n = nblk * p->D; p->lengths.assign(U, nblk * p->D);
  p->T = T; p->nb = std::max(T - p->pdS, 0); p->Gp = round_up(U * p->K, 128);
  p->streaming = false; p->Y_out = p->d_Y;
  CK(cudaMemcpyAsync(p->d_len, p->lengths.data(), U * sizeof(int), cudaMemcpyHostToDevice, p->stream));
  std::vector<float2> tmp((size_t)T * p->Cp * p->Gp, make_float2(0.f, 0.f));
  for (int u = 0; u < U; u++)
    for (int t = 0; t < T; t++)
      for (int c = 0; c < p->C; c++)
        memcpy(&tmp[((size_t)t * p->Cp + c) * p->Gp + (size_t)u * p->K], X + 2 * ((((size_t)u * T + t) * p->C + c) * p->K), sizeof(float2) * p->K);
  CK(cudaMemcpyAsync(p->d_X, tmp.data(), tmp.size() * sizeof(float2), cudaMemcpyHostToDevice, p->stream));
  CK(cudaMemsetAsync(p->d_E, 0, (size_t)std::max(T, 1) * U * sizeof(float), p->stream));   // channel-0 frame energy: not known for foreign snapshots
  CK(cudaStreamSynchronize(p->stream));
  p->have_X = true; p->have_Y = false; p->have_time = false; p->have_ua = false; p->pf_applied = false;
  return BTKB_OK;
}

int btkb_run_synthesis(btkb_pipeline* p) {
  if (!p) return fail(BTKB_ERR_INVALID, "null pipeline");
  CK(cudaSetDevice(p->cfg.device));
  p->launches = 0;
  CK(cudaEventRecord(p->ev[0], p->stream));
  CK(cudaEventRecord(p->ev[1], p->stream));
  CK(cudaEventRecord(p->ev[2], p->stream));
  int rc = do_synthesis(p); if (rc) return rc;
  CK(cudaEventRecord(p->ev[3], p->stream));
  return BTKB_OK;
}

int btkb_run(btkb_pipeline* p, int do_syn) {
  if (!p) return fail(BTKB_ERR_INVALID, "null pipeline");
  CK(cudaSetDevice(p->cfg.device));
  p->launches = 0;
  CK(cudaEventRecord(p->ev[0], p->stream));
  // BTKB_FUSED=1: analysis and the per-bin NLMS recurrence in one kernel, the snapshots never reaching HBM (btkb_fused.cu; measured
  // slower than the two kernels, DESIGN.md 10.1, hence opt-in).  Its time is reported in the "analysis" segment.
  const bool want_fused = [] { const char* e = getenv("BTKB_FUSED"); return e && atoi(e) != 0; }();   // read at every call so that one process can compare
  if (want_fused && !p->cfg.wpe.enabled && p->have_h && p->U > 0 && p->have_w && p->wU == p->U) {
    AnalysisArgs fa{p->x_cur, p->d_len, p->d_h, p->d_X, p->d_E, p->U, p->C, p->n, (p->x_cur == p->d_x) ? p->n_stride : p->n, p->T, p->M, p->m, p->D, p->laN,
                    p->Gp, 1, p->d_tw, 1, 0, (long long)(p->laN + 1) * p->D, 0, p->x16_cur, p->x16_stride, p->Cp};
    PerBinArgs fb = perbin_args(p);
    if (fused_supported(fa, fb)) {
      CK(launch_fused_analysis_nlms(fa, fb, p->stream));
      p->launches++;
      p->have_X = false; p->have_Y = true; p->pf_applied = true; p->have_ua = true;
      CK(cudaEventRecord(p->ev[1], p->stream));
      CK(cudaEventRecord(p->ev[2], p->stream));
      if (do_syn) { int rc2 = do_synthesis(p); if (rc2) return rc2; }
      CK(cudaEventRecord(p->ev[3], p->stream));
      return BTKB_OK;
    }
  }
  int rc = do_analysis(p); if (rc) return rc;
  if (p->cfg.wpe.enabled) { rc = do_wpe(p, 0, -1); if (rc) return rc; }   // counted in the "analysis" segment of btkb_last_timing; btkb_last_timing_wpe isolates it
  CK(cudaEventRecord(p->ev[1], p->stream));
  rc = do_beamformer(p); if (rc) return rc;
  CK(cudaEventRecord(p->ev[2], p->stream));
  if (do_syn) { rc = do_synthesis(p); if (rc) return rc; }
  CK(cudaEventRecord(p->ev[3], p->stream));
  return BTKB_OK;
}

int btkb_synchronize(btkb_pipeline* p) {
  if (!p) return fail(BTKB_ERR_INVALID, "null pipeline");
  CK(cudaSetDevice(p->cfg.device));
  CK(cudaStreamSynchronize(p->stream));
  return BTKB_OK;
}

int btkb_last_timing(btkb_pipeline* p, float* out5) {
  if (!p || !out5) return fail(BTKB_ERR_INVALID, "null argument");
  CK(cudaSetDevice(p->cfg.device));
  CK(cudaEventSynchronize(p->ev[3]));
  CK(cudaEventElapsedTime(&out5[0], p->ev[0], p->ev[3]));
  CK(cudaEventElapsedTime(&out5[1], p->ev[0], p->ev[1]));
  CK(cudaEventElapsedTime(&out5[2], p->ev[1], p->ev[2]));
  CK(cudaEventElapsedTime(&out5[3], p->ev[2], p->ev[3]));
  out5[4] = (float)p->launches;
  return BTKB_OK;
}

int btkb_num_frames(const btkb_pipeline* p) { return p ? p->T : 0; }
int btkb_num_frames_of(const btkb_pipeline* p, int u) {
  if (!p || u < 0 || u >= p->U) return 0;
  return p->streaming ? std::min(std::max(p->s_tu[u] - p->s_tlast, 0), p->T) : frames_of(p->lengths[u], p->D, p->laN, p->pdA);
}
int btkb_num_blocks(const btkb_pipeline* p) { return p ? p->nb : 0; }

static int ensure_scratch(btkb_pipeline* p, size_t bytes) {
  if (p->scratch_bytes >= bytes) return BTKB_OK;
  if (p->d_scratch) { cudaFree(p->d_scratch); p->d_scratch = nullptr; p->scratch_bytes = 0; }
  CK(cudaMalloc(&p->d_scratch, bytes));
  p->scratch_bytes = bytes;
  return BTKB_OK;
}

// device [T][rows][Gp] complex -> packed [U][T][rows][K]
__global__ void k_gather(const float2* src, float2* dst, int U, int T, int rows, int K, int Gp, int row_pitch) {   // row_pitch >= rows: rows per frame in src
  const size_t total = (size_t)U * T * rows * K;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    int k = (int)(i % K); size_t r = i / K;
    int c = (int)(r % rows); r /= rows;
    int t = (int)(r % T); int u = (int)(r / T);
    dst[i] = src[((size_t)t * row_pitch + c) * Gp + (size_t)u * K + k];
  }
}
__global__ void k_gather_f(const float* src, float* dst, int U, int T, int K, int Gp) {
  const size_t total = (size_t)U * T * K;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    int k = (int)(i % K); size_t r = i / K;
    int t = (int)(r % T); int u = (int)(r / T);
    dst[i] = src[(size_t)t * Gp + (size_t)u * K + k];
  }
}

int btkb_fetch_subband(btkb_pipeline* p, float* out) {
  if (!p || !out) return fail(BTKB_ERR_INVALID, "null argument");
  if (!p->have_Y) return fail(BTKB_ERR_STATE, "btkb_fetch_subband: nothing has been beamformed");
  CK(cudaSetDevice(p->cfg.device));
  const size_t bytes = (size_t)p->U * p->T * p->K * sizeof(float2);
  int rc = ensure_scratch(p, bytes); if (rc) return rc;
  k_gather<<<2048, 256, 0, p->stream>>>(p->Y_out, (float2*)p->d_scratch, p->U, p->T, 1, p->K, p->Gp, 1);
  CK(cudaGetLastError());
  CK(cudaMemcpyAsync(out, p->d_scratch, bytes, cudaMemcpyDeviceToHost, p->stream));
  CK(cudaStreamSynchronize(p->stream));
  return BTKB_OK;
}

int btkb_fetch_snapshots(btkb_pipeline* p, float* out) {
  if (!p || !out) return fail(BTKB_ERR_INVALID, "null argument");
  if (!p->have_X) return fail(BTKB_ERR_STATE, "btkb_fetch_snapshots: run the analysis first");
  CK(cudaSetDevice(p->cfg.device));
  const size_t bytes = (size_t)p->U * p->T * p->C * p->K * sizeof(float2);
  int rc = ensure_scratch(p, bytes); if (rc) return rc;
  k_gather<<<2048, 256, 0, p->stream>>>(p->d_X, (float2*)p->d_scratch, p->U, p->T, p->C, p->K, p->Gp, p->Cp);
  CK(cudaGetLastError());
  CK(cudaMemcpyAsync(out, p->d_scratch, bytes, cudaMemcpyDeviceToHost, p->stream));
  CK(cudaStreamSynchronize(p->stream));
  return BTKB_OK;
}

int btkb_get_postfilter_weights(btkb_pipeline* p, float* out) {
  if (!p || !out) return fail(BTKB_ERR_INVALID, "null argument");
  if (!p->d_PFW || !p->have_Y) return fail(BTKB_ERR_STATE, "btkb_get_postfilter_weights: no post-filter output");
  CK(cudaSetDevice(p->cfg.device));
  const size_t bytes = (size_t)p->U * p->T * p->K * sizeof(float);
  int rc = ensure_scratch(p, bytes); if (rc) return rc;
  k_gather_f<<<2048, 256, 0, p->stream>>>(p->d_PFW, (float*)p->d_scratch, p->U, p->T, p->K, p->Gp);
  CK(cudaGetLastError());
  CK(cudaMemcpyAsync(out, p->d_scratch, bytes, cudaMemcpyDeviceToHost, p->stream));
  CK(cudaStreamSynchronize(p->stream));
  return BTKB_OK;
}

int btkb_fetch_time(btkb_pipeline* p, float* out) {
  if (!p || !out) return fail(BTKB_ERR_INVALID, "null argument");
  if (!p->have_time) return fail(BTKB_ERR_STATE, "btkb_fetch_time: synthesis has not run");
  CK(cudaSetDevice(p->cfg.device));
  CK(cudaMemcpyAsync(out, p->d_time, (size_t)p->U * p->nb * p->D * sizeof(float), cudaMemcpyDeviceToHost, p->stream));
  CK(cudaStreamSynchronize(p->stream));
  return BTKB_OK;
}

int btkb_fetch_stats(btkb_pipeline* p, double* out) {
  if (!p || !out) return fail(BTKB_ERR_INVALID, "null argument");
  CK(cudaSetDevice(p->cfg.device));
  std::vector<float> upd(p->U, 0.f);
  if (p->have_time) CK(cudaMemcpyAsync(out, p->d_stats, (size_t)p->U * 3 * sizeof(double), cudaMemcpyDeviceToHost, p->stream));
  else memset(out, 0, (size_t)p->U * 3 * sizeof(double));
  if (p->have_ua) CK(cudaMemcpyAsync(upd.data(), p->d_upd, (size_t)p->U * sizeof(float), cudaMemcpyDeviceToHost, p->stream));
  CK(cudaStreamSynchronize(p->stream));
  for (int u = 0; u < p->U; u++) { out[3 * u + 1] = p->streaming ? p->s_tu[u] : frames_of(p->lengths[u], p->D, p->laN, p->pdA); out[3 * u + 2] = upd[u]; }
  return BTKB_OK;
}

static int get_rows(btkb_pipeline* p, const float2* d_src, int rows, float* out) {
  std::vector<float2> tmp((size_t)rows * p->Gp);
  CK(cudaMemcpyAsync(tmp.data(), d_src, tmp.size() * sizeof(float2), cudaMemcpyDeviceToHost, p->stream));
  CK(cudaStreamSynchronize(p->stream));
  from_device_layout(tmp, out, p->wU ? p->wU : p->U, p->K, rows, p->Gp);
  return BTKB_OK;
}

int btkb_get_weights(btkb_pipeline* p, float* out) {
  if (!p || !out) return fail(BTKB_ERR_INVALID, "null argument");
  if (!p->have_w) return fail(BTKB_ERR_STATE, "btkb_get_weights: no weights");
  CK(cudaSetDevice(p->cfg.device));
  return get_rows(p, p->d_W, p->C, out);
}

int btkb_get_active_weights(btkb_pipeline* p, float* out) {
  if (!p || !out) return fail(BTKB_ERR_INVALID, "null argument");
  CK(cudaSetDevice(p->cfg.device));
  if (p->cfg.beamformer == BTKB_BF_GSC_LMS || p->cfg.beamformer == BTKB_BF_GSC_RLS) {
    if (!p->have_ua) return fail(BTKB_ERR_STATE, "btkb_get_active_weights: the adaptive sidelobe canceller has not run");
    if (p->C > 8) return fail(BTKB_ERR_INVALID, "btkb_get_active_weights: the blocking-matrix export is built for <= 8 channels");
    CK(launch_ua_to_wa(p->d_UA, p->d_TA, p->d_WA, p->U, p->C, p->K, p->Gp, p->stream));
  } else if (!p->have_wl) {
    return fail(BTKB_ERR_STATE, "btkb_get_active_weights: no active weights set");
  }
  return get_rows(p, p->d_WA, p->C - p->NC, out);
}

int btkb_get_sidelobe_weights(btkb_pipeline* p, float* out) {
  if (!p || !out) return fail(BTKB_ERR_INVALID, "null argument");
  if (!p->have_w) return fail(BTKB_ERR_STATE, "btkb_get_sidelobe_weights: no weights");
  if (p->cfg.beamformer == BTKB_BF_GSC_LMS || p->cfg.beamformer == BTKB_BF_GSC_RLS)
    return fail(BTKB_ERR_INVALID, "btkb_get_sidelobe_weights: the adaptive kernels carry u = wa^H B^T, not wl; read btkb_get_active_weights");
  CK(cudaSetDevice(p->cfg.device));
  if (!p->have_wl) {  // zero_active_weights: wl = B 0 (beamformer.cc:1387-1399)
    memset(out, 0, (size_t)(p->wU ? p->wU : p->U) * p->K * p->C * sizeof(float2));
    return BTKB_OK;
  }
  return get_rows(p, p->d_WL, p->C, out);
}

int btkb_get_covariance(btkb_pipeline* p, float* out) {
  if (!p || !out) return fail(BTKB_ERR_INVALID, "null argument");
  if (!p->have_R) return fail(BTKB_ERR_STATE, "btkb_get_covariance: no covariance");
  CK(cudaSetDevice(p->cfg.device));
  int rc = get_rows(p, p->d_R, p->C * p->C, out); if (rc) return rc;
  if (p->R_is_sum) {  // finalize_stats (pybeamformer.py:994-1000): divide by the number of noise frames
    std::vector<int> cnt(p->U);
    CK(cudaMemcpy(cnt.data(), p->d_count, (size_t)p->U * sizeof(int), cudaMemcpyDeviceToHost));
    const size_t per = (size_t)p->K * p->C * p->C * 2;
    for (int u = 0; u < p->U; u++)
      if (cnt[u] > 0) for (size_t i = 0; i < per; i++) out[(size_t)u * per + i] /= (float)cnt[u];
  }
  return BTKB_OK;
}

int btkb_device_pointers(btkb_pipeline* p, void** X, void** Y, void** time_out) {
  if (!p) return fail(BTKB_ERR_INVALID, "null pipeline");
  if (X) *X = p->d_X; if (Y) *Y = p->Y_out; if (time_out) *time_out = p->d_time;
  return BTKB_OK;
}

// ------------------------------------------------------------------------------------- streamed chunks with carried state
// The reference's streams are pulled one frame at a time (FeatureStream::next, stream/stream.h:16-54): an utterance never has to be
// complete before its first frame comes out, and the look direction may change while it runs (unit_test/test_online_beamforming.py:
// 205-225).  btkb_stream_begin / btkb_stream_submit give the batch pipe the same property at chunk granularity.  The contract is
// pinned on the CPU in the reference's own fp64 arithmetic (tests/test_chunked_spec.py); here:
//   K1  sees the chunk's samples behind a history prefix of min(m M - D, samples so far) samples and numbers its frames from t_base;
//   K4  loads / stores the per-chain recurrence state (ST, UA) and tests absolute frame numbers;
//   K5  sees the chunk's Y rows behind m R - 1 rows of history and numbers its blocks from b_base.
// A chunked run equals the whole-utterance run bit for bit (tests/test_parity_gpu_r2.py::test_streamed_chunks_*).
static int stream_unsupported(btkb_pipeline* p) {
  if (p->cfg.wpe.enabled) return fail(BTKB_ERR_INVALID, "btkb_stream_begin: WPE buffers the whole utterance by definition (dereverberation.cc:500-534); submit whole utterances");
  if (p->C > 8) return fail(BTKB_ERR_INVALID, "btkb_stream_begin: streamed chunks are built for the register-path kernels (<= 8 channels)");
  if (p->cfg.beamformer == BTKB_BF_GSC_RLS_CPP) return fail(BTKB_ERR_INVALID, "btkb_stream_begin: the fp64 SubbandGSCRLS parity kernel processes whole utterances");
  return BTKB_OK;
}

int btkb_stream_begin(btkb_pipeline* p, int U) {
  if (!p) return fail(BTKB_ERR_INVALID, "btkb_stream_begin: null pipeline");
  if (U < 1 || U > p->Ucap) return fail(BTKB_ERR_INVALID, "btkb_stream_begin: U out of range");
  int rc = stream_unsupported(p); if (rc) return rc;
  CK(cudaSetDevice(p->cfg.device));
  const int Ha = p->m * p->M - p->D;
  if (!p->d_xs[0]) {
    p->xs_stride = round_up(Ha + p->ncap + p->D, 4);
    for (int i = 0; i < 2; i++) CK(cudaMalloc((void**)&p->d_xs[i], (size_t)p->Ucap * p->C * p->xs_stride * sizeof(float)));
    CK(cudaMalloc((void**)&p->d_ST, (size_t)ST_ROWS * p->Gpcap * sizeof(float)));
    CK(cudaMalloc((void**)&p->d_tu, (size_t)p->Ucap * sizeof(int)));
  }
  if (p->wU != U) { p->have_w = p->have_wl = p->have_ta = p->have_R = false; p->R_is_sum = false; p->NC = 1; p->wU = 0; }
  p->U = U; p->Gp = round_up(U * p->K, 128); p->n = 0; p->T = 0; p->nb = 0;
  p->have_X = p->have_Y = p->have_time = p->have_ua = false;
  p->streaming = true; p->stream_final = false; p->xs_cur = 0; p->xs_hist = 0; p->s_samples = 0;
  p->s_tnext = p->s_tlast = p->s_bnext = p->s_blast = p->s_chunks = 0;
  p->s_len.assign(U, 0); p->s_tu.assign(U, 0); p->lengths.assign(U, 0);
  p->Y_out = p->d_Y + (size_t)p->Hy * p->Gp;
  CK(cudaMemsetAsync(p->d_stats, 0, (size_t)U * 3 * sizeof(double), p->stream));
  CK(cudaMemsetAsync(p->d_Y, 0, (size_t)p->Hy * p->Gp * sizeof(float2), p->stream));
  return BTKB_OK;
}

__global__ void k_i16_to_f32_rows(const int16_t* __restrict__ src, float* __restrict__ dst, size_t rows, int n, int dst_stride, int dst_off) {
  const size_t total = rows * (size_t)n;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const size_t r = i / n; const int c = (int)(i - r * n);
    dst[r * dst_stride + dst_off + c] = (float)src[i];
  }
}

static int stream_submit_impl(btkb_pipeline* p, const float* samples, const int16_t* samples16, int n, const int* lengths, int final_chunk, int do_syn) {
  if (!p || (!samples && !samples16 && n > 0)) return fail(BTKB_ERR_INVALID, "btkb_stream_submit: null argument");
  if (!p->streaming) return fail(BTKB_ERR_STATE, "btkb_stream_submit: call btkb_stream_begin first");
  if (p->stream_final) return fail(BTKB_ERR_STATE, "btkb_stream_submit: the stream has ended (final chunk already submitted); call btkb_stream_begin");
  if (n < 0 || n > p->ncap) return fail(BTKB_ERR_INVALID, "btkb_stream_submit: n exceeds max_samples");
  if (!final_chunk && (n % p->D != 0 || n == 0)) return fail(BTKB_ERR_INVALID, "btkb_stream_submit: a non-final chunk must hold a positive whole number of D-sample blocks");
  if (!final_chunk && lengths) return fail(BTKB_ERR_INVALID, "btkb_stream_submit: ragged lengths only on the final chunk");
  if (!p->have_h) return fail(BTKB_ERR_STATE, "btkb_run: set the analysis prototype first");
  if (!p->have_w) {
    if (p->cfg.beamformer == BTKB_BF_MVDR) return fail(BTKB_ERR_STATE, "call calc_mvdr_weights() once");
    return fail(BTKB_ERR_STATE, "call calc_array_manifold_vectorsX() once");
  }
  if (p->wU != p->U) return fail(BTKB_ERR_INVALID, "btkb_stream_submit: weights were set for a different number of utterances");
  if (do_syn && !p->have_g) return fail(BTKB_ERR_STATE, "btkb_run: set the synthesis prototype first");
  const int U = p->U, C = p->C, D = p->D;
  std::vector<int> lloc(U);
  int Tabs = 0;
  for (int u = 0; u < U; u++) {
    const int ln = lengths ? lengths[u] : n;
    if (ln < 0 || ln > n) return fail(BTKB_ERR_INVALID, "btkb_stream_submit: lengths[u] out of range");
    lloc[u] = ln;
  }
  CK(cudaSetDevice(p->cfg.device));
  p->launches = 0;
  // ---- samples: [history | new] in the current buffer; `lengths` of the row count the history
  float* xs = p->d_xs[p->xs_cur];
  if (n > 0 && samples)
    CK(cudaMemcpy2DAsync(xs + p->xs_hist, (size_t)p->xs_stride * sizeof(float), samples, (size_t)n * sizeof(float), (size_t)n * sizeof(float), (size_t)U * C,
                         cudaMemcpyHostToDevice, p->stream));
  else if (n > 0) {   // 16-bit PCM chunk (live capture): upload as is, widen on the device behind the history prefix
    if (!p->d_x16) CK(cudaMalloc((void**)&p->d_x16, (size_t)p->Ucap * C * p->n_stride * sizeof(int16_t)));
    CK(cudaMemcpyAsync(p->d_x16, samples16, (size_t)U * C * n * sizeof(int16_t), cudaMemcpyHostToDevice, p->stream));
    k_i16_to_f32_rows<<<148 * 4, 256, 0, p->stream>>>(p->d_x16, xs, (size_t)U * C, n, p->xs_stride, p->xs_hist);
    CK(cudaGetLastError());
  }
  const long long s_base = p->s_samples - p->xs_hist;   // absolute index of the row's first sample
  for (int u = 0; u < U; u++) {
    p->s_len[u] += lloc[u];
    p->lengths[u] = p->xs_hist + lloc[u];
    // frames so far: a running stream has emitted every frame whose window is complete (blocks - laN); the end of the stream adds the
    // pd_A flush frames of zero blocks (modulated.cc:440-466)
    const long long blocks = final_chunk ? (p->s_len[u] + D - 1) / D : (p->s_samples + n) / D;
    p->s_tu[u] = (int)std::max<long long>(final_chunk ? blocks - p->laN + p->pdA : blocks - p->laN, 0);
    Tabs = std::max(Tabs, p->s_tu[u]);
  }
  const int t_base = p->s_tnext, b_base = p->s_bnext;
  const int Tloc = std::max(Tabs - t_base, 0);
  if (Tloc > p->Tcap) return fail(BTKB_ERR_INVALID, "btkb_stream_submit: the chunk yields more frames than the pipeline was sized for");
  // Two consecutive frames ride one synthesis transform and rounding makes a frame's result depend on its partner; tiles start at
  // even absolute blocks, so frame tau is the FIRST of its pair when tau - (pd_S - (m R - 1)) is even.  If the last frame available is
  // such a frame its partner has not arrived yet: the block that needs it is held back until the next chunk (the end of the
  // stream pairs it with zero, like the whole-utterance run).
  const int q0 = (((p->pdS - (p->m * p->R - 1)) % 2) + 2) % 2;
  const int hold = (!final_chunk && Tabs > 0 && (((Tabs - 1 - q0) % 2 + 2) % 2) == 0) ? 1 : 0;
  const int nb_abs = std::max(Tabs - p->pdS - hold, 0);
  const int nbloc = std::max(nb_abs - b_base, 0);
  p->T = Tloc; p->nb = nbloc; p->n = n;
  CK(cudaMemcpyAsync(p->d_len, p->lengths.data(), U * sizeof(int), cudaMemcpyHostToDevice, p->stream));
  CK(cudaMemcpyAsync(p->d_tu, p->s_tu.data(), U * sizeof(int), cudaMemcpyHostToDevice, p->stream));
  CK(cudaEventRecord(p->ev[0], p->stream));
  if (Tloc > 0) {
    AnalysisArgs a{xs, p->d_len, p->d_h, p->d_X, p->d_E, U, C, p->xs_hist + n, p->xs_stride, Tloc, p->M, p->m, D, p->laN, p->Gp, 1, p->d_tw, 1, 0,
                   (long long)(p->laN + t_base + 1) * D - s_base, t_base & 1, nullptr, 0, p->Cp};
    CK(launch_analysis(a, p->stream));
    p->launches++;
  }
  p->have_X = Tloc > 0;
  CK(cudaEventRecord(p->ev[1], p->stream));
  if (Tloc > 0) {
    PerBinArgs a = perbin_args(p);
    a.t_base = t_base; a.tu = p->d_tu; a.ST = p->d_ST; a.st_load = p->s_chunks > 0 ? 1 : 0;
    if (p->cfg.postfilter >= BTKB_PF_MCCOWAN) {
      if (!p->have_pfR) return fail(BTKB_ERR_STATE, "McCowanPostFilter:  construct/set a noise coherence matrix");
      const bool lef = p->cfg.postfilter == BTKB_PF_LEFKIMMIATIS;
      CK(launch_pf_prepare(p->d_pfR, p->d_pfInvR, p->d_pfQ, C, p->K, p->cfg.pf_threshold, p->cfg.pf_min_sv, lef ? 1 : 0, p->stream));
      p->launches++;
      if (lef) { CK(launch_pf_lambda(p->d_pfInvR, p->have_ta ? p->d_TA : p->d_W, p->d_LAM, U, C, p->K, p->Gp, p->cfg.pf_type, p->stream)); p->launches++; }
      a.PFQ = p->d_pfQ; a.LAM = p->d_LAM; a.pf_fbin1 = p->cfg.pf_fbin1;
    }
    CK(launch_perbin(a, p->stream));
    p->launches++;
    p->s_chunks++;
  }
  p->have_Y = Tloc > 0; p->pf_applied = true;
  p->have_ua = (p->cfg.beamformer == BTKB_BF_GSC_LMS || p->cfg.beamformer == BTKB_BF_GSC_RLS) && p->s_chunks > 0;
  CK(cudaEventRecord(p->ev[2], p->stream));
  p->have_time = false;
  if (do_syn && nbloc > 0) {
    SynthesisArgs a{p->d_Y, p->d_len, p->d_g, p->d_time, p->d_stats, U, n, Tloc, p->M, p->m, p->cfg.r, D, p->K, p->Gp, p->pdS, p->laN, p->pdA,
                    nbloc, nbloc * D, p->cfg.synthesis_gain, p->d_tw,
                    (p->cfg.postfilter >= BTKB_PF_MCCOWAN) ? p->cfg.pf_min_frames + 1 : 0, b_base, t_base - p->Hy, p->d_tu,
                    b_base & 1};
    CK(launch_synthesis(a, p->stream));
    p->launches++;
    p->have_time = true;
  }
  CK(cudaEventRecord(p->ev[3], p->stream));
  // ---- carry the histories over: the last min(m M - D, samples so far) samples become the head of the other sample buffer, the last
  // m R - 1 rows of Y move to the front (through the scratch buffer when the chunk is shorter than the history)
  const int Ha = p->m * p->M - D;
  const int have = p->xs_hist + n;
  const int keep = std::min(Ha, have);
  if (!final_chunk) {
    float* nx = p->d_xs[p->xs_cur ^ 1];
    if (keep > 0)
      CK(cudaMemcpy2DAsync(nx, (size_t)p->xs_stride * sizeof(float), xs + (have - keep), (size_t)p->xs_stride * sizeof(float), (size_t)keep * sizeof(float),
                           (size_t)U * C, cudaMemcpyDeviceToDevice, p->stream));
    p->xs_cur ^= 1; p->xs_hist = keep;
    if (Tloc > 0 && p->Hy > 0) {
      const size_t row = (size_t)p->Gp * sizeof(float2);
      if (Tloc >= p->Hy) CK(cudaMemcpyAsync(p->d_Y, p->d_Y + (size_t)Tloc * p->Gp, p->Hy * row, cudaMemcpyDeviceToDevice, p->stream));
      else {
        int rc = ensure_scratch(p, p->Hy * row); if (rc) return rc;
        CK(cudaMemcpyAsync(p->d_scratch, p->d_Y + (size_t)Tloc * p->Gp, p->Hy * row, cudaMemcpyDeviceToDevice, p->stream));
        CK(cudaMemcpyAsync(p->d_Y, p->d_scratch, p->Hy * row, cudaMemcpyDeviceToDevice, p->stream));
      }
    }
  }
  p->s_samples += n;
  p->s_tlast = t_base; p->s_blast = b_base; p->s_tnext = Tabs; p->s_bnext = std::max(nb_abs, b_base);
  p->stream_final = final_chunk != 0;
  return BTKB_OK;
}

int btkb_stream_submit(btkb_pipeline* p, const float* samples, int n, const int* lengths, int final_chunk, int do_syn) {
  return stream_submit_impl(p, samples, nullptr, n, lengths, final_chunk, do_syn);
}
int btkb_stream_submit_i16(btkb_pipeline* p, const int16_t* samples, int n, const int* lengths, int final_chunk, int do_syn) {
  return stream_submit_impl(p, nullptr, samples, n, lengths, final_chunk, do_syn);
}

int btkb_reset(btkb_pipeline* p) {   // FeatureStream::reset() of every stream of the graph (stream/stream.h:41): rewind, forget the adaptive state
  if (!p) return fail(BTKB_ERR_INVALID, "btkb_reset: null pipeline");
  CK(cudaSetDevice(p->cfg.device));
  CK(cudaStreamSynchronize(p->stream));
  if (p->streaming) return btkb_stream_begin(p, p->U);
  p->have_X = p->have_Y = p->have_time = p->have_ua = false;
  p->U = 0; p->T = 0; p->nb = 0;
  return BTKB_OK;
}

int btkb_set_stream(btkb_pipeline* p, void* cuda_stream) {
  if (!p) return fail(BTKB_ERR_INVALID, "btkb_set_stream: null pipeline");
  CK(cudaSetDevice(p->cfg.device));
  CK(cudaStreamSynchronize(p->stream));
  if (p->owns_stream && p->stream) cudaStreamDestroy(p->stream);
  if (cuda_stream) { p->stream = (cudaStream_t)cuda_stream; p->owns_stream = false; }
  else { CK(cudaStreamCreateWithFlags(&p->stream, cudaStreamNonBlocking)); p->owns_stream = true; }
  return BTKB_OK;
}

int btkb_stream_position(const btkb_pipeline* p, int* first_frame, int* first_block) {
  if (!p || !p->streaming) return fail(BTKB_ERR_STATE, "btkb_stream_position: not streaming");
  if (first_frame) *first_frame = p->s_tlast;
  if (first_block) *first_block = p->s_blast;
  return BTKB_OK;
}

}  // extern "C"
