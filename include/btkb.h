/*
 * btkb.h — C-ABI of the B200-native subband beamforming pipe ("btk batch").
 *
 * This is the drop-in boundary for btk2.0's hot path: everything the reference computes per frame inside
 *   OverSampledDFTAnalysisBank::next        (btk20_src/modulated/modulated.cc:375-409)
 *   SnapShotArray::update                   (btk20_src/beamformer/beamformer.cc:56-70)
 *   SubbandDS/GSC/MVDR/MVDRGSC::next        (beamformer.cc:1095-1157, 1251-1316, 2537-2587, 2719-2773)
 *   SubbandGSCLMSBeamformer.__iter__        (btk20_src/lib/pybeamformer.py:659-734)
 *   SubbandSMIMVDRBeamformer.accu_stats...  (pybeamformer.py:948-1023)  +  SubbandMVDR::calc_mvdr_weights (beamformer.cc:2350-2402)
 *   ZelinskiPostFilter::next                (btk20_src/postfilter/postfilter.cc:424-491)
 *   OverSampledDFTSynthesisBank::next       (modulated.cc:569-612)
 * is executed here for a whole batch of utterances by hand-written sm_100a CUDA kernels.
 *
 * Conventions: plain C types only (no torch / C++ types); every function returns BTKB_OK (0) or a negative
 * error code and never throws; btkb_last_error() returns a thread-local description of the last failure.  One host
 * thread per pipeline handle.  All work is stream-ordered on the pipeline's CUDA stream; functions that hand data to
 * the host synchronise that stream before returning.  There is NO CPU fallback: creation fails with
 * BTKB_ERR_NO_DEVICE when no CUDA device is usable.
 *
 * Shapes: U utterances, C channels, M subbands (fft_len), m prototype-length factor, r decimation exponent,
 * D = M >> r samples per frame, K = M/2 + 1 unique bins, T frames:
 *   T(n) = ceil(n / D) - laN + pd_A   (modulated.cc:246-264, 418-469; = ceil(n/D) + m 2^r / 2 for delay-comp. type 2).
 * Complex arrays are interleaved (re, im).
 */
#ifndef BTKB_H
#define BTKB_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BTKB_OK 0
#define BTKB_ERR_INVALID (-1)    /* bad argument / configuration (the reference throws jdimension_error / jconsistency_error) */
#define BTKB_ERR_NO_DEVICE (-2)  /* no usable CUDA device: the product never falls back to the CPU */
#define BTKB_ERR_CUDA (-3)       /* a CUDA runtime call failed; see btkb_last_error() */
#define BTKB_ERR_STATE (-4)      /* call order violated (the reference throws j_error "call calc_..._weights() once") */
#define BTKB_ERR_ALLOC (-5)      /* jallocation_error */

/* beamformer kinds */
#define BTKB_BF_DS 0       /* SubbandDS: y = wq^H x                                          (beamformer.cc:1095-1157) */
#define BTKB_BF_GSC 1      /* SubbandGSC: DC bin wq^H x, bins >= 1 (wq - wl)^H x, static wl  (beamformer.cc:1208-1316) */
#define BTKB_BF_MVDR 2     /* SubbandMVDR / SubbandMVDRGSC: y = (wmvdr - wl)^H x             (beamformer.cc:2537-2587, 2719-2773) */
#define BTKB_BF_GSC_LMS 3  /* SubbandGSCLMSBeamformer: leaky power-normalised NLMS           (pybeamformer.py:588-762) */
#define BTKB_BF_GSC_RLS 4  /* SubbandGSCRLSBeamformer: regularised RLS sidelobe canceller    (pybeamformer.py:765-928) */
#define BTKB_BF_GSC_RLS_CPP 5 /* the C++ class SubbandGSCRLS: RLS in the blocking-matrix basis, fp64 (beamformer.cc:1447-1699) */

/* post-filter kinds */
#define BTKB_PF_NONE 0
#define BTKB_PF_ZELINSKI 1 /* ZelinskiPostFilter (postfilter.cc:57-219, 424-491); type bits: 1 = Re, 2 = |.| */
#define BTKB_PF_MCCOWAN 2  /* McCowanPostFilter: noise-coherence-compensated Wiener gain (postfilter.cc:496-934) */
#define BTKB_PF_LEFKIMMIATIS 3 /* LefkimmiatisPostFilter: McCowan's clean PSD + noise PSD / (d^H R^-1 d) (postfilter.cc:935-1200) */

typedef struct btkb_pipeline btkb_pipeline;

typedef struct btkb_lms_params { /* defaults = unit_test/confs/gsclms.json / pybeamformer.py:597-607 */
  float beta, gamma, init_diagonal_load, regularization_param, energy_floor, sil_thresh, max_wa_l2norm;
  int min_frames, slowdown_after;
} btkb_lms_params;

typedef struct btkb_rls_params { /* defaults = unit_test/confs/gscrls.json / pybeamformer.py:773-786 */
  float beta, gamma, mu, init_diagonal_load, regularization_param, sil_thresh, alpha2, max_wa_l2norm;
  int constraint_option;  /* 0 none, 1 quadratic constraint, 2 norm normalisation, 3 both */
  int min_frames;
} btkb_rls_params;

typedef struct btkb_rls_cpp_params { /* SubbandGSCRLS(fftlen, half_band_shift, myu = 0.9, sigma2 = 0.01) (beamformer.i:306-330) */
  float mu;           /* forgetting factor myu */
  float sigma2;       /* diagonal_weights_[f]: the leak of the weight update (constructor argument sigma2) */
  float init_sigma2;  /* init_precision_matrix(sigma2 = 0.01): Pz = I / sigma2 */
  float alpha;        /* set_quadratic_constraint(alpha, qctype) */
  int qctype;         /* 0 none, 1 CONSTANT_NORM, 2 THRESHOLD_LIMITATION (beamformer.h) */
  int update;         /* update_active_weight_vecotrs(flag), default 1 */
} btkb_rls_cpp_params;

typedef struct btkb_wpe_params { /* MultiChannelWPEDereverberation ctor (dereverberation.cc:312-334); defaults = unit_test/confs/wpe.json */
  int enabled;         /* 1: btkb_run() dereverberates the snapshots between the analysis bank and the beamformer */
  int lower_num, upper_num, iterations_num;
  double load_db, band_width, diagonal_bias;
  int fp32_normal_equations; /* 0 (default): R_c, r_c and the Cholesky solve in fp64 like the reference; 1: fp32 (faster, for well-loaded problems) */
} btkb_wpe_params;

typedef struct btkb_config {
  int device;                  /* CUDA device ordinal */
  int channels;                /* C >= 1 (GSC kinds need C >= 2) */
  int fft_len;                 /* M, power of two in [256, 2048] */
  int m;                       /* prototype length factor (prototype length = m*M) */
  int r;                       /* decimation exponent, D = M >> r */
  int delay_compensation_type; /* 0, 1 or 2 (modulated.cc:246-264) */
  float samplerate;
  int beamformer;              /* BTKB_BF_* */
  int postfilter;              /* BTKB_PF_* */
  float pf_alpha;              /* Zelinski forgetting factor (default 0.6) */
  int pf_type;                 /* 1 or 2 (postfilter.h:41-47) */
  int pf_min_frames;
  btkb_lms_params lms;
  int max_utterances;          /* capacity of one batch */
  int max_samples;             /* capacity: samples per channel per utterance */
  int keep_snapshots;          /* 1: keep the analysis output X resident so btkb_fetch_snapshots works (always true today) */
  int synthesis_gain;          /* OverSampledDFTSynthesisBank gain_factor (modulated.cc:608-609); default 1 */
  int normalize_weight;        /* SubbandGSC::normalize_weight(flag): w <- w / (||w|| C) for bins >= 1 (beamformer.cc:1230-1236) */
  float pf_threshold;          /* McCowan / Lefkimmiatis: clip of the noise coherence R_ij (default 0.99, postfilter.h) */
  double pf_min_sv;            /* Lefkimmiatis: singular-value floor of the coherence pseudo-inverse (default 1e-8) */
  int pf_fbin1;                /* Lefkimmiatis: first bin that divides the noise PSD by Lambda = d^H R^-1 d (default 0) */
  btkb_rls_params rls;         /* BTKB_BF_GSC_RLS */
  btkb_wpe_params wpe;         /* multi-channel WPE dereverberation of the snapshots (C = 1, 2, 4 or 8) */
  btkb_rls_cpp_params rls_cpp; /* BTKB_BF_GSC_RLS_CPP */
} btkb_config;

/* ---- lifecycle ---------------------------------------------------------------------------------------------- */
void btkb_default_config(btkb_config* cfg);
int btkb_create(const btkb_config* cfg, btkb_pipeline** out);
void btkb_destroy(btkb_pipeline* p);
const char* btkb_last_error(void);
int btkb_device_count(void);

/* ---- setup (replaces OverSampledDFTFilterBank ctor, modulated.cc:232-268, and BeamformerWeights, beamformer.cc:485-965) */
/* analysis prototype h and synthesis prototype g, len = m*M doubles each (g may be NULL when synthesis is not used) */
int btkb_set_prototypes(btkb_pipeline* p, const double* h, const double* g, int len);
/* per-utterance time delays [U][C] (seconds) -> quiescent weights wq = calcMainlobe (beamformer.cc:502-565), ta = wq */
int btkb_set_delays(btkb_pipeline* p, int U, const double* delays);
/* LCMV quiescent weights with NC >= 2 linear constraints: target delays [U][C] + NC-1 jammer delay vectors [U][NC-1][C]
 * (calc_gsc_weights_n / calc_array_manifold_vectors_n -> BeamformerWeights::calcMainlobeN, beamformer.cc:573-721, with
 * calc_null_beamformer_, beamformer.cc:299-363).  The time-alignment manifold stays the delay-and-sum one; active
 * weights then have C-NC entries per bin. */
int btkb_set_delays_lcmv(btkb_pipeline* p, int U, int NC, const double* delaysT, const double* delaysJ);
/* explicit quiescent / MVDR weights [U][K][C] complex64 (replaces setQuiescentVector / wmvdr_) */
int btkb_set_weights(btkb_pipeline* p, int U, const float* w);
/* active weights wa [U][K][C-1] complex64 -> wl = B wa with B = calc_blocking_matrix_(wq) (beamformer.cc:373-454, 729-767) */
int btkb_set_active_weights(btkb_pipeline* p, int U, const float* wa);
/* BTKB_BF_MVDR only: which vector the blocking matrix of the next btkb_set_active_weights is orthogonal to.
 * 0 (default): the delay-and-sum manifold — SubbandMVDRGSC::calc_blocking_matrix1 (beamformer.cc:2638-2643);
 * 1: the MVDR weights of the last btkb_calc_mvdr_weights — calc_blocking_matrix2 (beamformer.cc:2649-2672).
 * Like the reference's alloc_bfweight_, the call discards active weights set before it. */
int btkb_set_blocking_source(btkb_pipeline* p, int from_mvdr_weights);
/* SubbandMVDRGSC::upgrade_blocking_matrix (beamformer.cc:2674-2691): from now on the blocking matrix of the bins >= 1 is the one
 * orthogonal to wq - wl (wq = the vector calc_blocking_matrix1 / 2 chose, wl = B wa of the last btkb_set_active_weights); bin 0
 * keeps its matrix, wl keeps its value until active weights are set again.  btkb_set_blocking_source undoes it. */
int btkb_upgrade_blocking_matrix(btkb_pipeline* p);
/* SubbandMVDRGSC::blocking_matrix_output(outChanX) (beamformer.cc:2693-2716) for every frame of the batch:
 * out [U][T][K] complex64 = b_outChanX^H x, b_i = column i of the current blocking matrix (what SubbandOrthogonalizer streams) */
int btkb_blocking_matrix_output(btkb_pipeline* p, int outChanX, float* out);
/* noise covariance R [U][K][C][C] complex64, row-major (set_noise_spatial_spectral_matrix, beamformer.cc:2410-2433) */
int btkb_set_noise_covariance(btkb_pipeline* p, int U, const float* R);
/* diffuse-noise coherence from microphone positions [C][3] (mm) (set_diffuse_noise_model, beamformer.cc:2442-2509) */
int btkb_set_diffuse_noise_model(btkb_pipeline* p, int U, const double* mpos, float sspeed);
/* ---- noise coherence of the McCowan / Lefkimmiatis post-filters: one [K][C][C] matrix set per pipeline (postfilter.h R_) */
/* McCowanPostFilter::set_diffuse_noise_model(micPositions [C][3], sampleRate, sspeed) (postfilter.cc:562-627) */
int btkb_pf_set_diffuse_noise_model(btkb_pipeline* p, const double* mpos, double samplerate, double sspeed);
/* McCowanPostFilter::set_noise_spatial_spectral_matrix for every bin: R [K][C][C] complex128 row-major (postfilter.cc:541-560) */
int btkb_pf_set_noise_coherence(btkb_pipeline* p, const double* R);
int btkb_pf_get_noise_coherence(btkb_pipeline* p, double* R);  /* noise_spatial_spectral_matrix(fbinX), all bins */
/* set_all_diagonal_loading(diagonalWeight): ADDS (float)mu to every diagonal, cumulatively like the reference (postfilter.cc:629-642) */
int btkb_pf_set_diagonal_loading(btkb_pipeline* p, float mu);
/* divide_all_nondiagonal_elements(mu): off-diagonals /= (1 + mu) (postfilter.cc:662-680) */
int btkb_pf_divide_nondiagonal(btkb_pipeline* p, float mu);
/* R += mu I (set_all_diagonal_loading, beamformer.cc:2511-2523) then w = R^-H d / (C d^H R^-1 d), w[0] = 1
 * (calc_mvdr_weights, beamformer.cc:2350-2402).  Uses the covariance from btkb_set_noise_covariance,
 * btkb_set_diffuse_noise_model or btkb_accumulate_covariance. */
int btkb_calc_mvdr_weights(btkb_pipeline* p, float mu);
/* the same with the reference's dThreshold (beamformer.i:414-486, default 1e-8): a bin whose loaded matrix has a singular value below
 * it gets the identity instead of an inverse (pseudoinverse returns false, beamformer.cc:267-274, 2381-2383).  <= 8 channels; the wide
 * solver keeps its pivot test.  btkb_calc_mvdr_weights(p, mu) == btkb_calc_mvdr_weights_ex(p, mu, 1e-8f). */
int btkb_calc_mvdr_weights_ex(btkb_pipeline* p, float mu, float dthreshold);

/* ---- data path ------------------------------------------------------------------------------------------------ */
/* host samples float32 [U][C][n] (int16 scale, like SampleFeature, feature/feature.cc:605-649), lengths[U] <= n (NULL: all n).
 * Asynchronous H2D on the pipeline stream (pinned host memory recommended). */
int btkb_submit(btkb_pipeline* p, const float* samples, int U, int n, const int* lengths);
/* 16-bit PCM variant (what SampleFeature::read gets from a wav file before its float conversion, feature/feature.cc:265-305,
 * norm == 0): host samples int16 [U][C][n]; halves the H2D bytes, the int16 -> float32 conversion runs on the device. */
int btkb_submit_i16(btkb_pipeline* p, const int16_t* samples, int U, int n, const int* lengths);
/* same, but `samples` is a DEVICE pointer in the same layout (no copy) */
int btkb_submit_device(btkb_pipeline* p, const float* d_samples, int U, int n, const int* lengths);
/* analysis only: time -> snapshots X (OverSampledDFTAnalysisBank + SnapShotArray) */
int btkb_run_analysis(btkb_pipeline* p);
/* SMI pass 1 (pybeamformer.py:948-1000): R[u][k] = mean over noise frames (outside [start,end] s, energy > threshold) of x x^H.
 * labels [U][2] seconds (NULL: every frame is noise).  Requires btkb_run_analysis. */
int btkb_accumulate_covariance(btkb_pipeline* p, const double* labels, float energy_threshold);
/* SpectralMatrixArray::update over every resident frame (beamformer.cc:122-143): R <- mu R + (1-mu) x x^T with NO
 * conjugate when legacy_noconj != 0 (the reference's arithmetic), x x^H otherwise.
 * Result readable with btkb_get_covariance.  Requires btkb_run_analysis. */
int btkb_spectral_matrix_update(btkb_pipeline* p, float mu, int legacy_noconj);
/* the per-bin beamformer (+ post-filter) over the resident snapshots, then synthesis when do_synthesis != 0 */
/* MultiChannelWPEDereverberation::estimate_filter(start_frame_no, end_frame_no) (dereverberation.cc:405-431) followed by
 * calc_every_channel_output for every frame (:441-497): the resident snapshots X are replaced by the dereverberated ones
 * (read them with btkb_fetch_snapshots; the beamformer / synthesis stages then run on them).  end_frame_no < 0 = all frames.
 * A non-positive Cholesky pivot reports BTKB_ERR_INVALID with the reference's jnumeric_error text (:676-678). */
int btkb_run_wpe(btkb_pipeline* p, int start_frame_no, int end_frame_no);
/* MultiChannelWPEDereverberationFeature::next over a NEW batch with the filters of the last btkb_run_wpe (the reference estimates once,
 * then re-reads the audio and pulls the output stage, unit_test/test_subband_dereverberator.py:147-170): run the analysis on the
 * new samples, then this call replaces the resident snapshots by x - G^H lags.  Same number of utterances as the estimation batch. */
int btkb_apply_wpe(btkb_pipeline* p);
/* prediction filters Gn_ [U][K][C][L] complex64, L = C * (upper_num - lower_num + 1), channel-major then lag (zero outside the band) */
int btkb_get_wpe_filter(btkb_pipeline* p, float* out);
/* hand prediction filters estimated elsewhere (another pipeline's btkb_get_wpe_filter, same shape [U][K][C][C*P] complex64) to this
 * pipeline; btkb_apply_wpe then dereverberates with them: MultiChannelWPEDereverberationFeature streams feeding a beamformer apply the
 * filters of the earlier estimate_filter() call, whatever audio they were estimated on (dereverberation.cc:441-497, 713-728). */
int btkb_set_wpe_filter(btkb_pipeline* p, int U, const float* G);
int btkb_run_beamformer(btkb_pipeline* p, int do_synthesis);
/* inject beamformed subband frames from the host, Y [U][T][K] complex64 (the stream an arbitrary upstream
 * VectorComplexFeatureStream would deliver to OverSampledDFTSynthesisBank::next, modulated.cc:533-549), then
 * btkb_run_synthesis resynthesises them.  lengths are taken from T: every utterance is treated as T frames long. */
int btkb_set_subband(btkb_pipeline* p, int U, int T, const float* Y);
int btkb_run_synthesis(btkb_pipeline* p);
/* snapshots computed elsewhere, X [U][T][C][K] complex64, take the place of the analysis output (the per-bin kernels, the covariance pass
 * and the post-filters then run on them; the frame energies the NLMS / SOS gates read are zero).  What a ZelinskiPostFilter wired with
 * set_snapshot_array / set_array_manifold_vector instead of set_beamformer needs (postfilter/postfilter.cc:384-417, 424-491). */
int btkb_set_snapshots(btkb_pipeline* p, int U, int T, const float* X);
/* whole pipe: analysis -> beamformer (+post-filter) -> synthesis */
int btkb_run(btkb_pipeline* p, int do_synthesis);
int btkb_synchronize(btkb_pipeline* p);

/* ---- streamed chunks with carried state: the per-frame pull contract of FeatureStream::next (stream/stream.h:16-54) at chunk
 * granularity.  An utterance need not be complete before its first frames come out, its length is unbounded (max_samples bounds one
 * CHUNK), and btkb_set_delays* may be called between chunks to move the look direction while the adaptive state is kept
 * (unit_test/test_online_beamforming.py:205-225 recomputes the weights inside its frame loop).
 *   btkb_stream_begin(p, U)      rewind: frame counters, filter-bank histories and the adaptive state (NLMS / RLS weights, post-filter
 *                                statistics) start afresh for U utterances that advance in lockstep; weights already set for U are kept.
 *   btkb_stream_submit(p, samples [U][C][n] float32, n, lengths, final, do_synthesis)
 *                                appends n samples per utterance (a positive multiple of D unless `final`) and runs analysis ->
 *                                beamformer / post-filter (-> synthesis) for exactly the frames this chunk completes: frames
 *                                [blocks_before - laN, blocks_now - laN) of every utterance; the final chunk (lengths[u] <= n valid new
 *                                samples, NULL = n) adds the pd_A flush frames the reference emits after its source ends
 *                                (modulated.cc:440-466).  Afterwards btkb_num_frames / btkb_num_blocks / btkb_fetch_subband /
 *                                btkb_fetch_time / btkb_fetch_snapshots / btkb_get_postfilter_weights describe THIS chunk;
 *                                btkb_fetch_stats is cumulative.  The concatenated chunks equal the whole-utterance run bit for bit.
 *   btkb_stream_position         absolute number of the last chunk's first frame and first output block.
 * Not offered for WPE (it buffers the whole utterance by definition, dereverberation.cc:500-534), the batch statistics
 * (btkb_accumulate_covariance, btkb_sos_*) and > 8 channels. */
int btkb_stream_begin(btkb_pipeline* p, int U);
int btkb_stream_submit(btkb_pipeline* p, const float* samples, int n, const int* lengths, int final_chunk, int do_synthesis);
int btkb_stream_submit_i16(btkb_pipeline* p, const int16_t* samples, int n, const int* lengths, int final_chunk, int do_synthesis);  /* 16-bit PCM chunks (live capture, wav) */
int btkb_stream_position(const btkb_pipeline* p, int* first_frame, int* first_block);
/* FeatureStream::reset() of the whole graph (stream/stream.h:41-47): drops the resident batch; while streaming, same as
 * btkb_stream_begin with the same U.  Weights are kept (the reference's reset() does not touch BeamformerWeights). */
int btkb_reset(btkb_pipeline* p);
/* Run all of this pipeline's work on the caller's CUDA stream (a cudaStream_t passed as void*; NULL = back to a private stream).
 * The pipeline synchronises its previous stream first.  SURVEY.md 8(b): "stream-ordered on a caller-supplied cudaStream_t". */
int btkb_set_stream(btkb_pipeline* p, void* cuda_stream);

/* ---- results (all synchronise the stream) ------------------------------------------------------------------- */
int btkb_num_frames(const btkb_pipeline* p);        /* T of the longest utterance in the batch */
int btkb_num_frames_of(const btkb_pipeline* p, int u);
int btkb_num_blocks(const btkb_pipeline* p);        /* synthesis output blocks (T - pd_S), D samples each */
/* beamformed subband spectra [U][T][K] complex64 (bins above M/2 are conjugate mirrors, beamformer.cc:1142-1149) */
int btkb_fetch_subband(btkb_pipeline* p, float* out);
/* resynthesised signal float32 [U][blocks*D] */
int btkb_fetch_time(btkb_pipeline* p, float* out);
/* snapshots X [U][T][C][K] complex64 (the SnapShotArray contents per frame) */
int btkb_fetch_snapshots(btkb_pipeline* p, float* out);
/* per-utterance statistics [U][3] float64: sum of squared output samples, frames, NLMS update count
 * (test_online_beamforming.py:208,336-337; pybeamformer.py:751) */
int btkb_fetch_stats(btkb_pipeline* p, double* out);
int btkb_get_weights(btkb_pipeline* p, float* out);            /* [U][K][C] complex64: wq or wmvdr */
int btkb_get_active_weights(btkb_pipeline* p, float* out);     /* [U][K][C-1] complex64 (NLMS: waH of pybeamformer.py) */
/* [U][K][C] complex64: wl = B wa of the static sidelobe canceller (BeamformerWeights::wl_f, beamformer.cc:729-767); zeros while no
 * active weights are set.  The reference's write_fir_coeff (beamformer.cc:775-828) exports conj(wq - wl). */
int btkb_get_sidelobe_weights(btkb_pipeline* p, float* out);
int btkb_get_covariance(btkb_pipeline* p, float* out);         /* [U][K][C][C] complex64 */
int btkb_get_postfilter_weights(btkb_pipeline* p, float* out); /* [U][T][K] float32 post-filter gains (wp1_) */

/* ---- second-order-statistics batch beamformers: blind MVDR (MMSE) and GEV ------------------------------------------
 * SubbandBlindMVDRBeamformer / SubbandGEVBeamformer (lib/pybeamformer.py:1026-1357; driver unit_test/test_sos_batch_beamforming.py:
 * 186-233).  Use a BTKB_BF_DS pipeline with C = 2, 4 or 8: after btkb_run_analysis the statistics are accumulated per
 * (utterance, bin) on the device, btkb_sos_calc_weights writes the beamformer weights (y = w^H x for every bin incl. DC, the
 * reference's wqH = conj(w), pybeamformer.py:1191-1207) and btkb_run_beamformer applies them. */
#define BTKB_SOS_BMVDR 0   /* w = Rn^-1 Rt u_ref / (offset + tr(Rn^-1 Rt))                       (pybeamformer.py:1257-1295) */
#define BTKB_SOS_GEV 1     /* principal generalised eigenvector of (Rt, Rn), v^H Rn v = 1, phase-aligned bin to bin (:1311-1357).
                            * scipy.linalg.eigh leaves the eigenvector's phase to LAPACK; here (first Cholesky column of Rn)^H v is
                            * real POSITIVE at bin 0, so results equal the reference's up to one global sign per utterance. */
/* SubbandSOSBatchBeamformer.reset_stats (:1209-1213) */
int btkb_sos_reset_stats(btkb_pipeline* p);
/* accu_stats_from_label (:1063-1127): labels [U][NL][2] = NL (start, end) target segments per utterance in seconds, walked
 * with the reference's running elapsed time and segment cursor (an open end < 0 only works for a segment that starts at 0);
 * frames with channel-0 energy <= energy_threshold are skipped.  Adds to the statistics already accumulated. */
int btkb_sos_accumulate_from_label(btkb_pipeline* p, const double* labels, int NL, float energy_threshold);
/* accu_stats_from_tfmask (:1129-1183): masks float32 [U][Tm][K], Tm >= frames of every utterance (the reference raises
 * IndexError otherwise); values <= 0 are ignored; counts follow the reference's integer-array truncation. */
int btkb_sos_accumulate_from_tfmask(btkb_pipeline* p, const float* mask_t, const float* mask_j, int Tm, float energy_threshold);
/* finalize_stats(gamma) + calc_beamformer_weights(ref_micx, offset) for kind = BTKB_SOS_BMVDR / BTKB_SOS_GEV.  BTKB_ERR_STATE when
 * a bin has no target / noise statistics (the reference's assertion), BTKB_ERR_INVALID when a factorisation fails
 * ("Matrix inversion failed" / "GEV failed", :1286-1287, 1336-1337). */
int btkb_sos_calc_weights(btkb_pipeline* p, int kind, double gamma, int ref_micx, double offset);
/* raw statistics: Rt, Rn complex128 [U][K][C][C] sums (not normalised), counts float64 [U][K][2] (target, noise); NULLs are skipped */
int btkb_sos_get_stats(btkb_pipeline* p, double* Rt, double* Rn, double* counts);

/* ---- measurement -------------------------------------------------------------------------------------------- */
/* device time (ms, CUDA events on the pipeline stream) of the last run: total and per kernel
 * out[0] total, out[1] analysis, out[2] per-bin beamformer, out[3] synthesis, out[4] launches */
int btkb_last_timing(btkb_pipeline* p, float* out5);
/* device pointers for zero-copy consumers (torch tensors via from_dlpack / data_ptr): X, Y, time */
/* device time (ms) of the last WPE pass (estimation + output stage) */
int btkb_last_timing_wpe(btkb_pipeline* p, float* ms);
/* which normal equations the last estimation solved: 0 = lag-domain (L x L, L = C x lags, what estimate_Gn_ builds,
 * dereverberation.cc:553-690), 1 = frame-domain (S x S, S = estimation frames - lower; the same filters through the push-through
 * identity, chosen per batch when S < L).  The environment variable BTKB_WPE_FORM=lag|frame, read at create, pins one. */
int btkb_last_wpe_form(btkb_pipeline* p, int* form);
int btkb_device_pointers(btkb_pipeline* p, void** X, void** Y, void** time_out);

#ifdef __cplusplus
}
#endif
#endif /* BTKB_H */
