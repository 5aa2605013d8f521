// btk20_host.h — C++ host side above the C-ABI: btk2.0's stream classes for the hot path, same names and semantics,
// with the arithmetic delegated to libbtkb.so (include/btkb.h).  No torch, no GSL.
//
// Reference interfaces mirrored (btk20_src/):
//   FeatureStream<T>::{next,current,reset,size,is_end,frame_no,name}      stream/stream.h:16-54
//   exception hierarchy j_error ...                                       common/jexception.h:44-161
//   SampleFeature::{read,setSamples,data,samplesN,next,reset}             feature/feature.h:153-208, feature.cc:238-389,605-679
//   OverSampledDFTAnalysisBank / OverSampledDFTSynthesisBank              modulated/modulated.h:270-349
//   SnapShotArray                                                         beamformer/spectralinfoarray.h:6-39
//   SubbandDS / SubbandGSC / SubbandMVDR / SubbandMVDRGSC                 beamformer/beamformer.h:130-429
//   ZelinskiPostFilter                                                    postfilter/postfilter.h:74-110
//   SubbandGSCLMS: the native body of lib/pybeamformer.py:588-762 (SubbandGSCLMSBeamformer, pure Python in the reference)
//
// Execution model: the reference pulls ONE frame through the whole graph per next().  Here the first next() on a node
// collects its upstream graph, submits the whole utterance to the GPU pipeline once, and subsequent next() calls hand out
// frames from the fetched host buffers.  Frame numbering, idempotent re-reads (frame_no == frame_no_), end-of-stream
// (jiterator_error) and reset() follow the reference.
#pragma once
#include <complex>
#include <cstdarg>
#include <cstdio>
#include <exception>
#include <memory>
#include <string>
#include <vector>

struct btkb_pipeline;
struct btkb_config;

namespace btk20 {

// ----------------------------------------------------------------------------------------------------------------
class j_error : public std::exception {
 public:
  j_error() {}
  explicit j_error(const char* fmt, ...) { va_list ap; va_start(ap, fmt); set(fmt, ap); va_end(ap); }
  const char* what() const noexcept override { return msg_.c_str(); }
 protected:
  void set(const char* fmt, va_list ap) { char buf[1024]; vsnprintf(buf, sizeof(buf), fmt, ap); msg_ = buf; }
  std::string msg_;
};
#define BTK20_EXC(NAME)                                                                                   \
  class NAME : public j_error {                                                                           \
   public:                                                                                                \
    explicit NAME(const char* fmt, ...) { va_list ap; va_start(ap, fmt); set(fmt, ap); va_end(ap); }      \
  };
BTK20_EXC(jallocation_error) BTK20_EXC(jarithmetic_error) BTK20_EXC(jconsistency_error) BTK20_EXC(jdimension_error)
BTK20_EXC(jindex_error) BTK20_EXC(jinitialization_error) BTK20_EXC(jio_error) BTK20_EXC(jiterator_error) BTK20_EXC(jkey_error)
BTK20_EXC(jnumeric_error) BTK20_EXC(jparameter_error) BTK20_EXC(jparse_error) BTK20_EXC(jtype_error)
#undef BTK20_EXC

typedef std::complex<double> cplx;

// ----------------------------------------------------------------------------------------------------------------
template <class T>
class FeatureStream {
 public:
  virtual ~FeatureStream() {}
  const std::string& name() const { return name_; }
  unsigned size() const { return size_; }
  virtual const T* next(int frame_no = -5) = 0;
  const T* current() {
    if (frame_no_ < 0) throw jconsistency_error("Frame index (%d) < 0.", frame_no_);
    return next(frame_no_);
  }
  bool is_end() const { return is_end_; }
  virtual void reset() { frame_no_ = frame_reset_no_; is_end_ = false; }
  virtual int frame_no() const { return frame_no_; }

 protected:
  FeatureStream(unsigned sz, const std::string& nm) : frame_reset_no_(-1), size_(sz), frame_no_(-1), vector_(sz), is_end_(false), name_(nm) {}
  void increment_() { frame_no_++; }
  const int frame_reset_no_;
  const unsigned size_;
  int frame_no_;
  std::vector<T> vector_;
  bool is_end_;
 private:
  std::string name_;
};
typedef FeatureStream<float> VectorFloatFeatureStream;
typedef FeatureStream<cplx> VectorComplexFeatureStream;
typedef FeatureStream<double> VectorFeatureStream;        // stream/stream.h:72-95: the remaining element types of the handle classes
typedef FeatureStream<short> VectorShortFeatureStream;
typedef FeatureStream<char> VectorCharFeatureStream;
typedef std::shared_ptr<VectorFloatFeatureStream> VectorFloatFeatureStreamPtr;
typedef std::shared_ptr<VectorComplexFeatureStream> VectorComplexFeatureStreamPtr;

// ----------------------------------------------------------------------------------------------------------------
class SampleFeature : public VectorFloatFeatureStream {
 public:
  SampleFeature(const std::string& fn = "", unsigned block_len = 320, unsigned shift_len = 160, bool pad_zeros = false, const std::string& nm = "Sample");
  unsigned read(const std::string& fn, int format = 0, int samplerate = 16000, int chX = 1, int chN = 1, int cfrom = 0, int to = -1,
                int outsamplerate = -1, float norm = 0.0f);
  void set_samples(const double* samples, unsigned n, unsigned samplerate);
  const std::vector<float>& samples() const { return samples_; }
  unsigned samplesN() const { return (unsigned)samples_.size(); }
  int samplerate() const { return samplerate_; }
  unsigned shift_len() const { return shift_len_; }
  unsigned long version() const { return version_; }   // bumps whenever the sample array is replaced
  const float* next(int frame_no = -5) override;
  void reset() override { cur_ = 0; VectorFloatFeatureStream::reset(); }
 private:
  std::vector<float> samples_;
  unsigned shift_len_, cur_;
  bool pad_zeros_;
  int samplerate_;
  unsigned long version_;
};
typedef std::shared_ptr<SampleFeature> SampleFeaturePtr;

// ----------------------------------------------------------------------------------------------------------------
class OverSampledDFTAnalysisBank : public VectorComplexFeatureStream {
 public:
  OverSampledDFTAnalysisBank(const VectorFloatFeatureStreamPtr& samp, const std::vector<double>& prototype, unsigned M, unsigned m, unsigned r,
                             unsigned delay_compensation_type = 0, const std::string& nm = "OverSampledDFTAnalysisBank");
  ~OverSampledDFTAnalysisBank();
  const cplx* next(int frame_no = -5) override;
  void reset() override;
  unsigned fftlen() const { return M_; }
  unsigned shiftlen() const { return D_; }
  double polyphase(unsigned m, unsigned n) const { return prototype_.at(m + M_ * n); }
  // graph introspection for the batch engine
  const VectorFloatFeatureStreamPtr& source() const { return samp_; }
  const std::vector<double>& prototype() const { return prototype_; }
  unsigned M() const { return M_; } unsigned m() const { return m_; } unsigned r() const { return r_; } unsigned dct() const { return dct_; }
 private:
  void realize_();
  VectorFloatFeatureStreamPtr samp_;
  std::vector<double> prototype_;
  unsigned M_, m_, r_, D_, dct_;
  btkb_pipeline* pipe_;
  std::vector<std::complex<float>> X_;  // [T][K]
  int T_; bool realized_;
};
typedef std::shared_ptr<OverSampledDFTAnalysisBank> OverSampledDFTAnalysisBankPtr;

// ----------------------------------------------------------------------------------------------------------------
class SnapShotArray {
 public:
  SnapShotArray(unsigned fftlen, unsigned nchan) : fftLen_(fftlen), nChan_(nchan), samples_((size_t)fftlen * nchan), snapshots_((size_t)fftlen * nchan) {}
  const cplx* snapshot(unsigned fbinX) const { return &snapshots_[(size_t)fbinX * nChan_]; }
  void set_samples(const cplx* samp, unsigned chanX) { for (unsigned k = 0; k < fftLen_; k++) samples_[(size_t)chanX * fftLen_ + k] = samp[k]; }
  virtual ~SnapShotArray() {}
  virtual void update() { for (unsigned k = 0; k < fftLen_; k++) for (unsigned c = 0; c < nChan_; c++) snapshots_[(size_t)k * nChan_ + c] = samples_[(size_t)c * fftLen_ + k]; }
  virtual void zero() { std::fill(samples_.begin(), samples_.end(), cplx(0, 0)); std::fill(snapshots_.begin(), snapshots_.end(), cplx(0, 0)); }
  const cplx* samples(unsigned chanX) const { return &samples_[(size_t)chanX * fftLen_]; }
  unsigned fftLen() const { return fftLen_; }
  unsigned nChan() const { return nChan_; }
 private:
  unsigned fftLen_, nChan_;
  std::vector<cplx> samples_, snapshots_;
};
typedef std::shared_ptr<SnapShotArray> SnapShotArrayPtr;

// SpectralMatrixArray (beamformer/spectralinfoarray.h:43-63, beamformer.cc:95-143): the per-frame container of the recursively
// averaged spectral matrices, R <- mu R + (1 - mu) x x^T per bin (no conjugate, as the reference).  An interface container like
// SnapShotArray: one frame of C^2 values per bin is kept and updated here; the whole-utterance recursion on the GPU is
// btkb_spectral_matrix_update.  FBSpectralMatrixArray is not mirrored: its update() indexes the channel-major sample vectors by
// the bin number (beamformer.cc:163: samples_[ifft], an array of nChan vectors), i.e. it reads out of bounds for bins >= nChan.
class SpectralMatrixArray : public SnapShotArray {
 public:
  SpectralMatrixArray(unsigned fftlen, unsigned nchan, float forget_factor = 0.95f)
      : SnapShotArray(fftlen, nchan), mu_(forget_factor), matrices_((size_t)fftlen * nchan * nchan) {}
  const cplx* matrix_f(unsigned idx) const { if (idx >= fftLen()) throw jindex_error("index %d >= fftLen %d\n", (int)idx, (int)fftLen()); return &matrices_[(size_t)idx * nChan() * nChan()]; }
  void update() override;
  void zero() override { SnapShotArray::zero(); std::fill(matrices_.begin(), matrices_.end(), cplx(0, 0)); }
 protected:
  cplx mu_;
  std::vector<cplx> matrices_;
};
typedef std::shared_ptr<SpectralMatrixArray> SpectralMatrixArrayPtr;

// ----------------------------------------------------------------------------------------------------------------
// WPE dereverberation (dereverberation/dereverberation.h).  The reference estimates the filters from a buffered pass over the
// input streams (estimate_filter), resets them, and applies the filters frame by frame in next().  Here estimate_filter runs
// analysis + estimation + output stage for the current sample arrays on the GPU; when the sources have been re-read since
// (test_subband_dereverberator.py:147-150) the first next() re-runs analysis + output stage with the stored filters.
struct WpeConfig {
  unsigned lower_num = 0, upper_num = 32, iterations_num = 2;
  double load_db = -20.0, band_width = 0.0, diagonal_bias = 0.001, samplerate = 16000.0;
};
class MultiChannelWPEDereverberation {
 public:
  MultiChannelWPEDereverberation(unsigned subbandsN, unsigned channelsN, unsigned lowerN = 0, unsigned upperN = 32, unsigned iterationsN = 2, double loadDb = -20.0,
                                 double bandWidth = 0.0, double diagonalBias = 0.001, double sampleRate = 16000.0);
  ~MultiChannelWPEDereverberation();
  unsigned size() const { return subbandsN_; }
  unsigned channelsN() const { return channelsN_; }
  void set_input(const VectorComplexFeatureStreamPtr& samples);
  unsigned estimate_filter(int start_frame_no = 0, int end_frame_no = -1);
  void reset_filter();
  void next_speaker();
  void print_objective_func(int /*subband_no*/) {}
  void reset();
  bool estimated() const { return estimated_; }
  // engine hooks
  int frames();                                               // frames of the (re)realised output
  const std::complex<float>* output_frame(unsigned channelX, int t);  // [K] dereverberated bins of frame t
  const std::vector<VectorComplexFeatureStreamPtr>& sources() const { return sources_; }
  const WpeConfig& config() const { return cfg_; }
  int est_start() const { return est_start_; }
  int est_end() const { return est_end_; }
  std::vector<float> filters() const;                         // [K][C][C*P] complex64 prediction filters of the last estimate_filter()
 private:
  unsigned gather_(std::vector<float>& x);   // [C][n] samples of the current sources; returns n
  void realize_();
  unsigned subbandsN_, channelsN_;
  WpeConfig cfg_;
  std::vector<VectorComplexFeatureStreamPtr> sources_;
  btkb_pipeline* pipe_ = nullptr; unsigned cap_ = 0;
  std::vector<std::complex<float>> Xd_;      // [T][C][K]
  std::vector<unsigned long> versions_;
  int T_ = 0, est_start_ = 0, est_end_ = -1;
  bool estimated_ = false, realized_ = false;
};
typedef std::shared_ptr<MultiChannelWPEDereverberation> MultiChannelWPEDereverberationPtr;

class MultiChannelWPEDereverberationFeature : public VectorComplexFeatureStream {
 public:
  MultiChannelWPEDereverberationFeature(const MultiChannelWPEDereverberationPtr& source, unsigned channelX, unsigned primaryChannelX = 0,
                                        const std::string& nm = "MultiChannelWPEDereverberationFeature");
  const cplx* next(int frame_no = -5) override;
  void reset() override;
  const MultiChannelWPEDereverberationPtr& source() const { return source_; }
  unsigned channel() const { return channelX_; }
 private:
  MultiChannelWPEDereverberationPtr source_; unsigned channelX_, primaryChannelX_;
};
typedef std::shared_ptr<MultiChannelWPEDereverberationFeature> MultiChannelWPEDereverberationFeaturePtr;

// SingleChannelWPEDereverberationFeature (dereverberation.cc:24-310) = the multi-channel estimator with one channel and no
// diagonal bias (bit-identical in the reference itself, tests/test_oracle.py::test_wpe_single_channel_golden)
class SingleChannelWPEDereverberationFeature : public VectorComplexFeatureStream {
 public:
  SingleChannelWPEDereverberationFeature(const VectorComplexFeatureStreamPtr& samples, unsigned lowerN = 0, unsigned upperN = 64, unsigned iterationsN = 2,
                                         double loadDb = -20.0, double bandWidth = 0.0, double sampleRate = 16000.0,
                                         const std::string& nm = "SingleChannelWPEDereverberationFeature");
  unsigned estimate_filter(int start_frame_no = 0, int end_frame_no = -1) { return impl_->estimate_filter(start_frame_no, end_frame_no); }
  void print_objective_func(int) {}
  void reset_filter() { impl_->reset_filter(); }
  void next_speaker() { impl_->next_speaker(); VectorComplexFeatureStream::reset(); }
  const cplx* next(int frame_no = -5) override;
  void reset() override;
 private:
  MultiChannelWPEDereverberationPtr impl_;
};
typedef std::shared_ptr<SingleChannelWPEDereverberationFeature> SingleChannelWPEDereverberationFeaturePtr;

struct LmsConfig {  // lib/pybeamformer.py:597-607 defaults (= unit_test/confs/gsclms.json)
  double beta = 0.97, gamma = 0.01, init_diagonal_load = 1.0e6, regularization_param = 1.0e-4, energy_floor = 90, sil_thresh = 1.0e8, max_wa_l2norm = 100.0;
  int min_frames = 128, slowdown_after = 4096;
};
class McCowanPostFilter;
struct RlsConfig {  // lib/pybeamformer.py:773-786 defaults (= unit_test/confs/gscrls.json)
  double beta = 0.97, gamma = 0.04, mu = 0.97, init_diagonal_load = 1.0e6, regularization_param = 1.0e-2, sil_thresh = 1.0e8, alpha2 = 10.0,
         max_wa_l2norm = 100.0;
  int constraint_option = 3, min_frames = 128;
};
struct PostFilterConfig {
  bool enabled = false; double alpha = 0.6; int type = 2; int min_frames = 0;
  int kind = 1;                              // BTKB_PF_ZELINSKI / MCCOWAN / LEFKIMMIATIS
  float threshold = 0.99f; double min_sv = 1.0e-8; unsigned fbin1 = 0;
  const McCowanPostFilter* coherence = nullptr; unsigned long coherence_version = 0;  // owner of the noise coherence R_
};
struct SynthesisConfig { bool enabled = false; std::vector<double> prototype; unsigned M = 0, m = 0, r = 0, dct = 0; int gain = 1; };

// ----------------------------------------------------------------------------------------------------------------
// Common base of the subband beamformers (beamformer.h:89-128): channel list, snapshot array, batch realisation.
class SubbandBeamformer : public VectorComplexFeatureStream {
 public:
  SubbandBeamformer(unsigned fftLen, bool half_band_shift, int kind, const std::string& nm);
  ~SubbandBeamformer();
  void set_channel(const VectorComplexFeatureStreamPtr& chan);
  virtual void clear_channel();
  const cplx* next(int frame_no = -5) override;
  void reset() override;
  unsigned fftLen() const { return fftLen_; }
  unsigned chanN() const { return (unsigned)channels_.size(); }
  SnapShotArrayPtr snapshot_array();      // snapshots of the CURRENT frame (filled on demand)
  // weights (all per bin f = 0..M/2, [C] each)
  std::vector<cplx> get_weights(unsigned fbinX);
  // ---- batch engine hooks used by downstream nodes (post-filter, synthesis bank)
  void run_graph(const PostFilterConfig& pf, const SynthesisConfig& syn);
  const std::vector<std::complex<float>>& Y() const { return Y_; }
  const std::vector<float>& time_out() const { return time_; }
  const std::vector<float>& pf_weights() const { return pfw_; }
  int frames() const { return T_; }
  int blocks() const { return nb_; }
  bool realized_with(const PostFilterConfig& pf, const SynthesisConfig& syn) const;
  // ---- chunked realisation (host-mirror extension; the reference pulls ONE frame per next(), stream/stream.h:16-54).  With
  // chunk_blocks > 0 the graph is not run on the whole utterance at the first next(): the samples are handed to the GPU
  // chunk_blocks D-sample blocks at a time through btkb_stream_submit, and a chunk is submitted only when a frame (or output
  // block) beyond the ones already computed is asked for.  Device memory is bounded by the chunk, and beamformer weights
  // recomputed between two next() calls (a look-direction change, unit_test/test_online_beamforming.py:205-225) take effect
  // with the next chunk while the adaptive state is kept.  Default: environment variable BTK20_CHUNK_BLOCKS, else 0 (whole
  // utterance).  Not used for WPE chains and the batch-statistics beamformers.
  void set_chunk_blocks(int blocks) { chunk_blocks_ = blocks > 0 ? blocks : 0; invalidate_(); }
  int chunk_blocks() const { return chunk_blocks_; }
  void ensure_frames(int t);    // make frame t available (no-op for a whole-utterance realisation)
  void ensure_blocks(int b);    // make synthesis output block b available
  double samplerate_hint() const { return samplerate_; }

 protected:
  virtual void configure_weights_(btkb_pipeline* p) = 0;   // push delays / weights / covariance into a fresh pipeline
  virtual bool stream_capable_() const { return true; }
  virtual void tune_config_(btkb_config&) {}   // last word of a subclass on the pipeline configuration (kind-specific parameter blocks)   // false: the weights need statistics of the whole utterance on the device
  void invalidate_() { realized_ = false; live_ = false; }
  // new weights for the SAME graph: a live chunked stream keeps running and picks them up with its next chunk
  void invalidate_weights_() { if (live_ && realized_) weights_dirty_ = true; else invalidate_(); }
  void require_weights_(bool ok, const char* msg) const { if (!ok) throw j_error("%s", msg); }
  unsigned fftLen_; bool halfBandShift_; int kind_;
  std::vector<VectorComplexFeatureStreamPtr> channels_;
  double samplerate_ = 16000.0;
  // realised results
  btkb_pipeline* pipe_ = nullptr;
  std::vector<std::complex<float>> Y_; std::vector<float> time_, pfw_; std::vector<std::complex<float>> X_;
  std::vector<std::complex<float>> W_;   // [K][C] weights read back
  int T_ = 0, nb_ = 0; bool realized_ = false, haveX_ = false;
  PostFilterConfig pf_used_; SynthesisConfig syn_used_;
  LmsConfig lms_; RlsConfig rls_;
  bool normalize_weight_ = false;          // SubbandGSC::normalize_weight (beamformer.h:177)
  // wq (or wmvdr) and wl = B wa of the configured weights, [K][C] each, realised on a one-block dummy batch
  void fetch_static_weights_(std::vector<std::complex<float>>& W, std::vector<std::complex<float>>& WL);
  MultiChannelWPEDereverberationPtr wpe_;   // set by run_graph when the channels are MultiChannelWPEDereverberationFeature streams
  SnapShotArrayPtr snap_;
  std::vector<unsigned long> src_versions_;
  void ensure_pipeline_(const PostFilterConfig& pf, const SynthesisConfig& syn, unsigned n_samples);
  // chunked realisation state
  int chunk_blocks_ = 0;
  bool live_ = false, weights_dirty_ = false;
  std::vector<const SampleFeature*> srcs_; unsigned n_total_ = 0, D_ = 0; size_t pos_ = 0;
  int T_ready_ = 0, nb_ready_ = 0, chunk_t0_ = 0;
  std::vector<float> xchunk_;
  void advance_();
};
typedef std::shared_ptr<SubbandBeamformer> SubbandBeamformerPtr;

class SubbandDS : public SubbandBeamformer {
 public:
  SubbandDS(unsigned fftLen = 512, bool half_band_shift = false, const std::string& nm = "SubbandDS", int kind = 0);
  virtual void calc_array_manifold_vectors(double samplerate, const std::vector<double>& delays);
  // LCMV weights without a blocking matrix: one target + NC-1 jammers, delaysJ flat [NC-1][C] (beamformer.cc:1057-1074)
  void calc_array_manifold_vectors_n(double samplerate, const std::vector<double>& delaysT, const std::vector<double>& delaysJ, unsigned NC = 2);
  void calc_array_manifold_vectors_2(double samplerate, const std::vector<double>& delaysT, const std::vector<double>& delaysJ) {
    calc_array_manifold_vectors_n(samplerate, delaysT, delaysJ, 2);
  }
  void clear_channel() override;
 protected:
  void configure_weights_(btkb_pipeline* p) override;
  void set_delays_(btkb_pipeline* p);   // btkb_set_delays or, with NC_ > 1, btkb_set_delays_lcmv
  std::vector<double> delays_; bool have_delays_ = false;
  std::vector<double> delaysJ_; unsigned NC_ = 1;
};
typedef std::shared_ptr<SubbandDS> SubbandDSPtr;

class SubbandGSC : public SubbandDS {
 public:
  SubbandGSC(unsigned fftLen = 512, bool half_band_shift = false, const std::string& nm = "SubbandGSC");
  void calc_gsc_weights(double samplerate, const std::vector<double>& delaysT) { NC_ = 1; delaysJ_.clear(); calc_array_manifold_vectors(samplerate, delaysT); }
  // LCMV quiescent weights: one target + NC-1 jammers, delaysJ flat [NC-1][C] (beamformer.cc:1332-1362)
  void calc_gsc_weights_n(double samplerate, const std::vector<double>& delaysT, const std::vector<double>& delaysJ, unsigned NC = 2);
  void calc_gsc_weights_2(double samplerate, const std::vector<double>& delaysT, const std::vector<double>& delaysJ) { calc_gsc_weights_n(samplerate, delaysT, delaysJ, 2); }
  void set_active_weights_f(unsigned fbinX, const std::vector<double>& packedWeight);
  void zero_active_weights();
  void normalize_weight(bool flag) { normalize_weight_ = flag; invalidate_(); }                // beamformer.h:177
  void calc_array_manifold_vectors(double samplerate, const std::vector<double>& delays) override { have_wq_explicit_ = false; SubbandDS::calc_array_manifold_vectors(samplerate, delays); }
  // beamformer.cc:1318-1324.  The reference re-allocates the whole BeamformerWeights object on every call (alloc_bfweight_(1, 1)),
  // so only the bin of the LAST call keeps a non-zero quiescent vector; reproduced.
  void set_quiescent_weights_f(unsigned fbinX, const std::vector<cplx>& srcWq);
  // beamformer.cc:775-828: conj(wq - wl) e^{j pi (f+1)} -> inverse DFT -> window -> text file ("C M" header, one row per channel)
  bool write_fir_coeff(const std::string& fn, unsigned winType = 1);
 protected:
  void configure_weights_(btkb_pipeline* p) override;
  std::vector<std::complex<float>> wa_;  // [K][C-NC]
  bool have_wa_ = false;
  std::vector<std::complex<float>> wq_explicit_; bool have_wq_explicit_ = false;   // set_quiescent_weights_f
};
typedef std::shared_ptr<SubbandGSC> SubbandGSCPtr;

// native body of pybeamformer.SubbandGSCLMSBeamformer
class SubbandGSCLMS : public SubbandDS {
 public:
  SubbandGSCLMS(unsigned fftLen, const LmsConfig& cfg, const std::string& nm = "SubbandGSCLMS");
  void calc_beamformer_weights(double samplerate, const std::vector<double>& delays) { calc_array_manifold_vectors(samplerate, delays); }
  std::vector<std::complex<float>> active_weights();   // [K][C-1] after the run (waH)
  int total_updates();
};
typedef std::shared_ptr<SubbandGSCLMS> SubbandGSCLMSPtr;

// native body of pybeamformer.SubbandGSCRLSBeamformer (lib/pybeamformer.py:765-928; NOT the reference's C++ SubbandGSCRLS,
// beamformer.cc:1447-1699, which forms Z = B^H x and is a different recursion)
class SubbandGSCRLSNative : public SubbandDS {
 public:
  SubbandGSCRLSNative(unsigned fftLen, const RlsConfig& cfg, const std::string& nm = "SubbandGSCRLSNative");
  void calc_beamformer_weights(double samplerate, const std::vector<double>& delays) { calc_array_manifold_vectors(samplerate, delays); }
  std::vector<std::complex<float>> active_weights();
  int total_updates();
};
typedef std::shared_ptr<SubbandGSCRLSNative> SubbandGSCRLSNativePtr;

// The reference's C++ class SubbandGSCRLS (beamformer.h:236-275, beamformer.cc:1447-1699, beamformer.i:289-339): GSC with an RLS
// sidelobe canceller in the blocking-matrix basis (Z = B^H x), fp64 on the GPU (BTKB_BF_GSC_RLS_CPP, csrc/btkb_rls_cpp.cu).
class SubbandGSCRLS : public SubbandGSC {
 public:
  SubbandGSCRLS(unsigned fftLen = 512, bool half_band_shift = false, float myu = 0.9f, float sigma2 = 0.01f, const std::string& nm = "SubbandGSCRLS");
  void init_precision_matrix(float sigma2 = 0.01f) { init_sigma2_ = sigma2; have_pz_ = true; invalidate_(); }                 // beamformer.cc:1479-1492
  void update_active_weight_vecotrs(bool flag) { update_ = flag; invalidate_(); }                                           // beamformer.h:252 (sic)
  void set_quadratic_constraint(float alpha, int qctype = 1) { alpha_ = alpha; qctype_ = qctype; invalidate_(); }            // beamformer.h:253-254
  // set_precision_matrix(fbinX, Pz) (per-bin start matrices, beamformer.cc:1494-1506) is not mirrored: see INTEGRATION.md
 protected:
  void configure_weights_(btkb_pipeline* p) override;
  void tune_config_(btkb_config& c) override;
  float mu_, sigma2_, init_sigma2_ = 0.01f, alpha_ = -1.0f; int qctype_ = 0; bool update_ = true, have_pz_ = false;
};
typedef std::shared_ptr<SubbandGSCRLS> SubbandGSCRLSPtr;

// native body of pybeamformer.SubbandSOSBatchBeamformer / SubbandBlindMVDRBeamformer / SubbandGEVBeamformer
// (lib/pybeamformer.py:1026-1357): the statistics live on the device in a pipeline of their own and accumulate over calls
// like the reference's _target/_noise_covariance_matrices; the weights are handed to the beamforming pipeline as explicit
// quiescent vectors (y = w^H x for every bin, pybeamformer.py:1191-1207).
class SubbandSOSNative : public SubbandDS {
 public:
  SubbandSOSNative(unsigned fftLen, const std::string& nm = "SubbandSOSNative");
  ~SubbandSOSNative();
  void reset_stats();
  // labels: flat [NL][2] (start, end) seconds
  void accu_stats_from_label(double samplerate, const std::vector<double>& labels, double energy_threshold);
  // masks: [rows][cols] float, cols >= fftLen/2+1 (only the first fftLen/2+1 columns are read, like the reference's loop)
  void accu_stats_from_tfmask(double samplerate, const std::vector<float>& mask_t, const std::vector<float>& mask_j, unsigned rows, unsigned cols,
                              double energy_threshold);
  void calc_weights(int kind, double gamma, int ref_micx, double offset);   // finalize_stats + calc_beamformer_weights
  std::vector<double> frame_counts();   // [K][2] target / noise
  void clear_channel() override;
 protected:
  void configure_weights_(btkb_pipeline* p) override;
  unsigned stage_(double samplerate);   // analysis of the current utterance into the statistics pipeline
  btkb_pipeline* stats_ = nullptr; unsigned stats_cap_ = 0;
  std::vector<std::complex<float>> wsos_; bool have_wsos_ = false;
};
typedef std::shared_ptr<SubbandSOSNative> SubbandSOSNativePtr;

class SubbandMVDR : public SubbandDS {
 public:
  SubbandMVDR(unsigned fftLen = 512, bool half_band_shift = false, const std::string& nm = "SubbandMVDR");
  bool calc_mvdr_weights(double samplerate, double dThreshold = 1.0e-8, bool calc_inverse_matrix = true);
  std::vector<cplx> mvdr_weights(unsigned fbinX) { return get_weights(fbinX); }
  bool set_noise_spatial_spectral_matrix(unsigned fbinX, const std::vector<cplx>& Rnn);   // row-major C x C
  bool set_diffuse_noise_model(const std::vector<double>& micPositions /* [C][3] */, double samplerate, double sspeed = 343740.0);
  void set_all_diagonal_loading(double diagonalWeight);
  void set_diagonal_looading(unsigned fbinX, float diagonalWeight);          // (sic) beamformer.cc:2525-2535: R[fbinX] += w I
  void divide_nondiagonal_elements(unsigned fbinX, float mu);               // beamformer.cc:2589-2599: off-diagonals /= (1 + mu)
  void divide_all_nondiagonal_elements(float mu);                           // beamformer.h:357-360, bins 0..M/2
  void clear_channel() override;
  // SMI statistics on the GPU (pybeamformer.py:948-1000): noise frames outside [start,end] with energy > threshold
  int accumulate_noise_covariance(double samplerate, double label_start, double label_end, double energy_threshold);
 protected:
  void configure_weights_(btkb_pipeline* p) override;
  std::vector<std::complex<float>> R_;   // [K][C][C]
  std::vector<std::complex<float>> wmvdr_;  // [K][C]
  bool have_R_ = false, have_w_ = false, diffuse_ = false, smi_ = false;
  std::vector<double> mpos_; double sspeed_ = 343740.0; double mu_ = 0.0;
  // per-bin edits of R applied when the weights are configured (they commute: loading touches the diagonal only, the division
  // the off-diagonals only): extra diagonal load and the accumulated off-diagonal divisor, [K] each, empty = none
  std::vector<double> load_f_, div_f_;
  // calc_mvdr_weights(samplerate, dThreshold, calc_inverse_matrix): the matrix the weights are solved with is the one present at the
  // last call with calc_inverse_matrix = true (the reference keeps invR_ and reuses it, beamformer.cc:2379-2384)
  struct InvSource { std::vector<std::complex<float>> R; bool diffuse = false; std::vector<double> mpos, load_f, div_f; double sspeed = 343740.0, mu = 0.0; bool valid = false; };
  InvSource inv_; float dthreshold_ = 1.0e-8f;
};
typedef std::shared_ptr<SubbandMVDR> SubbandMVDRPtr;

class SubbandMVDRGSC : public SubbandMVDR {
 public:
  SubbandMVDRGSC(unsigned fftLen = 512, bool half_band_shift = false, const std::string& nm = "SubbandMVDR") : SubbandMVDR(fftLen, half_band_shift, nm) {}
  void set_active_weights_f(unsigned fbinX, const std::vector<double>& packedWeight);
  void zero_active_weights() { wa_.clear(); have_wa_ = false; invalidate_(); }
  // B orthogonal to the delay-and-sum weights (beamformer.cc:2638-2643); like the reference's alloc_bfweight_ both variants drop
  // the active weights set before
  bool calc_blocking_matrix1(double samplerate, const std::vector<double>& delaysT) {
    calc_array_manifold_vectors(samplerate, delaysT); bm_from_mvdr_ = false; wa_.clear(); have_wa_ = false; upgrades_.clear(); return true;
  }
  // B orthogonal to the MVDR weights (beamformer.cc:2649-2672); false when calc_mvdr_weights() has not been called
  bool calc_blocking_matrix2() {
    if (!have_w_) return false;
    bm_from_mvdr_ = true; wa_.clear(); have_wa_ = false; upgrades_.clear(); invalidate_(); return true;
  }
  // beamformer.cc:2674-2691: the blocking matrix of the bins >= 1 becomes the one orthogonal to wq - wl (wl = B wa of the active
  // weights set so far); active weights set afterwards go through the new matrix.  The mirror keeps the active weights of each
  // upgrade and replays set / upgrade in order when the pipeline is configured (whole vectors: an upgrade between the per-bin
  // set_active_weights_f calls of ONE weight vector is not tracked bin by bin).
  void upgrade_blocking_matrix();
  // beamformer.cc:2693-2716: b_outChanX^H x of the CURRENT frame for the bins 0..M/2; the upper half of the returned vector is what
  // the reference's shared output vector holds there after next(): the conjugate-symmetric half of the beamformer output
  const cplx* blocking_matrix_output(int outChanX = 0);
 protected:
  void configure_weights_(btkb_pipeline* p) override;
  std::vector<std::complex<float>> wa_; bool have_wa_ = false, bm_from_mvdr_ = false;
  std::vector<std::vector<std::complex<float>>> upgrades_;   // active weights in force at each upgrade_blocking_matrix()
  bool wa_since_upgrade_ = false;
  std::vector<std::complex<float>> Z_; int Z_chan_ = -1; const void* Z_of_ = nullptr; int Z_T_ = 0;
  std::vector<cplx> bmout_;
};
typedef std::shared_ptr<SubbandMVDRGSC> SubbandMVDRGSCPtr;

// SubbandOrthogonalizer (beamformer.h / beamformer.cc:2776-2806): streams the beamformer output (outChanX <= 0) or the output of
// blocking-matrix branch outChanX - 1 of a SubbandMVDRGSC
class SubbandOrthogonalizer : public VectorComplexFeatureStream {
 public:
  SubbandOrthogonalizer(const SubbandMVDRGSCPtr& beamformer, int outChanX = 0, const std::string& nm = "SubbandOrthogonalizer")
      : VectorComplexFeatureStream(beamformer->fftLen(), nm), beamformer_(beamformer), outChanX_(outChanX) {}
  const cplx* next(int frame_no = -5) override;
  void reset() override { beamformer_->reset(); VectorComplexFeatureStream::reset(); }
 private:
  SubbandMVDRGSCPtr beamformer_; int outChanX_;
};
typedef std::shared_ptr<SubbandOrthogonalizer> SubbandOrthogonalizerPtr;

// ----------------------------------------------------------------------------------------------------------------
class ZelinskiPostFilter : public VectorComplexFeatureStream {
 public:
  ZelinskiPostFilter(const VectorComplexFeatureStreamPtr& output, unsigned fftLen, double alpha = 0.6, int type = 2, int min_frames = 0,
                     const std::string& nm = "ZelinskPostFilter");
  const cplx* next(int frame_no = -5) override;
  void reset() override;
  void set_beamformer(const SubbandDSPtr& bf) { bf_ = bf; }
  // postfilter.cc:384-417: the wiring WITHOUT a beamformer object — the snapshots of a foreign stream and one time-alignment vector per bin
  // (wq for the TYPE_ZELINSKI2 variants, the array manifold otherwise: both are "the vector the filter aligns the channels with").
  // The mirror drains the output stream at the first next(), copying the shared snapshot array after every frame, runs the filter
  // statistics on the GPU (btkb_set_snapshots + the Zelinski epilogue of the per-bin kernel) and multiplies the foreign output by the gains.
  void set_snapshot_array(const SnapShotArrayPtr& snap) { snap_foreign_ = snap; foreign_ready_ = false; }
  void set_array_manifold_vector(unsigned fbinX, const std::vector<cplx>& v, bool half_band_shift = false, unsigned NC = 1);
  ~ZelinskiPostFilter();
  std::vector<cplx> postfilter_weights();
  const SubbandDSPtr& beamformer() const { return bf_; }
  virtual PostFilterConfig config() const { PostFilterConfig c; c.enabled = true; c.alpha = alpha_; c.type = type_; c.min_frames = min_frames_; return c; }
  const VectorComplexFeatureStreamPtr& source() const { return samp_; }
 protected:
  virtual int onesided_frames_() const { return 0; }  // frames whose upper half-spectrum the reference leaves untouched
  unsigned fftLen_; VectorComplexFeatureStreamPtr samp_; double alpha_; int type_, min_frames_;
  SubbandDSPtr bf_;
  // foreign-stream wiring
  SnapShotArrayPtr snap_foreign_; std::vector<std::complex<float>> manifold_; unsigned manifold_C_ = 0; bool foreign_ready_ = false;
  btkb_pipeline* fpipe_ = nullptr; std::vector<cplx> fout_; std::vector<float> fgain_; int fT_ = 0;
  void realize_foreign_();
};
typedef std::shared_ptr<ZelinskiPostFilter> ZelinskiPostFilterPtr;

// McCowanPostFilter (postfilter.h / postfilter.cc:496-934).  The noise coherence R_[fbinX] lives on the device in a small
// pipeline of its own, so every setter below is the corresponding C-ABI call; the beamformer's pipeline receives a copy
// when the graph runs.
class McCowanPostFilter : public ZelinskiPostFilter {
 public:
  McCowanPostFilter(const VectorComplexFeatureStreamPtr& output, unsigned fftLen, double alpha = 0.6, int type = 2, int min_frames = 0, float threshold = 0.99f,
                    const std::string& nm = "McCowanPostFilterPtr");
  ~McCowanPostFilter();
  std::vector<cplx> noise_spatial_spectral_matrix(unsigned fbinX);                        // [C][C] row-major
  bool set_noise_spatial_spectral_matrix(unsigned fbinX, const std::vector<cplx>& Rnn, unsigned rows, unsigned cols);
  bool set_diffuse_noise_model(const std::vector<double>& mpos, unsigned rows, unsigned cols, double sampleRate, double sspeed = 343740.0);
  void set_all_diagonal_loading(float diagonalWeight);
  void set_diagonal_looading(unsigned fbinX, float diagonalWeight);
  void divide_all_nondiagonal_elements(float mu);
  void divide_nondiagonal_elements(unsigned fbinX, float mu);
  PostFilterConfig config() const override;
  void push_coherence(btkb_pipeline* dst) const;   // copy R_ into the beamformer's pipeline
  unsigned chanN() const { return chanN_; }
 protected:
  int onesided_frames_() const override { return min_frames_ + 1; }  // postfilter.cc:896-901
  void ensure_store_(unsigned C);
  void require_R_() const;
  std::vector<cplx> get_all_() const; void set_all_(const std::vector<cplx>& R);
  int kind_; float threshold_; double min_sv_ = 1.0e-8; unsigned fbin1_ = 0;
  btkb_pipeline* store_ = nullptr; unsigned chanN_ = 0; bool haveR_ = false; unsigned long version_ = 0;
};
typedef std::shared_ptr<McCowanPostFilter> McCowanPostFilterPtr;

// LefkimmiatisPostFilter (postfilter.cc:935-1200)
class LefkimmiatisPostFilter : public McCowanPostFilter {
 public:
  LefkimmiatisPostFilter(const VectorComplexFeatureStreamPtr& output, unsigned fftLen, double minSV = 1.0e-8, unsigned fbinX1 = 0, double alpha = 0.6, int type = 2,
                         int min_frames = 0, float threshold = 0.99f, const std::string& nm = "LefkimmiatisPostFilterPtr");
  void calc_inverse_noise_spatial_spectral_matrix();  // the inverse is (re)computed on the device whenever the graph runs
};
typedef std::shared_ptr<LefkimmiatisPostFilter> LefkimmiatisPostFilterPtr;

class OverSampledDFTSynthesisBank : public VectorFloatFeatureStream {
 public:
  OverSampledDFTSynthesisBank(const VectorComplexFeatureStreamPtr& samp, const std::vector<double>& prototype, unsigned M, unsigned m, unsigned r = 0,
                              unsigned delay_compensation_type = 0, int gain_factor = 1, const std::string& nm = "OverSampledDFTSynthesisBank");
  ~OverSampledDFTSynthesisBank();
  const float* next(int frame_no = -5) override;
  void reset() override;
  double polyphase(unsigned m, unsigned n) const { return prototype_.at(m + M_ * n); }
  // modulated.h:330: hand the bank one subband frame directly (update_buf_); the frames pushed before the first next() enter the
  // buffer ahead of the source's frames, so the outputs are those of the concatenated sequence from block (number pushed) on; r = 0 only
  void input_source_vector(const std::vector<cplx>& block);
 private:
  void realize_();
  std::vector<std::complex<float>> pushed_; int npushed_ = 0;
  VectorComplexFeatureStreamPtr samp_;
  std::vector<double> prototype_;
  unsigned M_, m_, r_, D_, dct_; int gain_; int pd_;
  btkb_pipeline* pipe_;            // only for the generic (arbitrary upstream) path
  std::vector<float> out_; int nb_; bool realized_;
  SubbandBeamformer* live_bf_ = nullptr;   // chunked upstream realisation: output blocks are pulled from it as next() asks for them
};
typedef std::shared_ptr<OverSampledDFTSynthesisBank> OverSampledDFTSynthesisBankPtr;

// delays (beamformer.cc:1170-1189 calc_all_delays)
std::vector<double> calc_all_delays(double x, double y, double z, const std::vector<double>& mpos /* [C][3] */);

}  // namespace btk20
