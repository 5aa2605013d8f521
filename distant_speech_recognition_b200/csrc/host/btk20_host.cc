// btk20_host.cc — implementation of the C++ host mirror (see btk20_host.h).  Everything numeric goes through the C-ABI.
#include "btk20_host.h"

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <fstream>

#include "../../../include/btkb.h"

namespace btk20 {

namespace {
void ck(int rc) {
  if (rc == BTKB_OK) return;
  const char* m = btkb_last_error();
  switch (rc) {
    case BTKB_ERR_ALLOC: throw jallocation_error("%s", m);
    case BTKB_ERR_INVALID: throw jdimension_error("%s", m);
    default: throw j_error("%s", m);
  }
}
int analysis_frames(int n, int D, int m, int r, int dct) {  // modulated.cc:246-264,418-469
  const int R = 1 << r;
  int pd, la = 0;
  if (dct == 1) pd = m * R - 1; else if (dct == 2) { pd = m * R - 1; la = m * R / 2 - 1; } else pd = 2 * m - 1;
  return (n + D - 1) / D - la + pd;
}
int synthesis_delay(int m, int r, int dct) { const int R = 1 << r; return dct == 1 ? m * R - 1 : (dct == 2 ? m * R / 2 : 2 * m - 1); }
}  // namespace

// ================================================================================================ SampleFeature
SampleFeature::SampleFeature(const std::string& fn, unsigned block_len, unsigned shift_len, bool pad_zeros, const std::string& nm)
    : VectorFloatFeatureStream(block_len, nm), shift_len_(shift_len), cur_(0), pad_zeros_(pad_zeros), samplerate_(16000), version_(0) {
  if (!fn.empty()) read(fn);
}

// Minimal RIFF/WAVE reader (PCM 16/32-bit, IEEE float32) standing in for libsndfile (feature.cc:238-389): norm == 0
// keeps int16-scale floats (SFC_SET_NORM_FLOAT off), otherwise samples are scaled to [-1, 1) * norm.
unsigned SampleFeature::read(const std::string& fn, int /*format*/, int samplerate, int chX, int /*chN*/, int cfrom, int to, int /*outsamplerate*/, float norm) {
  std::ifstream f(fn, std::ios::binary);
  if (!f) throw jio_error("Could not open file %s.", fn.c_str());
  char id[4]; uint32_t sz;
  f.read(id, 4); f.read((char*)&sz, 4); char wave[4]; f.read(wave, 4);
  if (!f || std::memcmp(id, "RIFF", 4) || std::memcmp(wave, "WAVE", 4)) throw jio_error("sndfile error: %s is not a RIFF/WAVE file.", fn.c_str());
  uint16_t fmt = 1, nch = 1, bits = 16; uint32_t rate = (uint32_t)samplerate;
  std::vector<char> data;
  while (f.read(id, 4) && f.read((char*)&sz, 4)) {
    if (!std::memcmp(id, "fmt ", 4)) {
      std::vector<char> b(sz); f.read(b.data(), sz);
      std::memcpy(&fmt, &b[0], 2); std::memcpy(&nch, &b[2], 2); std::memcpy(&rate, &b[4], 4); std::memcpy(&bits, &b[14], 2);
    } else if (!std::memcmp(id, "data", 4)) {
      data.resize(sz); f.read(data.data(), sz); break;
    } else f.seekg(sz + (sz & 1), std::ios::cur);
  }
  if (data.empty()) throw jio_error("sndfile error: no data chunk in %s.", fn.c_str());
  const int bytes = bits / 8; const long frames = (long)(data.size() / (size_t)(bytes * nch));
  if (to < 0 || to >= frames) to = (int)frames - 1;
  if (cfrom < 0) cfrom = 0;
  if (cfrom > to || cfrom > frames) throw jio_error("Cannot load samples from %d to %d.", cfrom, to);
  if (chX < 1 || chX > nch) throw jio_error("Channel %d is not in the file (%d channels).", chX, (int)nch);
  const int n = to - cfrom + 1;
  samples_.resize(n);
  for (int i = 0; i < n; i++) {
    const char* p = &data[((size_t)(cfrom + i) * nch + (chX - 1)) * bytes];
    float v;
    // libsndfile semantics of sf_readf_float (feature.cc:265-276, 305): float files are returned as they are; integer files unscaled with
    // SFC_SET_NORM_FLOAT off (norm == 0), scaled to [-1, 1) with it on; then samples *= norm unless norm is 0 or 1 (feature.cc:339-341)
    if (fmt == 3 && bits == 32) { std::memcpy(&v, p, 4); if (norm != 0.0f && norm != 1.0f) v *= norm; }
    else if (bits == 16) { int16_t s; std::memcpy(&s, p, 2); v = (norm == 0.0f) ? (float)s : (float)s / 32768.0f * norm; }
    else if (bits == 32) { int32_t s; std::memcpy(&s, p, 4); v = (norm == 0.0f) ? (float)s : (float)s / 2147483648.0f * norm; }
    else throw jio_error("sndfile error: unsupported sample format (%d bits).", (int)bits);
    samples_[i] = v;
  }
  samplerate_ = (int)rate;
  version_++;
  reset();
  return (unsigned)n;
}

void SampleFeature::set_samples(const double* samples, unsigned n, unsigned samplerate) {  // feature.cc:669-679
  samples_.resize(n);
  for (unsigned i = 0; i < n; i++) samples_[i] = (float)samples[i];
  samplerate_ = (int)samplerate;
  version_++;
  reset();
}

const float* SampleFeature::next(int frame_no) {  // feature.cc:605-649
  if (is_end_) throw jiterator_error("end of samples!");
  if (frame_no == frame_no_) return vector_.data();
  if (frame_no >= 0 && frame_no - 1 != frame_no_) throw jindex_error("Problem in Feature %s: %d != %d\n", name().c_str(), frame_no - 1, frame_no_);
  const unsigned ttl = (unsigned)samples_.size();
  if (cur_ >= ttl) { is_end_ = true; throw jiterator_error("end of samples!"); }
  if (cur_ + size() >= ttl) {
    if (pad_zeros_) {
      std::fill(vector_.begin(), vector_.end(), 0.f);
      for (unsigned i = 0; i < ttl - cur_; i++) vector_[i] = samples_[cur_ + i];
    } else { is_end_ = true; throw jiterator_error("end of samples!"); }
  } else {
    for (unsigned i = 0; i < size(); i++) vector_[i] = samples_[cur_ + i];
  }
  cur_ += shift_len_;
  increment_();
  return vector_.data();
}

// ================================================================================================ analysis bank
OverSampledDFTAnalysisBank::OverSampledDFTAnalysisBank(const VectorFloatFeatureStreamPtr& samp, const std::vector<double>& prototype, unsigned M, unsigned m,
                                                       unsigned r, unsigned dct, const std::string& nm)
    : VectorComplexFeatureStream(M, nm), samp_(samp), prototype_(prototype), M_(M), m_(m), r_(r), D_(M >> r), dct_(dct), pipe_(nullptr), T_(0), realized_(false) {
  if (prototype.size() != (size_t)M * m) throw jconsistency_error("Prototype sizes do not match (%d vs. %d).", (int)prototype.size(), (int)(M * m));  // modulated.cc:239-241
  if (samp_->size() != D_) throw jdimension_error("Input block length (%d) != D_ (%d)\n", samp_->size(), D_);                                     // modulated.cc:337-338
}
OverSampledDFTAnalysisBank::~OverSampledDFTAnalysisBank() { if (pipe_) btkb_destroy(pipe_); }

void OverSampledDFTAnalysisBank::realize_() {
  auto* sf = dynamic_cast<SampleFeature*>(samp_.get());
  if (!sf) throw j_error("OverSampledDFTAnalysisBank: the GPU engine needs a SampleFeature source (whole-utterance sample array)");
  const unsigned n = sf->samplesN();
  if (n == 0) throw jiterator_error("end of samples!");
  if (pipe_) { btkb_destroy(pipe_); pipe_ = nullptr; }
  btkb_config c; btkb_default_config(&c);
  c.channels = 1; c.fft_len = (int)M_; c.m = (int)m_; c.r = (int)r_; c.delay_compensation_type = (int)dct_; c.max_utterances = 1; c.max_samples = (int)n;
  ck(btkb_create(&c, &pipe_));
  ck(btkb_set_prototypes(pipe_, prototype_.data(), nullptr, (int)prototype_.size()));
  ck(btkb_submit(pipe_, sf->samples().data(), 1, (int)n, nullptr));
  ck(btkb_run_analysis(pipe_));
  T_ = btkb_num_frames(pipe_);
  const unsigned K = M_ / 2 + 1;
  X_.resize((size_t)T_ * K);
  ck(btkb_fetch_snapshots(pipe_, reinterpret_cast<float*>(X_.data())));
  realized_ = true;
}

const cplx* OverSampledDFTAnalysisBank::next(int frame_no) {  // modulated.cc:375-409
  if (frame_no == frame_no_) return vector_.data();
  if (!realized_) realize_();
  if (frame_no_ + 1 >= T_) { is_end_ = true; throw jiterator_error("end of samples!"); }
  increment_();
  const unsigned K = M_ / 2 + 1;
  const std::complex<float>* x = &X_[(size_t)frame_no_ * K];
  for (unsigned k = 0; k < K; k++) vector_[k] = cplx(x[k].real(), x[k].imag());
  for (unsigned k = 1; k < M_ / 2; k++) vector_[M_ - k] = std::conj(vector_[k]);
  return vector_.data();
}
void OverSampledDFTAnalysisBank::reset() { samp_->reset(); VectorComplexFeatureStream::reset(); realized_ = false; }

// ================================================================================================ WPE dereverberation
MultiChannelWPEDereverberation::MultiChannelWPEDereverberation(unsigned subbandsN, unsigned channelsN, unsigned lowerN, unsigned upperN, unsigned iterationsN,
                                                               double loadDb, double bandWidth, double diagonalBias, double sampleRate)
    : subbandsN_(subbandsN), channelsN_(channelsN) {
  cfg_.lower_num = lowerN; cfg_.upper_num = upperN; cfg_.iterations_num = iterationsN; cfg_.load_db = loadDb; cfg_.band_width = bandWidth;
  cfg_.diagonal_bias = diagonalBias; cfg_.samplerate = sampleRate;
  if (bandWidth > sampleRate / 2.0) throw jdimension_error("Bandwidth is greater than the Nyquist rate.\n");   // dereverberation.cc:369-370
  if (channelsN < 1 || channelsN > 8) throw jdimension_error("the GPU WPE is built for 1..8 channels (got %d)", channelsN);
}
MultiChannelWPEDereverberation::~MultiChannelWPEDereverberation() { if (pipe_) btkb_destroy(pipe_); }
void MultiChannelWPEDereverberation::set_input(const VectorComplexFeatureStreamPtr& samples) {
  if (sources_.size() == channelsN_) throw jallocation_error("Channel capacity exceeded.");   // dereverberation.cc:395-396
  if (samples->size() != subbandsN_) throw jdimension_error("Input block length (%d) != the number of subbands (%d)\n", samples->size(), subbandsN_);
  sources_.push_back(samples);
}
unsigned MultiChannelWPEDereverberation::gather_(std::vector<float>& x) {
  if (sources_.size() != channelsN_) throw j_error("MultiChannelWPEDereverberation: %d of %d inputs set", (int)sources_.size(), (int)channelsN_);
  unsigned n = 0;
  versions_.assign(channelsN_, 0);
  std::vector<const SampleFeature*> sf(channelsN_);
  for (unsigned c = 0; c < channelsN_; c++) {
    auto* ab = dynamic_cast<OverSampledDFTAnalysisBank*>(sources_[c].get());
    if (!ab) throw j_error("MultiChannelWPEDereverberation: the GPU engine needs OverSampledDFTAnalysisBank inputs");
    if (ab->M() != subbandsN_) throw jdimension_error("channel %d: inconsistent FFT length (%d vs. %d)", c, ab->M(), subbandsN_);
    sf[c] = dynamic_cast<const SampleFeature*>(ab->source().get());
    if (!sf[c]) throw j_error("MultiChannelWPEDereverberation: the GPU engine needs SampleFeature sources");
    if (c == 0) n = sf[c]->samplesN();
    else if (sf[c]->samplesN() != n) throw jdimension_error("channel %d: %d samples, channel 0: %d", c, sf[c]->samplesN(), n);
    versions_[c] = sf[c]->version();
  }
  if (n == 0) throw jiterator_error("end of samples!");
  x.resize((size_t)channelsN_ * n);
  for (unsigned c = 0; c < channelsN_; c++) std::memcpy(&x[(size_t)c * n], sf[c]->samples().data(), sizeof(float) * n);
  return n;
}
unsigned MultiChannelWPEDereverberation::estimate_filter(int start_frame_no, int end_frame_no) {   // dereverberation.cc:405-431
  std::vector<float> x;
  const unsigned n = gather_(x);
  auto* a0 = dynamic_cast<OverSampledDFTAnalysisBank*>(sources_[0].get());
  if (!pipe_ || n > cap_) {
    if (pipe_) { btkb_destroy(pipe_); pipe_ = nullptr; }
    btkb_config c; btkb_default_config(&c);
    c.channels = (int)channelsN_; c.fft_len = (int)subbandsN_; c.m = (int)a0->m(); c.r = (int)a0->r(); c.delay_compensation_type = (int)a0->dct();
    c.samplerate = (float)cfg_.samplerate; c.beamformer = BTKB_BF_DS; c.max_utterances = 1; c.max_samples = (int)n;
    c.wpe.enabled = 1; c.wpe.lower_num = (int)cfg_.lower_num; c.wpe.upper_num = (int)cfg_.upper_num; c.wpe.iterations_num = (int)cfg_.iterations_num;
    c.wpe.load_db = cfg_.load_db; c.wpe.band_width = cfg_.band_width; c.wpe.diagonal_bias = cfg_.diagonal_bias;
    ck(btkb_create(&c, &pipe_));
    ck(btkb_set_prototypes(pipe_, a0->prototype().data(), nullptr, (int)a0->prototype().size()));
    cap_ = n;
  }
  ck(btkb_submit(pipe_, x.data(), 1, (int)n, nullptr));
  ck(btkb_run_analysis(pipe_));
  ck(btkb_run_wpe(pipe_, start_frame_no, end_frame_no));
  T_ = btkb_num_frames(pipe_);
  Xd_.resize((size_t)T_ * channelsN_ * (subbandsN_ / 2 + 1));
  ck(btkb_fetch_snapshots(pipe_, reinterpret_cast<float*>(Xd_.data())));
  est_start_ = start_frame_no; est_end_ = end_frame_no;
  estimated_ = true; realized_ = true;
  for (auto& s : sources_) s->reset();
  // framesN_: fill_buffer_ (:500-534) takes frames from the head of the stream, (end - start) of them when end >= 0
  const int want = (end_frame_no < 0) ? T_ : std::max(0, end_frame_no - std::max(start_frame_no, 0));
  return (unsigned)std::min(T_, want);
}
void MultiChannelWPEDereverberation::realize_() {
  // the sources have been re-read since the estimation: analysis + output stage with the stored filters
  std::vector<float> x;
  const unsigned n = gather_(x);
  if (n > cap_) throw j_error("MultiChannelWPEDereverberation: the utterance (%d samples) is longer than the one the filter was estimated on (%d)", n, cap_);
  ck(btkb_submit(pipe_, x.data(), 1, (int)n, nullptr));
  ck(btkb_run_analysis(pipe_));
  ck(btkb_apply_wpe(pipe_));
  T_ = btkb_num_frames(pipe_);
  Xd_.resize((size_t)T_ * channelsN_ * (subbandsN_ / 2 + 1));
  ck(btkb_fetch_snapshots(pipe_, reinterpret_cast<float*>(Xd_.data())));
  realized_ = true;
}
std::vector<float> MultiChannelWPEDereverberation::filters() const {
  if (!estimated_ || !pipe_) throw jinitialization_error("Call MultiChannelWPEDereverberation::estimate_filter()\n");
  const size_t P = cfg_.upper_num - cfg_.lower_num + 1;
  std::vector<float> G((size_t)(subbandsN_ / 2 + 1) * channelsN_ * channelsN_ * P * 2);
  ck(btkb_get_wpe_filter(pipe_, G.data()));
  return G;
}
int MultiChannelWPEDereverberation::frames() {
  if (!estimated_) throw jinitialization_error("Call SingleChannelWPEDereverberationFeature::estimate_filter()\n");   // dereverberation.cc:446-447 (sic)
  bool stale = !realized_;
  for (unsigned c = 0; c < channelsN_ && !stale; c++) {
    auto* ab = dynamic_cast<OverSampledDFTAnalysisBank*>(sources_[c].get());
    auto* sf = ab ? dynamic_cast<const SampleFeature*>(ab->source().get()) : nullptr;
    if (!sf || sf->version() != versions_[c]) stale = true;
  }
  if (stale) realize_();
  return T_;
}
const std::complex<float>* MultiChannelWPEDereverberation::output_frame(unsigned channelX, int t) {
  if (channelX >= channelsN_) throw jindex_error("Invalid channel index: it exceeds the number of channels: %u >= %u\n", channelX, channelsN_);   // :434-437
  return &Xd_[((size_t)t * channelsN_ + channelX) * (subbandsN_ / 2 + 1)];
}
void MultiChannelWPEDereverberation::reset() { for (auto& s : sources_) s->reset(); }
void MultiChannelWPEDereverberation::reset_filter() { estimated_ = false; realized_ = false; }
void MultiChannelWPEDereverberation::next_speaker() { reset(); estimated_ = false; realized_ = false; }   // :692-698 zeroes the filters: the output would equal the input

MultiChannelWPEDereverberationFeature::MultiChannelWPEDereverberationFeature(const MultiChannelWPEDereverberationPtr& source, unsigned channelX,
                                                                             unsigned primaryChannelX, const std::string& nm)
    : VectorComplexFeatureStream(source->size(), nm), source_(source), channelX_(channelX), primaryChannelX_(primaryChannelX) {}
const cplx* MultiChannelWPEDereverberationFeature::next(int frame_no) {   // dereverberation.cc:713-728
  if (frame_no == frame_no_) return vector_.data();
  const int T = source_->frames();
  if (frame_no >= 0 && frame_no - 1 != frame_no_) throw jindex_error("Problem in 'MultiChannelWPEDereverberationFeature': %d - 1 != %d\n", frame_no, frame_no_);
  if (frame_no_ + 1 >= T) { is_end_ = true; throw jiterator_error("end of samples!"); }
  increment_();
  const unsigned M = size(), K = M / 2 + 1;
  const std::complex<float>* x = source_->output_frame(channelX_, frame_no_);
  for (unsigned k = 0; k < K; k++) vector_[k] = cplx(x[k].real(), x[k].imag());
  for (unsigned k = 1; k < M / 2; k++) vector_[M - k] = std::conj(vector_[k]);
  return vector_.data();
}
void MultiChannelWPEDereverberationFeature::reset() { source_->reset(); VectorComplexFeatureStream::reset(); is_end_ = false; }

SingleChannelWPEDereverberationFeature::SingleChannelWPEDereverberationFeature(const VectorComplexFeatureStreamPtr& samples, unsigned lowerN, unsigned upperN,
                                                                               unsigned iterationsN, double loadDb, double bandWidth, double sampleRate,
                                                                               const std::string& nm)
    : VectorComplexFeatureStream(samples->size(), nm),
      impl_(std::make_shared<MultiChannelWPEDereverberation>(samples->size(), 1, lowerN, upperN, iterationsN, loadDb, bandWidth, 0.0, sampleRate)) {
  impl_->set_input(samples);
}
const cplx* SingleChannelWPEDereverberationFeature::next(int frame_no) {   // dereverberation.cc:227-275
  if (!impl_->estimated()) throw jinitialization_error("Call SingleChannelWPEDereverberationFeature::estimate_filter()\n");
  if (frame_no == frame_no_) return vector_.data();
  if (frame_no >= 0 && frame_no - 1 != frame_no_) throw jindex_error("Problem in Feature %s: %d != %d\n", name().c_str(), frame_no - 1, frame_no_);
  const int T = impl_->frames();
  if (frame_no_ + 1 >= T) { is_end_ = true; throw jiterator_error("end of samples!"); }
  increment_();
  const unsigned M = size(), K = M / 2 + 1;
  const std::complex<float>* x = impl_->output_frame(0, frame_no_);
  for (unsigned k = 0; k < K; k++) vector_[k] = cplx(x[k].real(), x[k].imag());
  for (unsigned k = 1; k < M / 2; k++) vector_[M - k] = std::conj(vector_[k]);
  return vector_.data();
}
void SingleChannelWPEDereverberationFeature::reset() { impl_->reset(); VectorComplexFeatureStream::reset(); is_end_ = false; }

// ================================================================================================ beamformers
SubbandBeamformer::SubbandBeamformer(unsigned fftLen, bool hbs, int kind, const std::string& nm)
    : VectorComplexFeatureStream(fftLen, nm), fftLen_(fftLen), halfBandShift_(hbs), kind_(kind) {
  if (const char* e = std::getenv("BTK20_CHUNK_BLOCKS")) chunk_blocks_ = std::max(0, std::atoi(e));
  if (hbs) throw jallocation_error("halfBandShift==true is not yet supported\n");  // beamformer.cc:2283-2285 (MVDR); the GPU path is M/2+1 bins only
}
SubbandBeamformer::~SubbandBeamformer() { if (pipe_) btkb_destroy(pipe_); }
void SubbandBeamformer::set_channel(const VectorComplexFeatureStreamPtr& chan) { channels_.push_back(chan); invalidate_(); }
void SubbandBeamformer::clear_channel() { channels_.clear(); snap_.reset(); invalidate_(); if (pipe_) { btkb_destroy(pipe_); pipe_ = nullptr; } }
void SubbandBeamformer::reset() {
  for (auto& c : channels_) c->reset();
  if (snap_) snap_->zero();
  VectorComplexFeatureStream::reset();
  is_end_ = false;
  realized_ = false;  // adaptive state restarts with the utterance (pybeamformer.py:759-762)
  live_ = false;
}

bool SubbandBeamformer::realized_with(const PostFilterConfig& pf, const SynthesisConfig& syn) const {
  if (!realized_) return false;
  if (pf.enabled != pf_used_.enabled || (pf.enabled && (pf.alpha != pf_used_.alpha || pf.type != pf_used_.type || pf.min_frames != pf_used_.min_frames))) return false;
  if (pf.enabled && (pf.kind != pf_used_.kind || pf.threshold != pf_used_.threshold || pf.min_sv != pf_used_.min_sv || pf.fbin1 != pf_used_.fbin1 ||
                     pf.coherence != pf_used_.coherence || pf.coherence_version != pf_used_.coherence_version)) return false;
  if (syn.enabled && !syn_used_.enabled) return false;
  return true;
}

// The analysis bank behind channel c: the channel itself, or — for the configs[4] chain analysis -> WPE -> beamformer — the
// input of the MultiChannelWPEDereverberation a MultiChannelWPEDereverberationFeature channel belongs to.
static OverSampledDFTAnalysisBank* bank_behind(const VectorComplexFeatureStreamPtr& ch, MultiChannelWPEDereverberationPtr* wpe) {
  if (auto* ab = dynamic_cast<OverSampledDFTAnalysisBank*>(ch.get())) return ab;
  if (auto* wf = dynamic_cast<MultiChannelWPEDereverberationFeature*>(ch.get())) {
    const MultiChannelWPEDereverberationPtr& pre = wf->source();
    if (wf->channel() >= pre->sources().size()) throw j_error("MultiChannelWPEDereverberationFeature: channel %d has no input", (int)wf->channel());
    if (wpe) {
      if (*wpe && wpe->get() != pre.get()) throw j_error("SubbandBeamformer: all channels must come from one MultiChannelWPEDereverberation");
      *wpe = pre;
    }
    return dynamic_cast<OverSampledDFTAnalysisBank*>(pre->sources()[wf->channel()].get());
  }
  return nullptr;
}

void SubbandBeamformer::ensure_pipeline_(const PostFilterConfig& pf, const SynthesisConfig& syn, unsigned n_samples) {
  auto* a0 = bank_behind(channels_[0], nullptr);
  btkb_config c; btkb_default_config(&c);
  c.channels = (int)channels_.size(); c.fft_len = (int)fftLen_; c.m = (int)a0->m(); c.r = (int)a0->r(); c.delay_compensation_type = (int)a0->dct();
  c.samplerate = (float)samplerate_; c.beamformer = kind_;
  c.postfilter = pf.enabled ? pf.kind : BTKB_PF_NONE; c.pf_alpha = (float)pf.alpha; c.pf_type = pf.type; c.pf_min_frames = pf.min_frames;
  c.pf_threshold = pf.threshold; c.pf_min_sv = pf.min_sv; c.pf_fbin1 = (int)pf.fbin1;
  c.lms.beta = (float)lms_.beta; c.lms.gamma = (float)lms_.gamma; c.lms.init_diagonal_load = (float)lms_.init_diagonal_load;
  c.lms.regularization_param = (float)lms_.regularization_param; c.lms.energy_floor = (float)lms_.energy_floor; c.lms.sil_thresh = (float)lms_.sil_thresh;
  c.lms.max_wa_l2norm = (float)lms_.max_wa_l2norm; c.lms.min_frames = lms_.min_frames; c.lms.slowdown_after = lms_.slowdown_after;
  c.rls.beta = (float)rls_.beta; c.rls.gamma = (float)rls_.gamma; c.rls.mu = (float)rls_.mu; c.rls.init_diagonal_load = (float)rls_.init_diagonal_load;
  c.rls.regularization_param = (float)rls_.regularization_param; c.rls.sil_thresh = (float)rls_.sil_thresh; c.rls.alpha2 = (float)rls_.alpha2;
  c.rls.max_wa_l2norm = (float)rls_.max_wa_l2norm; c.rls.constraint_option = rls_.constraint_option; c.rls.min_frames = rls_.min_frames;
  c.max_utterances = 1; c.max_samples = (int)n_samples; c.synthesis_gain = syn.enabled ? syn.gain : 1;
  c.normalize_weight = normalize_weight_ ? 1 : 0;
  if (wpe_) {
    const WpeConfig& w = wpe_->config();
    c.wpe.enabled = 1; c.wpe.lower_num = (int)w.lower_num; c.wpe.upper_num = (int)w.upper_num; c.wpe.iterations_num = (int)w.iterations_num;
    c.wpe.load_db = w.load_db; c.wpe.band_width = w.band_width; c.wpe.diagonal_bias = w.diagonal_bias;
  }
  tune_config_(c);
  if (pipe_) { btkb_destroy(pipe_); pipe_ = nullptr; }
  ck(btkb_create(&c, &pipe_));
  ck(btkb_set_prototypes(pipe_, a0->prototype().data(), syn.enabled ? syn.prototype.data() : nullptr, (int)a0->prototype().size()));
}

// Collect the upstream graph, run the whole utterance on the GPU once, cache the results.
void SubbandBeamformer::run_graph(const PostFilterConfig& pf, const SynthesisConfig& syn) {
  if (channels_.empty()) throw j_error("set_channel() has not been called\n");
  const unsigned C = (unsigned)channels_.size();
  std::vector<const SampleFeature*> srcs(C);
  unsigned n = 0;
  wpe_.reset();
  for (unsigned c = 0; c < C; c++) {
    auto* ab = bank_behind(channels_[c], &wpe_);
    if (!ab) throw j_error("SubbandBeamformer: the GPU engine needs OverSampledDFTAnalysisBank channels (directly or behind MultiChannelWPEDereverberationFeature)");
    if (ab->M() != fftLen_) throw jdimension_error("channel %d: inconsistent FFT length (%d vs. %d)", c, ab->M(), fftLen_);
    srcs[c] = dynamic_cast<const SampleFeature*>(ab->source().get());
    if (!srcs[c]) throw j_error("SubbandBeamformer: the GPU engine needs SampleFeature sources");
    if (c == 0) n = srcs[c]->samplesN();
    else if (srcs[c]->samplesN() != n) throw jdimension_error("channel %d: %d samples, channel 0: %d", c, srcs[c]->samplesN(), n);
  }
  if (n == 0) { T_ = 0; nb_ = 0; realized_ = true; pf_used_ = pf; syn_used_ = syn; return; }
  if (syn.enabled && (syn.M != fftLen_)) throw jdimension_error("synthesis bank: inconsistent FFT length (%d vs. %d)", syn.M, fftLen_);
  live_ = false;
  const bool chunked = chunk_blocks_ > 0 && !wpe_ && C <= 8 && stream_capable_();
  {
    auto* a0 = bank_behind(channels_[0], nullptr);
    D_ = fftLen_ >> a0->r();
  }
  ensure_pipeline_(pf, syn, chunked ? std::min<unsigned>(n, (unsigned)chunk_blocks_ * D_) : n);
  configure_weights_(pipe_);
  if (pf.enabled && pf.kind != BTKB_PF_ZELINSKI) {
    if (!pf.coherence) throw j_error("McCowanPostFilter:  construct/set a noise coherence matrix\n");  // postfilter.cc:828-830
    if (pf.coherence->chanN() != C) throw jdimension_error("noise coherence matrix is %d x %d but the beamformer has %d channels\n", pf.coherence->chanN(), pf.coherence->chanN(), C);
    pf.coherence->push_coherence(pipe_);
  }
  if (chunked) {
    // totals in closed form (modulated.cc:246-264, 418-469): T = ceil(n / D) - laN + pd_A, blocks = T - pd_S
    auto* a0 = bank_behind(channels_[0], nullptr);
    const int R = 1 << a0->r(), mm = (int)a0->m(), dct = (int)a0->dct();
    int pdA, laN = 0, pdS;
    if (dct == 1) { pdA = mm * R - 1; pdS = mm * R - 1; } else if (dct == 2) { pdA = mm * R - 1; laN = mm * R / 2 - 1; pdS = mm * R / 2; } else { pdA = 2 * mm - 1; pdS = 2 * mm - 1; }
    T_ = (int)((n + D_ - 1) / D_) - laN + pdA; nb_ = std::max(T_ - pdS, 0);
    ck(btkb_stream_begin(pipe_, 1));
    srcs_ = srcs; n_total_ = n; pos_ = 0; T_ready_ = nb_ready_ = chunk_t0_ = 0; weights_dirty_ = false;
    const unsigned K = fftLen_ / 2 + 1;
    Y_.assign((size_t)T_ * K, std::complex<float>(0.f, 0.f));
    time_.assign(syn.enabled ? (size_t)nb_ * D_ : 0, 0.f);
    pfw_.assign(pf.enabled ? (size_t)T_ * K : 0, 0.f);
    W_.resize((size_t)K * C);
    ck(btkb_get_weights(pipe_, reinterpret_cast<float*>(W_.data())));
    haveX_ = false;
    realized_ = true; live_ = true; pf_used_ = pf; syn_used_ = syn;
    return;
  }
  std::vector<float> x((size_t)C * n);
  for (unsigned c = 0; c < C; c++) std::memcpy(&x[(size_t)c * n], srcs[c]->samples().data(), sizeof(float) * n);
  ck(btkb_submit(pipe_, x.data(), 1, (int)n, nullptr));
  if (wpe_) {
    if (!wpe_->estimated()) throw jinitialization_error("Call SingleChannelWPEDereverberationFeature::estimate_filter()\n");
    if (wpe_->channelsN() != C) throw jdimension_error("the dereverberator has %d channels, the beamformer %d", (int)wpe_->channelsN(), (int)C);
    // the filters of the earlier estimate_filter() call are applied to the audio the sources hold NOW, whatever they were estimated on
    // (MultiChannelWPEDereverberationFeature::next -> calc_every_channel_output, dereverberation.cc:441-497, 713-728)
    ck(btkb_run_analysis(pipe_));
    const std::vector<float> G = wpe_->filters();
    ck(btkb_set_wpe_filter(pipe_, 1, G.data()));
    ck(btkb_apply_wpe(pipe_));
    ck(btkb_run_beamformer(pipe_, syn.enabled ? 1 : 0));
  } else {
    ck(btkb_run(pipe_, syn.enabled ? 1 : 0));
  }
  T_ = btkb_num_frames(pipe_); nb_ = btkb_num_blocks(pipe_);
  const unsigned K = fftLen_ / 2 + 1;
  Y_.resize((size_t)T_ * K);
  ck(btkb_fetch_subband(pipe_, reinterpret_cast<float*>(Y_.data())));
  if (syn.enabled) { time_.resize((size_t)nb_ * (fftLen_ >> syn.r)); ck(btkb_fetch_time(pipe_, time_.data())); }
  if (pf.enabled) { pfw_.resize((size_t)T_ * K); ck(btkb_get_postfilter_weights(pipe_, pfw_.data())); }
  W_.resize((size_t)K * C);
  ck(btkb_get_weights(pipe_, reinterpret_cast<float*>(W_.data())));
  haveX_ = false;
  realized_ = true; pf_used_ = pf; syn_used_ = syn;
}

// Chunked realisation: hand the next chunk_blocks_ blocks of every channel to the GPU and append what comes back.
void SubbandBeamformer::advance_() {
  if (!live_ || pos_ >= n_total_) return;
  if (weights_dirty_) { configure_weights_(pipe_); weights_dirty_ = false; }   // btkb_set_delays* between chunks keeps the adaptive state
  const unsigned C = chanN(), K = fftLen_ / 2 + 1;
  const size_t want = (size_t)chunk_blocks_ * D_;
  const size_t nc = std::min<size_t>(want, n_total_ - pos_);
  const bool final_chunk = pos_ + nc >= n_total_;
  xchunk_.resize((size_t)C * nc);
  for (unsigned c = 0; c < C; c++) std::memcpy(&xchunk_[(size_t)c * nc], srcs_[c]->samples().data() + pos_, sizeof(float) * nc);
  ck(btkb_stream_submit(pipe_, xchunk_.data(), (int)nc, nullptr, final_chunk ? 1 : 0, syn_used_.enabled ? 1 : 0));
  pos_ += nc;
  const int Tl = btkb_num_frames(pipe_), nbl = btkb_num_blocks(pipe_);
  chunk_t0_ = T_ready_;
  if (Tl > 0) {
    if (T_ready_ + Tl > T_) throw j_error("SubbandBeamformer: chunked realisation produced more frames than the closed form predicts");
    ck(btkb_fetch_subband(pipe_, reinterpret_cast<float*>(&Y_[(size_t)T_ready_ * K])));
    if (pf_used_.enabled) ck(btkb_get_postfilter_weights(pipe_, &pfw_[(size_t)T_ready_ * K]));
    T_ready_ += Tl;
  }
  if (syn_used_.enabled && nbl > 0) {
    if (nb_ready_ + nbl > nb_) throw j_error("SubbandBeamformer: chunked realisation produced more blocks than the closed form predicts");
    ck(btkb_fetch_time(pipe_, &time_[(size_t)nb_ready_ * D_]));
    nb_ready_ += nbl;
  }
  haveX_ = false;
}
void SubbandBeamformer::ensure_frames(int t) {
  if (!live_) return;
  while (T_ready_ <= t && pos_ < n_total_) advance_();
}
void SubbandBeamformer::ensure_blocks(int b) {
  if (!live_) return;
  while (nb_ready_ <= b && pos_ < n_total_) advance_();
}

const cplx* SubbandBeamformer::next(int frame_no) {
  if (frame_no == frame_no_) return vector_.data();
  if (!realized_) run_graph(PostFilterConfig(), SynthesisConfig());
  if (frame_no_ + 1 >= T_) { is_end_ = true; throw jiterator_error("end of samples!"); }
  ensure_frames(frame_no_ + 1);
  increment_();
  const unsigned K = fftLen_ / 2 + 1;
  const std::complex<float>* y = &Y_[(size_t)frame_no_ * K];
  for (unsigned k = 0; k < K; k++) vector_[k] = cplx(y[k].real(), y[k].imag());
  for (unsigned k = 1; k < fftLen_ / 2; k++) vector_[fftLen_ - k] = std::conj(vector_[k]);   // beamformer.cc:1142-1149
  return vector_.data();
}

SnapShotArrayPtr SubbandBeamformer::snapshot_array() {
  if (!snap_) snap_ = std::make_shared<SnapShotArray>(fftLen_, chanN());
  if (realized_ && pipe_ && frame_no_ >= 0 && frame_no_ < T_) {
    const unsigned K = fftLen_ / 2 + 1, C = chanN();
    const int tl = live_ ? frame_no_ - chunk_t0_ : frame_no_;   // chunked realisation: the device holds the snapshots of the current chunk
    if (!haveX_) { X_.resize((size_t)btkb_num_frames(pipe_) * C * K); ck(btkb_fetch_snapshots(pipe_, reinterpret_cast<float*>(X_.data()))); haveX_ = true; }
    if (tl < 0 || (size_t)(tl + 1) * C * K > X_.size()) return snap_;
    std::vector<cplx> full(fftLen_);
    for (unsigned c = 0; c < C; c++) {
      const std::complex<float>* x = &X_[((size_t)tl * C + c) * K];
      for (unsigned k = 0; k < K; k++) full[k] = cplx(x[k].real(), x[k].imag());
      for (unsigned k = 1; k < fftLen_ / 2; k++) full[fftLen_ - k] = std::conj(full[k]);
      snap_->set_samples(full.data(), c);
    }
    snap_->update();
  }
  return snap_;
}

std::vector<cplx> SubbandBeamformer::get_weights(unsigned fbinX) {
  const unsigned K = fftLen_ / 2 + 1, C = chanN();
  if (fbinX >= K) throw jindex_error("fbinX %d must be <= %d", fbinX, fftLen_ / 2);
  if (W_.size() != (size_t)K * C) {
    // weights do not depend on the samples: realise them on a one-block dummy batch
    btkb_pipeline* keep = pipe_; pipe_ = nullptr;
    auto* a0 = channels_.empty() ? nullptr : dynamic_cast<OverSampledDFTAnalysisBank*>(channels_[0].get());
    if (!a0) throw j_error("call set_channel() before asking for weights");
    ensure_pipeline_(PostFilterConfig(), SynthesisConfig(), fftLen_);
    configure_weights_(pipe_);
    W_.resize((size_t)K * C);
    ck(btkb_get_weights(pipe_, reinterpret_cast<float*>(W_.data())));
    btkb_destroy(pipe_); pipe_ = keep;
  }
  std::vector<cplx> w(C);
  for (unsigned c = 0; c < C; c++) w[c] = cplx(W_[(size_t)fbinX * C + c].real(), W_[(size_t)fbinX * C + c].imag());
  return w;
}

void SubbandBeamformer::fetch_static_weights_(std::vector<std::complex<float>>& W, std::vector<std::complex<float>>& WL) {
  const unsigned K = fftLen_ / 2 + 1, C = chanN();
  if (channels_.empty() || !bank_behind(channels_[0], nullptr)) throw j_error("call set_channel() before asking for weights");
  btkb_pipeline* keep = pipe_; pipe_ = nullptr;
  try {
    ensure_pipeline_(PostFilterConfig(), SynthesisConfig(), fftLen_);
    configure_weights_(pipe_);
    W.resize((size_t)K * C); WL.resize((size_t)K * C);
    ck(btkb_get_weights(pipe_, reinterpret_cast<float*>(W.data())));
    ck(btkb_get_sidelobe_weights(pipe_, reinterpret_cast<float*>(WL.data())));
  } catch (...) {
    if (pipe_) btkb_destroy(pipe_);
    pipe_ = keep;
    throw;
  }
  btkb_destroy(pipe_); pipe_ = keep;
}

// ---- SubbandDS
SubbandDS::SubbandDS(unsigned fftLen, bool hbs, const std::string& nm, int kind) : SubbandBeamformer(fftLen, hbs, kind, nm) {}
void SubbandDS::clear_channel() { SubbandBeamformer::clear_channel(); have_delays_ = false; W_.clear(); }
void SubbandDS::calc_array_manifold_vectors(double samplerate, const std::vector<double>& delays) {
  if (delays.size() != chanN())  // beamformer.cc:504-506
    throw jdimension_error("Number of delays does not match number of channels (%d vs. %d).\n", (int)delays.size(), (int)chanN());
  samplerate_ = samplerate; delays_ = delays; have_delays_ = true; NC_ = 1; delaysJ_.clear(); W_.clear(); invalidate_weights_();   // alloc_bfweight_(1, 1)
}
void SubbandDS::calc_array_manifold_vectors_n(double samplerate, const std::vector<double>& delaysT, const std::vector<double>& delaysJ, unsigned NC) {
  const unsigned C = chanN();
  if (NC < 2 || NC > C) throw jdimension_error("1 < the number of constraints %d <= the number of sensors %d.\n", (int)NC, (int)C);   // beamformer.cc:592-594
  if (delaysJ.size() != (size_t)(NC - 1) * C) throw jdimension_error("delays of the interference signals must be %d x %d\n", (int)(NC - 1), (int)C);
  calc_array_manifold_vectors(samplerate, delaysT);
  delaysJ_ = delaysJ; NC_ = NC;
}
void SubbandDS::set_delays_(btkb_pipeline* p) {
  if (NC_ > 1) ck(btkb_set_delays_lcmv(p, 1, (int)NC_, delays_.data(), delaysJ_.data()));
  else ck(btkb_set_delays(p, 1, delays_.data()));
}
void SubbandDS::configure_weights_(btkb_pipeline* p) {
  require_weights_(have_delays_, "call calc_array_manifold_vectorsX() once\n");  // beamformer.cc:1098-1100
  set_delays_(p);
}

// ---- SubbandGSC
SubbandGSC::SubbandGSC(unsigned fftLen, bool hbs, const std::string& nm) : SubbandDS(fftLen, hbs, nm, BTKB_BF_GSC) {}
void SubbandGSC::calc_gsc_weights_n(double samplerate, const std::vector<double>& delaysT, const std::vector<double>& delaysJ, unsigned NC) {
  const unsigned C = chanN();
  if (NC < 2 || NC > C) throw jdimension_error("1 < the number of constraints %d <= the number of sensors %d.\n", (int)NC, (int)C);   // beamformer.cc:592-594
  if (delaysJ.size() != (size_t)(NC - 1) * C) throw jdimension_error("delays of the interference signals must be %d x %d\n", (int)(NC - 1), (int)C);
  calc_array_manifold_vectors(samplerate, delaysT);
  delaysJ_ = delaysJ; NC_ = NC; wa_.clear(); have_wa_ = false;
}
void SubbandGSC::set_quiescent_weights_f(unsigned fbinX, const std::vector<cplx>& srcWq) {   // beamformer.cc:1318-1324
  const unsigned C = chanN(), K = fftLen_ / 2 + 1;
  if (C == 0) throw j_error("set_channel() has not been called\n");
  if (srcWq.size() != C) throw jdimension_error("the quiescent vector must have %d elements but it has %d\n", (int)C, (int)srcWq.size());
  if (fbinX >= fftLen_) throw jindex_error("fbinX %d must be less than %d\n", fbinX, fftLen_);
  // alloc_bfweight_(1, 1): a fresh, zeroed BeamformerWeights (beamformer.cc:1082-1092); bins above M/2 are never read by next()
  wq_explicit_.assign((size_t)K * C, std::complex<float>(0, 0));
  if (fbinX < K) for (unsigned c = 0; c < C; c++) wq_explicit_[(size_t)fbinX * C + c] = std::complex<float>((float)srcWq[c].real(), (float)srcWq[c].imag());
  have_wq_explicit_ = true; have_delays_ = false; NC_ = 1; delaysJ_.clear(); wa_.clear(); have_wa_ = false; W_.clear(); invalidate_();
}
bool SubbandGSC::write_fir_coeff(const std::string& fn, unsigned winType) {   // beamformer.cc:775-828, 1364-1371
  if (!have_delays_ && !have_wq_explicit_) { fprintf(stderr, "call calc_array_manifold_vectorsX() once\n"); return false; }
  const unsigned M = fftLen_, M2 = M / 2, C = chanN();
  std::vector<std::complex<float>> W, WL;
  fetch_static_weights_(W, WL);
  FILE* fp = fopen(fn.c_str(), "w");
  if (!fp) { printf("could not open %s\n", fn.c_str()); return false; }
  fprintf(fp, "%d %d\n", (int)C, (int)M);
  std::vector<double> window(M);   // get_window (modulated.cc:47-73): 0 rectangle, 2 Hanning, otherwise Hamming
  for (unsigned i = 0; i < M; i++) {
    const double ph = 2.0 * M_PI * i / (double)(M - 1);
    window[i] = (winType == 0) ? 1.0 : (winType == 2 ? 0.5 * (1.0 - std::cos(ph)) : 0.54 - 0.46 * std::cos(ph));
  }
  std::vector<cplx> spec(M), twid(M);
  for (unsigned i = 0; i < M; i++) twid[i] = std::polar(1.0, 2.0 * M_PI * i / (double)M);
  for (unsigned c = 0; c < C; c++) {
    for (unsigned f = 0; f <= M2; f++) {
      const cplx wq(W[(size_t)f * C + c].real(), W[(size_t)f * C + c].imag()), wl(WL[(size_t)f * C + c].real(), WL[(size_t)f * C + c].imag());
      const cplx val = std::polar(1.0, M_PI * (double)(f + 1)) * std::conj(wq - wl);   // shift by fftLen/2
      spec[f] = val;
      if (f > 0 && f < M2) spec[M - f] = std::conj(val);
    }
    for (unsigned n = 0; n < M; n++) {   // gsl_fft_complex_radix2_inverse: (1/M) sum_f spec[f] e^{+2 pi i f n / M}; the real part is written
      cplx acc(0, 0);
      for (unsigned f = 0; f < M; f++) acc += spec[f] * twid[(size_t)f * n % M];
      fprintf(fp, "%e ", window[n] * acc.real() / (double)M);
    }
    fprintf(fp, "\n");
  }
  fclose(fp);
  return true;
}
void SubbandGSC::set_active_weights_f(unsigned fbinX, const std::vector<double>& packed) {  // beamformer.cc:729-748,1365-1372
  require_weights_(have_delays_ || have_wq_explicit_, "call calc_gsc_weights_x() once\n");
  const unsigned C = chanN(), K = fftLen_ / 2 + 1, NA = C - NC_;
  if (packed.size() != 2 * NA) throw jdimension_error("the size of an active weight vector must be %d but it is %d\n", (int)(2 * NA), (int)packed.size());
  if (fbinX >= fftLen_) throw jdimension_error("Must be a frequency bin %d < the length of FFT %d\n", fbinX, fftLen_);
  if (fbinX >= K) return;  // mirrored bins are never used by next()
  if (wa_.size() != (size_t)K * NA) wa_.assign((size_t)K * NA, std::complex<float>(0, 0));
  for (unsigned i = 0; i < NA; i++) wa_[(size_t)fbinX * NA + i] = std::complex<float>((float)packed[2 * i], (float)packed[2 * i + 1]);
  have_wa_ = true; invalidate_();
}
void SubbandGSC::zero_active_weights() { require_weights_(have_delays_ || have_wq_explicit_, "call calc_gsc_weights_x() once\n"); wa_.clear(); have_wa_ = false; invalidate_(); }
void SubbandGSC::configure_weights_(btkb_pipeline* p) {
  require_weights_(have_delays_ || have_wq_explicit_, "call calc_gsc_weights_X() once\n");  // beamformer.cc:1262-1264
  if (have_wq_explicit_) ck(btkb_set_weights(p, 1, reinterpret_cast<const float*>(wq_explicit_.data())));
  else set_delays_(p);
  if (have_wa_) ck(btkb_set_active_weights(p, 1, reinterpret_cast<const float*>(wa_.data())));
}

// ---- SubbandGSCRLS (the reference's C++ class)
SubbandGSCRLS::SubbandGSCRLS(unsigned fftLen, bool hbs, float myu, float sigma2, const std::string& nm) : SubbandGSC(fftLen, hbs, nm), mu_(myu), sigma2_(sigma2) {
  kind_ = BTKB_BF_GSC_RLS_CPP;
}
void SubbandGSCRLS::tune_config_(btkb_config& c) {
  c.rls_cpp.mu = mu_; c.rls_cpp.sigma2 = sigma2_; c.rls_cpp.init_sigma2 = init_sigma2_; c.rls_cpp.alpha = alpha_; c.rls_cpp.qctype = qctype_; c.rls_cpp.update = update_ ? 1 : 0;
}
void SubbandGSCRLS::configure_weights_(btkb_pipeline* p) {
  require_weights_(have_delays_, "call calc_gsc_weights_x() once\n");   // beamformer.cc:1518-1519
  if (!have_pz_) throw j_error("set the precision matrix with init_precision_matrix() or set_precision_matrix()\n");   // :1520-1521
  if (NC_ > 1) throw j_error("SubbandGSCRLS: the GPU kernel implements one linear constraint\n");
  if (chunk_blocks_ > 0) chunk_blocks_ = 0;   // whole utterances only (btkb_stream_begin refuses this kind)
  ck(btkb_set_delays(p, 1, delays_.data()));
}

// ---- SubbandGSCLMS
SubbandGSCLMS::SubbandGSCLMS(unsigned fftLen, const LmsConfig& cfg, const std::string& nm) : SubbandDS(fftLen, false, nm, BTKB_BF_GSC_LMS) { lms_ = cfg; }
std::vector<std::complex<float>> SubbandGSCLMS::active_weights() {
  if (!realized_ || !pipe_) throw j_error("run the beamformer first");
  std::vector<std::complex<float>> wa((size_t)(fftLen_ / 2 + 1) * (chanN() - 1));
  ck(btkb_get_active_weights(pipe_, reinterpret_cast<float*>(wa.data())));
  return wa;
}
int SubbandGSCLMS::total_updates() {
  if (!realized_ || !pipe_) return 0;
  double st[3]; ck(btkb_fetch_stats(pipe_, st));
  return (int)st[2];
}

// ---- SubbandGSCRLSNative
SubbandGSCRLSNative::SubbandGSCRLSNative(unsigned fftLen, const RlsConfig& cfg, const std::string& nm) : SubbandDS(fftLen, false, nm, BTKB_BF_GSC_RLS) { rls_ = cfg; }
std::vector<std::complex<float>> SubbandGSCRLSNative::active_weights() {
  if (!realized_ || !pipe_) throw j_error("run the beamformer first");
  std::vector<std::complex<float>> wa((size_t)(fftLen_ / 2 + 1) * (chanN() - 1));
  ck(btkb_get_active_weights(pipe_, reinterpret_cast<float*>(wa.data())));
  return wa;
}
int SubbandGSCRLSNative::total_updates() {
  if (!realized_ || !pipe_) return 0;
  double st[3]; ck(btkb_fetch_stats(pipe_, st));
  return (int)st[2];
}

// ---- SubbandSOSNative
SubbandSOSNative::SubbandSOSNative(unsigned fftLen, const std::string& nm) : SubbandDS(fftLen, false, nm, BTKB_BF_DS) {}
SubbandSOSNative::~SubbandSOSNative() { if (stats_) btkb_destroy(stats_); }
void SubbandSOSNative::clear_channel() { SubbandDS::clear_channel(); if (stats_) { btkb_destroy(stats_); stats_ = nullptr; } stats_cap_ = 0; have_wsos_ = false; }
void SubbandSOSNative::reset_stats() { if (stats_) ck(btkb_sos_reset_stats(stats_)); }   // pybeamformer.py:1209-1213
unsigned SubbandSOSNative::stage_(double samplerate) {
  if (channels_.empty()) throw j_error("set_channel() has not been called\n");
  const unsigned C = chanN();
  std::vector<const SampleFeature*> srcs(C); unsigned n = 0;
  for (unsigned c = 0; c < C; c++) {
    auto* ab = dynamic_cast<OverSampledDFTAnalysisBank*>(channels_[c].get());
    if (!ab) throw j_error("SubbandSOS: the GPU engine needs OverSampledDFTAnalysisBank channels");
    srcs[c] = dynamic_cast<const SampleFeature*>(ab->source().get());
    if (!srcs[c]) throw j_error("SubbandSOS: the GPU engine needs SampleFeature sources");
    if (c == 0) n = srcs[c]->samplesN();
    else if (srcs[c]->samplesN() != n) throw jdimension_error("channel %d: %d samples, channel 0: %d", c, srcs[c]->samplesN(), n);
  }
  if (n == 0) throw jiterator_error("end of samples!");
  samplerate_ = samplerate;
  if (!stats_ || n > stats_cap_) {
    if (stats_) throw j_error("SubbandSOS: a later utterance (%d samples) is longer than the first one (%d); call reset_stats() and start with the longest", n, stats_cap_);
    auto* a0 = dynamic_cast<OverSampledDFTAnalysisBank*>(channels_[0].get());
    btkb_config c; btkb_default_config(&c);
    c.channels = (int)C; c.fft_len = (int)fftLen_; c.m = (int)a0->m(); c.r = (int)a0->r(); c.delay_compensation_type = (int)a0->dct();
    c.samplerate = (float)samplerate; c.beamformer = BTKB_BF_DS; c.max_utterances = 1; c.max_samples = (int)n;
    ck(btkb_create(&c, &stats_));
    ck(btkb_set_prototypes(stats_, a0->prototype().data(), nullptr, (int)a0->prototype().size()));
    stats_cap_ = n;
  }
  std::vector<float> x((size_t)C * n);
  for (unsigned c = 0; c < C; c++) std::memcpy(&x[(size_t)c * n], srcs[c]->samples().data(), sizeof(float) * n);
  ck(btkb_submit(stats_, x.data(), 1, (int)n, nullptr));
  ck(btkb_run_analysis(stats_));
  return n;
}
void SubbandSOSNative::accu_stats_from_label(double samplerate, const std::vector<double>& labels, double thr) {
  if (labels.empty() || labels.size() % 2) throw jdimension_error("target_labs must be a list of (start, end) pairs");
  stage_(samplerate);
  ck(btkb_sos_accumulate_from_label(stats_, labels.data(), (int)(labels.size() / 2), (float)thr));
}
void SubbandSOSNative::accu_stats_from_tfmask(double samplerate, const std::vector<float>& mt, const std::vector<float>& mj, unsigned rows, unsigned cols, double thr) {
  const unsigned K = fftLen_ / 2 + 1;
  if (cols < K || mt.size() != (size_t)rows * cols || mj.size() != mt.size()) throw jdimension_error("TF masks must be [frames][>= %d] and of equal shape", K);
  stage_(samplerate);
  const unsigned T = (unsigned)btkb_num_frames(stats_);
  if (rows < T) throw jindex_error("index %d is out of bounds for axis 0 with size %d", rows, rows);   // mask_t[frame_no] (pybeamformer.py:1151)
  std::vector<float> a((size_t)rows * K), b((size_t)rows * K);
  for (unsigned t = 0; t < rows; t++)
    for (unsigned k = 0; k < K; k++) { a[(size_t)t * K + k] = mt[(size_t)t * cols + k]; b[(size_t)t * K + k] = mj[(size_t)t * cols + k]; }
  ck(btkb_sos_accumulate_from_tfmask(stats_, a.data(), b.data(), (int)rows, (float)thr));
}
void SubbandSOSNative::calc_weights(int kind, double gamma, int ref_micx, double offset) {
  if (!stats_) throw j_error("No target signal SOS");   // pybeamformer.py:1270-1273
  ck(btkb_sos_calc_weights(stats_, kind, gamma, ref_micx, offset));
  const unsigned K = fftLen_ / 2 + 1, C = chanN();
  wsos_.resize((size_t)K * C);
  ck(btkb_get_weights(stats_, reinterpret_cast<float*>(wsos_.data())));
  have_wsos_ = true; W_.clear(); invalidate_();
}
std::vector<double> SubbandSOSNative::frame_counts() {
  if (!stats_) throw j_error("no statistics accumulated");
  std::vector<double> c((size_t)(fftLen_ / 2 + 1) * 2);
  ck(btkb_sos_get_stats(stats_, nullptr, nullptr, c.data()));
  return c;
}
void SubbandSOSNative::configure_weights_(btkb_pipeline* p) {
  // before calc_beamformer_weights the reference's wqH is all ones (pybeamformer.py:1044)
  const unsigned K = fftLen_ / 2 + 1, C = chanN();
  if (!have_wsos_) { wsos_.assign((size_t)K * C, std::complex<float>(1.f, 0.f)); }
  ck(btkb_set_weights(p, 1, reinterpret_cast<const float*>(wsos_.data())));
}

// ---- SubbandMVDR
SubbandMVDR::SubbandMVDR(unsigned fftLen, bool hbs, const std::string& nm) : SubbandDS(fftLen, hbs, nm, BTKB_BF_MVDR) {}
void SubbandMVDR::clear_channel() { SubbandDS::clear_channel(); R_.clear(); wmvdr_.clear(); load_f_.clear(); div_f_.clear(); have_R_ = have_w_ = diffuse_ = smi_ = false; inv_ = InvSource(); }
void SubbandMVDR::set_diagonal_looading(unsigned fbinX, float w) {   // beamformer.cc:2525-2535
  if (!have_R_) throw j_error("Construct first a noise covariance matrix\n");
  const unsigned K = fftLen_ / 2 + 1;
  if (fbinX >= K) throw jindex_error("fbinX %d must be <= %d", fbinX, fftLen_ / 2);
  if (load_f_.size() != K) load_f_.assign(K, 0.0);
  load_f_[fbinX] += (double)w; have_w_ = false; W_.clear(); invalidate_();
}
void SubbandMVDR::divide_nondiagonal_elements(unsigned fbinX, float mu) {   // beamformer.cc:2589-2599
  if (!have_R_) throw j_error("Construct first a noise covariance matrix\n");
  const unsigned K = fftLen_ / 2 + 1;
  if (fbinX >= K) throw jindex_error("fbinX %d must be <= %d", fbinX, fftLen_ / 2);
  if (div_f_.size() != K) div_f_.assign(K, 1.0);
  div_f_[fbinX] *= 1.0 + (double)mu; have_w_ = false; W_.clear(); invalidate_();
}
void SubbandMVDR::divide_all_nondiagonal_elements(float mu) {   // beamformer.h:357-360
  for (unsigned f = 0; f <= fftLen_ / 2; f++) divide_nondiagonal_elements(f, mu);
}
bool SubbandMVDR::set_noise_spatial_spectral_matrix(unsigned fbinX, const std::vector<cplx>& Rnn) {  // beamformer.cc:2410-2433
  const unsigned C = chanN(), K = fftLen_ / 2 + 1;
  if (Rnn.size() != (size_t)C * C) { fprintf(stderr, "The number of the rows of the matrix must be %d\n", C); return false; }
  if (fbinX >= K) throw jindex_error("fbinX %d must be <= %d", fbinX, fftLen_ / 2);
  if (R_.size() != (size_t)K * C * C) R_.assign((size_t)K * C * C, std::complex<float>(0, 0));
  for (size_t i = 0; i < (size_t)C * C; i++) R_[(size_t)fbinX * C * C + i] = std::complex<float>((float)Rnn[i].real(), (float)Rnn[i].imag());
  if (load_f_.size() == K) load_f_[fbinX] = 0.0;   // the bin's matrix is replaced, earlier edits of it are gone
  if (div_f_.size() == K) div_f_[fbinX] = 1.0;
  have_R_ = true; diffuse_ = false; smi_ = false; mu_ = 0.0; have_w_ = false; invalidate_();
  return true;
}
bool SubbandMVDR::set_diffuse_noise_model(const std::vector<double>& mpos, double samplerate, double sspeed) {  // beamformer.cc:2442-2509
  if (mpos.size() != (size_t)chanN() * 3) { fprintf(stderr, "The number of microphones must be %d but it is %d\n", chanN(), (int)(mpos.size() / 3)); return false; }
  mpos_ = mpos; samplerate_ = samplerate; sspeed_ = sspeed; have_R_ = true; diffuse_ = true; smi_ = false; mu_ = 0.0; have_w_ = false; invalidate_();
  load_f_.clear(); div_f_.clear();
  return true;
}
void SubbandMVDR::set_all_diagonal_loading(double w) {  // beamformer.cc:2511-2523 (R += w I, cumulative like the reference)
  if (!have_R_) throw j_error("Construct first a noise covariance matrix\n");
  mu_ += (double)(float)w; have_w_ = false; invalidate_();
}
bool SubbandMVDR::calc_mvdr_weights(double samplerate, double dThreshold, bool calc_inverse_matrix) {  // beamformer.cc:2350-2402
  if (!have_R_) throw jallocation_error("Set a spatial spectral matrix before calling calc_mvdr_weights()\n");
  require_weights_(have_delays_, "call calc_array_manifold_vectorsX() once\n");
  if (calc_inverse_matrix) {   // pseudoinverse(R_[f], invR_[f], dThreshold) now; later calls with calc_inverse_matrix = false reuse it
    inv_.R = R_; inv_.diffuse = diffuse_; inv_.mpos = mpos_; inv_.sspeed = sspeed_; inv_.mu = mu_; inv_.load_f = load_f_; inv_.div_f = div_f_; inv_.valid = true;
    dthreshold_ = (float)dThreshold;
  } else if (!inv_.valid) {
    throw j_error("calc_mvdr_weights(calc_inverse_matrix=False) needs an earlier call that computed the inverse matrices\n");   // the reference would read unset invR_
  }
  samplerate_ = samplerate; have_w_ = true; W_.clear(); invalidate_();
  return true;
}
int SubbandMVDR::accumulate_noise_covariance(double samplerate, double start, double end, double thr) {
  // pass 1 of pybeamformer.py:948-1000 on the GPU: analysis + masked covariance; the statistics stay on the host as R
  const unsigned C = chanN(), K = fftLen_ / 2 + 1;
  std::vector<const SampleFeature*> srcs(C); unsigned n = 0;
  for (unsigned c = 0; c < C; c++) {
    auto* ab = dynamic_cast<OverSampledDFTAnalysisBank*>(channels_[c].get());
    if (!ab) throw j_error("SubbandMVDR: the GPU engine needs OverSampledDFTAnalysisBank channels");
    srcs[c] = dynamic_cast<const SampleFeature*>(ab->source().get());
    if (!srcs[c]) throw j_error("SubbandMVDR: the GPU engine needs SampleFeature sources");
    n = srcs[0]->samplesN();
  }
  samplerate_ = samplerate;
  ensure_pipeline_(PostFilterConfig(), SynthesisConfig(), n);
  std::vector<float> x((size_t)C * n);
  for (unsigned c = 0; c < C; c++) std::memcpy(&x[(size_t)c * n], srcs[c]->samples().data(), sizeof(float) * n);
  ck(btkb_submit(pipe_, x.data(), 1, (int)n, nullptr));
  ck(btkb_run_analysis(pipe_));
  double lab[2] = {start, end};
  ck(btkb_accumulate_covariance(pipe_, lab, (float)thr));
  std::vector<std::complex<float>> Rn((size_t)K * C * C);
  ck(btkb_get_covariance(pipe_, reinterpret_cast<float*>(Rn.data())));
  R_ = Rn; have_R_ = true; diffuse_ = false; smi_ = true; mu_ = 0.0; have_w_ = false; invalidate_();
  load_f_.clear(); div_f_.clear();
  return 0;
}
void SubbandMVDR::configure_weights_(btkb_pipeline* p) {
  require_weights_(have_delays_, "call calc_array_manifold_vectorsX() once\n");
  if (!have_w_) throw j_error("call calc_mvdr_weights() once\n");  // beamformer.cc:2544-2546
  if (NC_ > 1) throw j_error("SubbandMVDR: the GPU MVDR solve takes the delay-and-sum manifold (calc_array_manifold_vectors), not LCMV weights\n");
  ck(btkb_set_delays(p, 1, delays_.data()));
  // the matrix of the last calc_mvdr_weights(calc_inverse_matrix = true)
  const std::vector<std::complex<float>>& R_ = inv_.R; const bool diffuse_ = inv_.diffuse; const std::vector<double>& mpos_ = inv_.mpos;
  const std::vector<double>& load_f_ = inv_.load_f; const std::vector<double>& div_f_ = inv_.div_f; const double sspeed_ = inv_.sspeed, mu_ = inv_.mu;
  if (diffuse_) ck(btkb_set_diffuse_noise_model(p, 1, mpos_.data(), (float)sspeed_));
  else ck(btkb_set_noise_covariance(p, 1, reinterpret_cast<const float*>(R_.data())));
  if (!load_f_.empty() || !div_f_.empty()) {   // per-bin edits (set_diagonal_looading, divide_nondiagonal_elements): O(K C^2) parameter edits
    const unsigned K = fftLen_ / 2 + 1, C = chanN();
    std::vector<std::complex<float>> R((size_t)K * C * C);
    ck(btkb_get_covariance(p, reinterpret_cast<float*>(R.data())));
    for (unsigned f = 0; f < K; f++)
      for (unsigned i = 0; i < C; i++)
        for (unsigned j = 0; j < C; j++) {
          std::complex<float>& v = R[((size_t)f * C + i) * C + j];
          if (i == j) { if (!load_f_.empty()) v += std::complex<float>((float)load_f_[f], 0.f); }
          else if (!div_f_.empty()) v = std::complex<float>((float)(v.real() / div_f_[f]), (float)(v.imag() / div_f_[f]));
        }
    ck(btkb_set_noise_covariance(p, 1, reinterpret_cast<const float*>(R.data())));
  }
  ck(btkb_calc_mvdr_weights_ex(p, (float)mu_, dthreshold_));
}
void SubbandMVDRGSC::set_active_weights_f(unsigned fbinX, const std::vector<double>& packed) {
  const unsigned C = chanN(), K = fftLen_ / 2 + 1;
  if (packed.size() != 2 * (C - 1)) throw jdimension_error("the size of an active weight vector must be %d but it is %d\n", (int)(2 * (C - 1)), (int)packed.size());
  if (fbinX >= K) return;
  if (wa_.size() != (size_t)K * (C - 1)) wa_.assign((size_t)K * (C - 1), std::complex<float>(0, 0));
  bool nz = false;
  for (unsigned i = 0; i < C - 1; i++) { wa_[(size_t)fbinX * (C - 1) + i] = std::complex<float>((float)packed[2 * i], (float)packed[2 * i + 1]); nz |= packed[2 * i] != 0 || packed[2 * i + 1] != 0; }
  have_wa_ = have_wa_ || nz; wa_since_upgrade_ = true; invalidate_();
}
void SubbandMVDRGSC::configure_weights_(btkb_pipeline* p) {
  SubbandMVDR::configure_weights_(p);
  if (bm_from_mvdr_) ck(btkb_set_blocking_source(p, 1));
  const unsigned C = chanN(), K = fftLen_ / 2 + 1;
  for (const auto& w : upgrades_) {   // replay: the active weights in force at each upgrade, then the upgrade
    if (w.size() == (size_t)K * (C - 1)) ck(btkb_set_active_weights(p, 1, reinterpret_cast<const float*>(w.data())));
    ck(btkb_upgrade_blocking_matrix(p));
  }
  if (have_wa_ && (upgrades_.empty() || wa_since_upgrade_)) ck(btkb_set_active_weights(p, 1, reinterpret_cast<const float*>(wa_.data())));
}
void SubbandMVDRGSC::upgrade_blocking_matrix() {
  if (halfBandShift_) throw j_error("halfBandShift is not implemented\n");
  upgrades_.push_back(have_wa_ ? wa_ : std::vector<std::complex<float>>());
  wa_since_upgrade_ = false; invalidate_();
}
const cplx* SubbandMVDRGSC::blocking_matrix_output(int outChanX) {
  if (chunk_blocks_ > 0) throw j_error("blocking_matrix_output: not offered for a chunked realisation (set_chunk_blocks(0))\n");
  if (!realized_) run_graph(PostFilterConfig(), SynthesisConfig());
  const unsigned K = fftLen_ / 2 + 1;
  if (Z_chan_ != outChanX || Z_of_ != (const void*)Y_.data() || Z_T_ != T_) {
    Z_.resize((size_t)btkb_num_frames(pipe_) * K);
    ck(btkb_blocking_matrix_output(pipe_, outChanX, reinterpret_cast<float*>(Z_.data())));
    Z_chan_ = outChanX; Z_of_ = (const void*)Y_.data(); Z_T_ = T_;
  }
  bmout_.assign(fftLen_, cplx(0, 0));
  const int t = frame_no_ < 0 ? 0 : frame_no_;
  if (t >= T_) throw jiterator_error("end of samples!");
  for (unsigned k = 0; k < K; k++) { const std::complex<float> z = Z_[(size_t)t * K + k]; bmout_[k] = cplx(z.real(), z.imag()); }
  if (frame_no_ >= 0) for (unsigned k = 1; k < fftLen_ / 2; k++) bmout_[fftLen_ - k] = vector_[fftLen_ - k];   // what next() left in the shared vector
  return bmout_.data();
}
const cplx* SubbandOrthogonalizer::next(int frame_no) {   // beamformer.cc:2786-2806
  if (frame_no == frame_no_) return vector_.data();
  const cplx* v = (outChanX_ <= 0) ? beamformer_->next(frame_no) : beamformer_->blocking_matrix_output(outChanX_ - 1);
  std::copy(v, v + size(), vector_.begin());
  increment_();
  return vector_.data();
}
void SpectralMatrixArray::update() {   // beamformer.cc:122-143
  SnapShotArray::update();
  const unsigned C = nChan();
  const cplx alpha = cplx(1.0, 0) - mu_;
  for (unsigned f = 0; f < fftLen(); f++) {
    cplx* R = &matrices_[(size_t)f * C * C];
    const cplx* x = snapshot(f);
    for (unsigned i = 0; i < C; i++)
      for (unsigned j = 0; j < C; j++) R[i * C + j] = R[i * C + j] * mu_ + alpha * (x[i] * x[j]);
  }
}

// ================================================================================================ post-filter
ZelinskiPostFilter::ZelinskiPostFilter(const VectorComplexFeatureStreamPtr& output, unsigned fftLen, double alpha, int type, int min_frames, const std::string& nm)
    : VectorComplexFeatureStream(fftLen, nm), fftLen_(fftLen), samp_(output), alpha_(alpha), type_(type), min_frames_(min_frames) {
  if (output->size() != fftLen) throw jdimension_error("Input block length (%d) != fftLen (%d)\n", output->size(), fftLen);  // postfilter.cc:366-368
  bf_ = std::dynamic_pointer_cast<SubbandDS>(output);
}
void ZelinskiPostFilter::reset() { samp_->reset(); if (bf_) bf_->reset(); VectorComplexFeatureStream::reset(); is_end_ = false; foreign_ready_ = false; }
ZelinskiPostFilter::~ZelinskiPostFilter() { if (fpipe_) btkb_destroy(fpipe_); }
void ZelinskiPostFilter::set_array_manifold_vector(unsigned fbinX, const std::vector<cplx>& v, bool half_band_shift, unsigned NC) {   // postfilter.cc:393-417
  if (fbinX >= fftLen_) throw jdimension_error("fbinX %d must be less than %d\n", (int)fbinX, (int)fftLen_);
  if (half_band_shift) throw j_error("halfBandShift is not implemented\n");
  if (NC != 1) throw j_error("set_array_manifold_vector: one constraint only\n");
  const unsigned K = fftLen_ / 2 + 1;
  if (manifold_C_ != v.size()) { manifold_C_ = (unsigned)v.size(); manifold_.assign((size_t)K * manifold_C_, std::complex<float>(0, 0)); }
  if (fbinX < K) for (unsigned c = 0; c < manifold_C_; c++) manifold_[(size_t)fbinX * manifold_C_ + c] = std::complex<float>((float)v[c].real(), (float)v[c].imag());
  foreign_ready_ = false;
}
void ZelinskiPostFilter::realize_foreign_() {
  const unsigned K = fftLen_ / 2 + 1, C = manifold_C_;
  if (C != snap_foreign_->nChan() || fftLen_ != snap_foreign_->fftLen()) throw jdimension_error("snapshot array (%d channels) and manifold vectors (%d) do not match\n", (int)snap_foreign_->nChan(), (int)C);
  if (config().kind != BTKB_PF_ZELINSKI) throw j_error("set_snapshot_array: offered for the Zelinski filter (McCowan / Lefkimmiatis take set_beamformer)\n");
  // drain the foreign output; whoever produces it updates the shared snapshot array as a side effect of its own next()
  std::vector<std::complex<float>> X;
  fout_.clear(); fT_ = 0;
  for (int t = 0;; t++) {
    const cplx* f;
    try { f = samp_->next(t); } catch (jiterator_error&) { break; }
    fout_.insert(fout_.end(), f, f + fftLen_);
    X.resize((size_t)(t + 1) * C * K);
    for (unsigned c = 0; c < C; c++)
      for (unsigned k = 0; k < K; k++) { const cplx v = snap_foreign_->snapshot(k)[c]; X[((size_t)t * C + c) * K + k] = std::complex<float>((float)v.real(), (float)v.imag()); }
    fT_ = t + 1;
  }
  fgain_.assign((size_t)fT_ * K, 1.f);
  if (fT_ > 0) {
    if (fpipe_) { btkb_destroy(fpipe_); fpipe_ = nullptr; }
    btkb_config c; btkb_default_config(&c);
    c.channels = (int)C; c.fft_len = (int)fftLen_; c.m = 4; c.r = 1; c.delay_compensation_type = 2; c.max_utterances = 1;
    c.max_samples = (fT_ + 8) * (int)(fftLen_ >> 1);
    c.beamformer = BTKB_BF_DS; c.postfilter = BTKB_PF_ZELINSKI; c.pf_alpha = (float)alpha_; c.pf_type = type_; c.pf_min_frames = min_frames_;
    ck(btkb_create(&c, &fpipe_));
    ck(btkb_set_weights(fpipe_, 1, reinterpret_cast<const float*>(manifold_.data())));
    ck(btkb_set_snapshots(fpipe_, 1, fT_, reinterpret_cast<const float*>(X.data())));
    ck(btkb_run_beamformer(fpipe_, 0));
    ck(btkb_get_postfilter_weights(fpipe_, fgain_.data()));
  }
  foreign_ready_ = true;
}
const cplx* ZelinskiPostFilter::next(int frame_no) {  // postfilter.cc:424-491
  if (frame_no == frame_no_) return vector_.data();
  if (!bf_ && snap_foreign_ && manifold_C_ > 0) {
    if (!foreign_ready_) realize_foreign_();
    if (frame_no_ + 1 >= fT_) { is_end_ = true; throw jiterator_error("end of samples!"); }
    increment_();
    const unsigned K = fftLen_ / 2 + 1;
    const cplx* y = &fout_[(size_t)frame_no_ * fftLen_];
    for (unsigned k = 0; k < fftLen_; k++) vector_[k] = y[k];
    // the reference tests frame_no_ < min_frames_ BEFORE it increments (postfilter.cc:470-475): frames 0 .. min_frames only update the
    // statistics (NO_USE_POST_FILTER returns before the filtering loop, :196-198) and the output passes as it came
    if (frame_no_ > min_frames_) {
      for (unsigned k = 0; k < K; k++) vector_[k] = y[k] * (double)fgain_[(size_t)frame_no_ * K + k];   // ZelinskiFilter: vector_ <- wp1 * output (postfilter.cc:57-219)
      for (unsigned k = 1; k < fftLen_ / 2; k++) vector_[fftLen_ - k] = std::conj(vector_[k]);
    }
    return vector_.data();
  }
  if (!bf_) throw j_error("set beamformer's weights \n");  // postfilter.cc:443-445
  const PostFilterConfig pf = config();
  if (!bf_->realized_with(pf, SynthesisConfig())) bf_->run_graph(pf, SynthesisConfig());
  if (frame_no_ + 1 >= bf_->frames()) { is_end_ = true; throw jiterator_error("end of samples!"); }
  bf_->ensure_frames(frame_no_ + 1);
  increment_();
  const unsigned K = fftLen_ / 2 + 1;
  const std::complex<float>* y = &bf_->Y()[(size_t)frame_no_ * K];
  for (unsigned k = 0; k < K; k++) vector_[k] = cplx(y[k].real(), y[k].imag());
  if (frame_no_ < onesided_frames_()) { for (unsigned k = K; k < fftLen_; k++) vector_[k] = cplx(0, 0); }
  else for (unsigned k = 1; k < fftLen_ / 2; k++) vector_[fftLen_ - k] = std::conj(vector_[k]);
  return vector_.data();
}

// ---- McCowan / Lefkimmiatis
McCowanPostFilter::McCowanPostFilter(const VectorComplexFeatureStreamPtr& output, unsigned fftLen, double alpha, int type, int min_frames, float threshold,
                                     const std::string& nm)
    : ZelinskiPostFilter(output, fftLen, alpha, type, min_frames, nm), kind_(BTKB_PF_MCCOWAN), threshold_(threshold) {}
McCowanPostFilter::~McCowanPostFilter() { if (store_) btkb_destroy(store_); }
void McCowanPostFilter::ensure_store_(unsigned C) {
  if (store_ && chanN_ == C) return;
  if (store_) { btkb_destroy(store_); store_ = nullptr; }
  btkb_config c; btkb_default_config(&c);
  c.channels = (int)C; c.fft_len = (int)fftLen_; c.postfilter = BTKB_PF_MCCOWAN; c.max_utterances = 1; c.max_samples = (int)fftLen_;
  ck(btkb_create(&c, &store_));
  chanN_ = C; haveR_ = false;
}
void McCowanPostFilter::require_R_() const { if (!store_ || !haveR_) throw j_error("Construct/set first a noise coherence matrix\n"); }  // postfilter.cc:631-633
std::vector<cplx> McCowanPostFilter::get_all_() const {
  std::vector<cplx> R((size_t)(fftLen_ / 2 + 1) * chanN_ * chanN_);
  ck(btkb_pf_get_noise_coherence(store_, reinterpret_cast<double*>(R.data())));
  return R;
}
void McCowanPostFilter::set_all_(const std::vector<cplx>& R) {
  ck(btkb_pf_set_noise_coherence(store_, reinterpret_cast<const double*>(R.data())));
  haveR_ = true; version_++;
}
std::vector<cplx> McCowanPostFilter::noise_spatial_spectral_matrix(unsigned fbinX) {
  require_R_();
  if (fbinX > fftLen_ / 2) throw jindex_error("frequency bin %d out of range", fbinX);
  const std::vector<cplx> R = get_all_();
  const size_t n = (size_t)chanN_ * chanN_;
  return std::vector<cplx>(R.begin() + fbinX * n, R.begin() + (fbinX + 1) * n);
}
bool McCowanPostFilter::set_noise_spatial_spectral_matrix(unsigned fbinX, const std::vector<cplx>& Rnn, unsigned rows, unsigned cols) {
  if (rows != cols) { fprintf(stderr, "The noise coherence matrix should be the square matrix\n"); return false; }  // postfilter.cc:543-546
  if (fbinX > fftLen_ / 2) throw jindex_error("frequency bin %d out of range", fbinX);
  const bool fresh = !(store_ && chanN_ == rows && haveR_);
  ensure_store_(rows);
  const size_t n = (size_t)rows * rows;
  std::vector<cplx> R = fresh ? std::vector<cplx>((size_t)(fftLen_ / 2 + 1) * n, cplx(0, 0)) : get_all_();
  std::copy(Rnn.begin(), Rnn.begin() + n, R.begin() + fbinX * n);
  set_all_(R);
  return true;
}
bool McCowanPostFilter::set_diffuse_noise_model(const std::vector<double>& mpos, unsigned rows, unsigned cols, double sampleRate, double sspeed) {
  if (cols < 3) { fprintf(stderr, "The microphone positions should be described in the three dimensions\n"); return false; }  // postfilter.cc:567-570
  ensure_store_(rows);
  std::vector<double> mp((size_t)rows * 3);
  for (unsigned c = 0; c < rows; c++) for (unsigned j = 0; j < 3; j++) mp[c * 3 + j] = mpos[(size_t)c * cols + j];
  ck(btkb_pf_set_diffuse_noise_model(store_, mp.data(), sampleRate, sspeed));
  haveR_ = true; version_++;
  return true;
}
void McCowanPostFilter::set_all_diagonal_loading(float w) { require_R_(); ck(btkb_pf_set_diagonal_loading(store_, w)); version_++; }
void McCowanPostFilter::divide_all_nondiagonal_elements(float mu) { require_R_(); ck(btkb_pf_divide_nondiagonal(store_, mu)); version_++; }
void McCowanPostFilter::set_diagonal_looading(unsigned fbinX, float w) {  // postfilter.cc:644-655
  require_R_();
  std::vector<cplx> R = get_all_();
  const size_t n = (size_t)chanN_ * chanN_;
  for (unsigned c = 0; c < chanN_; c++) R[fbinX * n + (size_t)c * chanN_ + c] += (double)w;
  set_all_(R);
}
void McCowanPostFilter::divide_nondiagonal_elements(unsigned fbinX, float mu) {  // postfilter.cc:669-680
  require_R_();
  std::vector<cplx> R = get_all_();
  const size_t n = (size_t)chanN_ * chanN_;
  for (unsigned i = 0; i < chanN_; i++) for (unsigned j = 0; j < chanN_; j++) if (i != j) R[fbinX * n + (size_t)i * chanN_ + j] /= (1.0 + mu);
  set_all_(R);
}
PostFilterConfig McCowanPostFilter::config() const {
  PostFilterConfig c = ZelinskiPostFilter::config();
  c.kind = kind_; c.threshold = threshold_; c.min_sv = min_sv_; c.fbin1 = fbin1_;
  c.coherence = haveR_ ? this : nullptr; c.coherence_version = version_;
  return c;
}
void McCowanPostFilter::push_coherence(btkb_pipeline* dst) const {
  const std::vector<cplx> R = get_all_();
  ck(btkb_pf_set_noise_coherence(dst, reinterpret_cast<const double*>(R.data())));
}
LefkimmiatisPostFilter::LefkimmiatisPostFilter(const VectorComplexFeatureStreamPtr& output, unsigned fftLen, double minSV, unsigned fbinX1, double alpha, int type,
                                               int min_frames, float threshold, const std::string& nm)
    : McCowanPostFilter(output, fftLen, alpha, type, min_frames, threshold, nm) { kind_ = BTKB_PF_LEFKIMMIATIS; min_sv_ = minSV; fbin1_ = fbinX1; }
void LefkimmiatisPostFilter::calc_inverse_noise_spatial_spectral_matrix() { require_R_(); }

std::vector<cplx> ZelinskiPostFilter::postfilter_weights() {
  std::vector<cplx> w(fftLen_, cplx(0, 0));
  if (!bf_ || frame_no_ < 0 || bf_->pf_weights().empty()) return w;
  const unsigned K = fftLen_ / 2 + 1;
  for (unsigned k = 0; k < K; k++) w[k] = cplx(bf_->pf_weights()[(size_t)frame_no_ * K + k], 0);
  for (unsigned k = 1; k < fftLen_ / 2; k++) w[fftLen_ - k] = w[k];
  return w;
}

// ================================================================================================ synthesis bank
OverSampledDFTSynthesisBank::OverSampledDFTSynthesisBank(const VectorComplexFeatureStreamPtr& samp, const std::vector<double>& prototype, unsigned M, unsigned m,
                                                         unsigned r, unsigned dct, int gain, const std::string& nm)
    : VectorFloatFeatureStream(M >> r, nm), samp_(samp), prototype_(prototype), M_(M), m_(m), r_(r), D_(M >> r), dct_(dct), gain_(gain),
      pd_(synthesis_delay((int)m, (int)r, (int)dct)), pipe_(nullptr), nb_(0), realized_(false) {
  if (prototype.size() != (size_t)M * m) throw jconsistency_error("Prototype sizes do not match (%d vs. %d).", (int)prototype.size(), (int)(M * m));
}
OverSampledDFTSynthesisBank::~OverSampledDFTSynthesisBank() { if (pipe_) btkb_destroy(pipe_); }
void OverSampledDFTSynthesisBank::reset() { samp_->reset(); VectorFloatFeatureStream::reset(); realized_ = false; live_bf_ = nullptr; pushed_.clear(); npushed_ = 0; }
void OverSampledDFTSynthesisBank::input_source_vector(const std::vector<cplx>& block) {   // modulated.h:330 -> update_buf_ (modulated.cc:551-567)
  if (block.size() != M_) throw jdimension_error("Input block length (%d) != fftLen (%d)\n", (int)block.size(), (int)M_);
  if (realized_) throw j_error("input_source_vector: frames can be pushed until the first next() (the mirror synthesises the utterance in one pass)\n");
  // pushed frames enter the buffer without a polyphase sum (update_buf_ only).  With R = 1 an output depends on the buffer alone and the
  // concatenated sequence reproduces the reference exactly; with R > 1 the first R - 1 outputs would also need the polyphase sums the
  // reference did NOT form for the pushed frames, which a one-pass synthesis of the concatenation cannot leave out.
  if (r_ != 0) throw j_error("input_source_vector: offered for critically sampled banks (r = 0) only\n");
  const unsigned K = M_ / 2 + 1;
  for (unsigned k = 0; k < K; k++) pushed_.push_back(std::complex<float>((float)block[k].real(), (float)block[k].imag()));
  npushed_++;
}

void OverSampledDFTSynthesisBank::realize_() {
  SynthesisConfig syn; syn.enabled = true; syn.prototype = prototype_; syn.M = M_; syn.m = m_; syn.r = r_; syn.dct = dct_; syn.gain = gain_;
  SubbandBeamformer* bf = nullptr; PostFilterConfig pf;
  if (auto* z = dynamic_cast<ZelinskiPostFilter*>(samp_.get())) { bf = z->beamformer().get(); pf = z->config(); }
  else bf = dynamic_cast<SubbandBeamformer*>(samp_.get());
  if (bf && npushed_ == 0) {  // fast path: the whole graph runs on the GPU in one submission
    if (!bf->realized_with(pf, syn)) bf->run_graph(pf, syn);
    nb_ = bf->blocks(); live_bf_ = bf;
    if (bf->chunk_blocks() == 0) { out_ = bf->time_out(); live_bf_ = nullptr; }   // chunked realisation: blocks are pulled as next() asks for them
  } else {
    live_bf_ = nullptr;   // arbitrary upstream stream (e.g. a Python object behind PyVectorComplexFeatureStream): drain it, synthesise on the GPU
    const unsigned K = M_ / 2 + 1;
    std::vector<std::complex<float>> Y = pushed_;   // frames handed in through input_source_vector sit in the buffer ahead of the source's
    int T = npushed_;
    for (int ts = 0;; ts++) {
      const cplx* f;
      try { f = samp_->next(ts); } catch (jiterator_error&) { break; }
      Y.resize((size_t)(T + 1) * K);
      for (unsigned k = 0; k < K; k++) Y[(size_t)T * K + k] = std::complex<float>((float)f[k].real(), (float)f[k].imag());
      T++;
    }
    nb_ = std::max(T - pd_, 0);
    out_.assign((size_t)nb_ * D_, 0.f);
    if (nb_ > 0) {
      if (pipe_) { btkb_destroy(pipe_); pipe_ = nullptr; }
      btkb_config c; btkb_default_config(&c);
      c.channels = 1; c.fft_len = (int)M_; c.m = (int)m_; c.r = (int)r_; c.delay_compensation_type = (int)dct_; c.max_utterances = 1;
      c.max_samples = (T + 8) * (int)D_; c.synthesis_gain = gain_;
      ck(btkb_create(&c, &pipe_));
      ck(btkb_set_prototypes(pipe_, prototype_.data(), prototype_.data(), (int)prototype_.size()));
      ck(btkb_set_subband(pipe_, 1, T, reinterpret_cast<const float*>(Y.data())));
      ck(btkb_run_synthesis(pipe_));
      nb_ = btkb_num_blocks(pipe_);
      out_.resize((size_t)nb_ * D_);
      ck(btkb_fetch_time(pipe_, out_.data()));
    }
  }
  realized_ = true;
}

const float* OverSampledDFTSynthesisBank::next(int frame_no) {  // modulated.cc:569-612
  if (frame_no == frame_no_ + pd_) return vector_.data();
  if (!realized_) realize_();
  if (frame_no_ + 1 >= nb_) { is_end_ = true; throw jiterator_error("end of samples!"); }
  if (frame_no >= 0 && frame_no - 1 != frame_no_) printf("The output might not be continuous %s: %d != %d\n", name().c_str(), frame_no - 1, frame_no_);
  increment_();
  if (live_bf_) { live_bf_->ensure_blocks(frame_no_); std::memcpy(vector_.data(), &live_bf_->time_out()[(size_t)frame_no_ * D_], sizeof(float) * D_); }
  else std::memcpy(vector_.data(), &out_[(size_t)frame_no_ * D_], sizeof(float) * D_);
  return vector_.data();
}

std::vector<double> calc_all_delays(double, double, double, const std::vector<double>& mpos) {  // beamformer.cc:1170-1189 (sic: ignores x,y,z)
  const size_t C = mpos.size() / 3;
  std::vector<double> d(C);
  for (size_t c = 0; c < C; c++) d[c] = std::sqrt(mpos[3 * c] * mpos[3 * c] + mpos[3 * c + 1] * mpos[3 * c + 1] + mpos[3 * c + 2] * mpos[3 * c + 2]) / 343740.0;
  const double mid = d[C / 2];
  for (auto& v : d) v -= mid;
  return d;
}

}  // namespace btk20
