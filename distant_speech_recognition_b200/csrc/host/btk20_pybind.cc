// btk20_pybind.cc — Python surface of the C++ host mirror (pybind11 stands in for SWIG, which is not installed).
// Mirrors the SWIG interface files of the reference: stream/stream.i, feature/feature.i:205-250, modulated/modulated.i:91-192,
// beamformer/beamformer.i:46-540, postfilter/postfilter.i:46-90, include/jexception.i:20-86 (exception map).
#include <pybind11/complex.h>
#include <pybind11/numpy.h>
#include <cstring>
#include <pybind11/pybind11.h>
#include <pybind11/stl.h>

#include "btk20_host.h"

namespace py = pybind11;
using namespace btk20;

namespace {

// stream/pyStream.h:25-133 — a Python iterable exposing __iter__/next/size/reset as a C++ stream (one adapter per element type:
// PyVectorShortFeatureStream, PyVectorFloatFeatureStream, PyVectorFeatureStream, PyVectorComplexFeatureStream)
template <class T>
class PyFeatureStream : public FeatureStream<T> {
 public:
  PyFeatureStream(py::object obj, const std::string& nm)
      : FeatureStream<T>(py::cast<unsigned>(obj.attr("size")()), nm), obj_(obj), iter_(py::none()) {}
  const T* next(int frame_no = -5) override {
    if (frame_no == this->frame_no_) return this->vector_.data();
    py::gil_scoped_acquire gil;
    if (iter_.is_none()) iter_ = obj_.attr("__iter__")();
    py::object item;
    try {
      item = py::hasattr(iter_, "__next__") ? iter_.attr("__next__")() : iter_.attr("next")();
    } catch (py::error_already_set& e) {
      if (e.matches(PyExc_StopIteration)) { this->is_end_ = true; throw jiterator_error("end of samples!"); }
      throw;
    }
    auto arr = py::array_t<T, py::array::c_style | py::array::forcecast>::ensure(item);
    if (!arr || (unsigned)arr.size() != this->size()) throw jdimension_error("PyFeatureStream: expected a vector of length %d", this->size());
    std::memcpy(this->vector_.data(), arr.data(), sizeof(T) * this->size());
    this->increment_();
    return this->vector_.data();
  }
  void reset() override {
    py::gil_scoped_acquire gil;
    if (py::hasattr(obj_, "reset")) obj_.attr("reset")();
    iter_ = py::none();
    FeatureStream<T>::reset();
  }
 private:
  py::object obj_, iter_;
};
typedef PyFeatureStream<cplx> PyVectorComplexFeatureStream;

template <class T, class S>
py::array_t<T> view(S& self, const T* p, size_t n) {  // aliases internal storage, like PyArray_FromDimsAndData (include/vector.i:302)
  return py::array_t<T>({n}, {sizeof(T)}, p, py::cast(&self, py::return_value_policy::reference));
}

template <class Cls, class PyCls>
void add_complex_stream_api(PyCls& c) {
  c.def("next", [](Cls& s, int frame_no) { const cplx* p = s.next(frame_no); return view<cplx>(s, p, s.size()); }, py::arg("frame_no") = -5)
      .def("__next__", [](Cls& s) { const cplx* p = s.next(-5); return view<cplx>(s, p, s.size()); })
      .def("current", [](Cls& s) { const cplx* p = s.current(); return view<cplx>(s, p, s.size()); })
      .def("__iter__", [](py::object self) { self.attr("reset")(); return self; })
      .def("reset", &Cls::reset)
      .def("size", &Cls::size)
      .def("is_end", &Cls::is_end)
      .def("frame_no", &Cls::frame_no)
      .def("name", &Cls::name);
}
template <class Cls, class PyCls>
void add_float_stream_api(PyCls& c) {
  c.def("next", [](Cls& s, int frame_no) { const float* p = s.next(frame_no); return view<float>(s, p, s.size()); }, py::arg("frame_no") = -5)
      .def("__next__", [](Cls& s) { const float* p = s.next(-5); return view<float>(s, p, s.size()); })
      .def("current", [](Cls& s) { const float* p = s.current(); return view<float>(s, p, s.size()); })
      .def("__iter__", [](py::object self) { self.attr("reset")(); return self; })
      .def("reset", &Cls::reset)
      .def("size", &Cls::size)
      .def("is_end", &Cls::is_end)
      .def("frame_no", &Cls::frame_no)
      .def("name", &Cls::name);
}
template <class T, class PyCls>
void add_stream_api(PyCls& c) {
  typedef FeatureStream<T> Cls;
  c.def("next", [](Cls& s, int frame_no) { const T* p = s.next(frame_no); return view<T>(s, p, s.size()); }, py::arg("frame_no") = -5)
      .def("__next__", [](Cls& s) { const T* p = s.next(-5); return view<T>(s, p, s.size()); })
      .def("__iter__", [](py::object self) { self.attr("reset")(); return self; })
      .def("current", [](Cls& s) { const T* p = s.current(); return view<T>(s, p, s.size()); })
      .def("reset", &Cls::reset)
      .def("size", &Cls::size)
      .def("is_end", &Cls::is_end)
      .def("frame_no", &Cls::frame_no)
      .def("name", &Cls::name);
}
template <class T>
void add_stream_classes(py::module_& m, const char* base_name, const char* py_name) {
  py::class_<FeatureStream<T>, std::shared_ptr<FeatureStream<T>>> base(m, base_name);
  add_stream_api<T>(base);
  py::class_<PyFeatureStream<T>, FeatureStream<T>, std::shared_ptr<PyFeatureStream<T>>>(m, py_name)
      .def(py::init<py::object, const std::string&>(), py::arg("obj"), py::arg("nm") = py_name);
}
std::vector<double> vec_d(py::array_t<double, py::array::c_style | py::array::forcecast> a) { return std::vector<double>(a.data(), a.data() + a.size()); }

}  // namespace

PYBIND11_MODULE(_btk20host, m) {
  m.doc() = "C++ host mirror of btk2.0's hot-path stream classes over libbtkb.so (sm_100a CUDA)";

  // include/jexception.i:20-86
  static py::exception<j_error> ex_base(m, "j_error", PyExc_Exception);
  py::register_exception_translator([](std::exception_ptr p) {
    try { if (p) std::rethrow_exception(p); }
    catch (const jiterator_error& e) { PyErr_SetString(PyExc_StopIteration, "stop iteration"); }
    catch (const jallocation_error& e) { PyErr_SetString(PyExc_MemoryError, e.what()); }
    catch (const jarithmetic_error& e) { PyErr_SetString(PyExc_ArithmeticError, e.what()); }
    catch (const jnumeric_error& e) { PyErr_SetString(PyExc_FloatingPointError, e.what()); }
    catch (const jindex_error& e) { PyErr_SetString(PyExc_IndexError, e.what()); }
    catch (const jio_error& e) { PyErr_SetString(PyExc_IOError, e.what()); }
    catch (const jkey_error& e) { PyErr_SetString(PyExc_KeyError, e.what()); }
    catch (const jparameter_error& e) { PyErr_SetString(PyExc_ValueError, e.what()); }
    catch (const jparse_error& e) { PyErr_SetString(PyExc_SyntaxError, e.what()); }
    catch (const jtype_error& e) { PyErr_SetString(PyExc_TypeError, e.what()); }
    catch (const j_error& e) { PyErr_SetString(PyExc_Exception, e.what()); }
  });

  py::class_<VectorFloatFeatureStream, VectorFloatFeatureStreamPtr> vf(m, "VectorFloatFeatureStreamPtr");
  add_float_stream_api<VectorFloatFeatureStream>(vf);
  py::class_<VectorComplexFeatureStream, VectorComplexFeatureStreamPtr> vc(m, "VectorComplexFeatureStreamPtr");
  add_complex_stream_api<VectorComplexFeatureStream>(vc);

  py::class_<PyVectorComplexFeatureStream, VectorComplexFeatureStream, std::shared_ptr<PyVectorComplexFeatureStream>>(m, "PyVectorComplexFeatureStreamPtr")
      .def(py::init<py::object, const std::string&>(), py::arg("obj"), py::arg("nm") = "PyVectorComplexFeatureStream");
  py::class_<PyFeatureStream<float>, VectorFloatFeatureStream, std::shared_ptr<PyFeatureStream<float>>>(m, "PyVectorFloatFeatureStreamPtr")
      .def(py::init<py::object, const std::string&>(), py::arg("obj"), py::arg("nm") = "PyVectorFloatFeatureStream");
  // the remaining element types of stream/stream.i:24-237 (no hot-path class produces or consumes them; plumbing only)
  add_stream_classes<double>(m, "VectorFeatureStreamPtr", "PyVectorFeatureStreamPtr");
  add_stream_classes<short>(m, "VectorShortFeatureStreamPtr", "PyVectorShortFeatureStreamPtr");
  add_stream_classes<char>(m, "VectorCharFeatureStreamPtr", "PyVectorCharFeatureStreamPtr");

  py::class_<SampleFeature, VectorFloatFeatureStream, SampleFeaturePtr>(m, "SampleFeaturePtr")
      .def(py::init<const std::string&, unsigned, unsigned, bool, const std::string&>(), py::arg("fn") = "", py::arg("block_len") = 320,
           py::arg("shift_len") = 160, py::arg("pad_zeros") = false, py::arg("nm") = "Sample")
      .def("read", &SampleFeature::read, py::arg("fn"), py::arg("format") = 0, py::arg("samplerate") = 16000, py::arg("chX") = 1, py::arg("chN") = 1,
           py::arg("cfrom") = 0, py::arg("to") = -1, py::arg("outsamplerate") = -1, py::arg("norm") = 0.0f)
      .def("setSamples", [](SampleFeature& s, py::array_t<double, py::array::c_style | py::array::forcecast> a, unsigned rate) { s.set_samples(a.data(), (unsigned)a.size(), rate); },
           py::arg("samples"), py::arg("samplerate"))
      .def("set_samples", [](SampleFeature& s, py::array_t<double, py::array::c_style | py::array::forcecast> a, unsigned rate) { s.set_samples(a.data(), (unsigned)a.size(), rate); },
           py::arg("samples"), py::arg("samplerate"))
      .def("data", [](SampleFeature& s) { return view<float>(s, s.samples().data(), s.samples().size()); })
      .def("samplesN", &SampleFeature::samplesN)
      .def("getSampleRate", &SampleFeature::samplerate);

  py::class_<OverSampledDFTAnalysisBank, VectorComplexFeatureStream, OverSampledDFTAnalysisBankPtr>(m, "OverSampledDFTAnalysisBankPtr")
      .def(py::init([](VectorFloatFeatureStreamPtr samp, py::array_t<double, py::array::c_style | py::array::forcecast> prototype, unsigned M, unsigned mm, unsigned r,
                       unsigned dct, const std::string& nm) { return std::make_shared<OverSampledDFTAnalysisBank>(samp, vec_d(prototype), M, mm, r, dct, nm); }),
           py::arg("samp"), py::arg("prototype"), py::arg("M") = 256, py::arg("m") = 3, py::arg("r") = 0, py::arg("delay_compensation_type") = 0,
           py::arg("nm") = "OverSampledDFTAnalysisBankFloat")
      .def("fftlen", &OverSampledDFTAnalysisBank::fftlen)
      .def("shiftlen", &OverSampledDFTAnalysisBank::shiftlen)
      .def("fftLen", &OverSampledDFTAnalysisBank::fftlen)
      .def("polyphase", &OverSampledDFTAnalysisBank::polyphase, py::arg("m"), py::arg("n"));

  py::class_<OverSampledDFTSynthesisBank, VectorFloatFeatureStream, OverSampledDFTSynthesisBankPtr>(m, "OverSampledDFTSynthesisBankPtr")
      .def(py::init([](VectorComplexFeatureStreamPtr samp, py::array_t<double, py::array::c_style | py::array::forcecast> prototype, unsigned M, unsigned mm, unsigned r,
                       unsigned dct, int gain, const std::string& nm) { return std::make_shared<OverSampledDFTSynthesisBank>(samp, vec_d(prototype), M, mm, r, dct, gain, nm); }),
           py::arg("samp"), py::arg("prototype"), py::arg("M"), py::arg("m"), py::arg("r") = 0, py::arg("delay_compensation_type") = 0, py::arg("gain_factor") = 1,
           py::arg("nm") = "OverSampledDFTSynthesisBank")
      .def("polyphase", &OverSampledDFTSynthesisBank::polyphase, py::arg("m"), py::arg("n"))
      .def("input_source_vector", [](OverSampledDFTSynthesisBank& s, py::array_t<std::complex<double>, py::array::c_style | py::array::forcecast> block) {
             s.input_source_vector(std::vector<cplx>(block.data(), block.data() + block.size())); }, py::arg("block"));

  // dereverberation/dereverberation.i:46-185
  py::class_<SingleChannelWPEDereverberationFeature, VectorComplexFeatureStream, SingleChannelWPEDereverberationFeaturePtr>(m, "SingleChannelWPEDereverberationFeaturePtr")
      .def(py::init([](VectorComplexFeatureStreamPtr samples, unsigned lower_num, unsigned upper_num, unsigned iterations_num, double load_db, double band_width,
                       double samplerate, const std::string& nm) {
             return std::make_shared<SingleChannelWPEDereverberationFeature>(samples, lower_num, upper_num, iterations_num, load_db, band_width, samplerate, nm); }),
           py::arg("samples"), py::arg("lower_num") = 0, py::arg("upper_num") = 64, py::arg("iterations_num") = 2, py::arg("load_db") = -20.0,
           py::arg("band_width") = 0.0, py::arg("samplerate") = 16000.0, py::arg("nm") = "SingleChannelWPEDereverberationFeature")
      .def("estimate_filter", &SingleChannelWPEDereverberationFeature::estimate_filter, py::arg("start_frame_no") = 0, py::arg("frame_num") = -1)
      .def("print_objective_func", &SingleChannelWPEDereverberationFeature::print_objective_func, py::arg("subband_no"))
      .def("reset_filter", &SingleChannelWPEDereverberationFeature::reset_filter)
      .def("next_speaker", &SingleChannelWPEDereverberationFeature::next_speaker);
  py::class_<MultiChannelWPEDereverberation, MultiChannelWPEDereverberationPtr>(m, "MultiChannelWPEDereverberationPtr")
      .def(py::init([](unsigned subbands_num, unsigned channels_num, unsigned lower_num, unsigned upper_num, unsigned iterations_num, double load_db,
                       double band_width, double diagonal_bias, double samplerate) {
             return std::make_shared<MultiChannelWPEDereverberation>(subbands_num, channels_num, lower_num, upper_num, iterations_num, load_db, band_width,
                                                                     diagonal_bias, samplerate); }),
           py::arg("subbands_num"), py::arg("channels_num"), py::arg("lower_num") = 0, py::arg("upper_num") = 32, py::arg("iterations_num") = 2,
           py::arg("load_db") = -20.0, py::arg("band_width") = 0.0, py::arg("diagonal_bias") = 0.001, py::arg("samplerate") = 16000.0)
      .def("size", &MultiChannelWPEDereverberation::size)
      .def("set_input", &MultiChannelWPEDereverberation::set_input, py::arg("samples"))
      .def("estimate_filter", &MultiChannelWPEDereverberation::estimate_filter, py::arg("start_frame_no") = 0, py::arg("frame_num") = -1)
      .def("reset_filter", &MultiChannelWPEDereverberation::reset_filter)
      .def("next_speaker", &MultiChannelWPEDereverberation::next_speaker)
      .def("print_objective_func", &MultiChannelWPEDereverberation::print_objective_func, py::arg("subband_no"))
      .def("reset", &MultiChannelWPEDereverberation::reset);
  py::class_<MultiChannelWPEDereverberationFeature, VectorComplexFeatureStream, MultiChannelWPEDereverberationFeaturePtr>(m, "MultiChannelWPEDereverberationFeaturePtr")
      .def(py::init([](MultiChannelWPEDereverberationPtr source, unsigned channel_no, unsigned primary_channel_no, const std::string& nm) {
             return std::make_shared<MultiChannelWPEDereverberationFeature>(source, channel_no, primary_channel_no, nm); }),
           py::arg("source"), py::arg("channel_no"), py::arg("primary_channel_no") = 0, py::arg("nm") = "MultiChannelWPEDereverberationFeature")
      // extension: the frame shift of the analysis bank behind the channel, so that the feature streams can feed the btk20.pybeamformer
      // classes (MultiChannelSource reads spec_sources[0].shiftlen(), lib/pybeamformer.py:251), not only the C++ beamformers
      .def("shiftlen", [](MultiChannelWPEDereverberationFeature& f) {
        auto* ab = dynamic_cast<OverSampledDFTAnalysisBank*>(f.source()->sources().at(f.channel()).get());
        if (!ab) throw j_error("MultiChannelWPEDereverberationFeature: no analysis bank behind channel %d", (int)f.channel());
        return ab->shiftlen(); });

  py::class_<SnapShotArray, SnapShotArrayPtr>(m, "SnapShotArrayPtr")
      .def(py::init<unsigned, unsigned>(), py::arg("fftlen"), py::arg("chan_num"))
      .def("snapshot", [](SnapShotArray& s, unsigned f) { return view<cplx>(s, s.snapshot(f), s.nChan()); }, py::arg("fbinX"))
      .def("set_samples", [](SnapShotArray& s, py::array_t<std::complex<double>, py::array::c_style | py::array::forcecast> a, unsigned ch) {
        if ((unsigned)a.size() != s.fftLen()) throw jdimension_error("set_samples: expected %d bins", s.fftLen());
        s.set_samples(a.data(), ch); }, py::arg("samp"), py::arg("chanX"))
      .def("update", &SnapShotArray::update).def("zero", &SnapShotArray::zero).def("fftLen", &SnapShotArray::fftLen).def("nChan", &SnapShotArray::nChan);

  // beamformer/beamformer.i:83-114
  py::class_<SpectralMatrixArray, SnapShotArray, SpectralMatrixArrayPtr>(m, "SpectralMatrixArrayPtr")
      .def(py::init<unsigned, unsigned, float>(), py::arg("fftLn"), py::arg("nChn"), py::arg("forgetFact") = 0.95f)
      .def("matrix_f", [](SpectralMatrixArray& s, unsigned idx) {
             const cplx* p = s.matrix_f(idx);
             return py::array_t<cplx>({(size_t)s.nChan(), (size_t)s.nChan()}, {sizeof(cplx) * s.nChan(), sizeof(cplx)}, p, py::cast(&s, py::return_value_policy::reference)); },
           py::arg("idx"))
      .def("update", &SpectralMatrixArray::update).def("zero", &SpectralMatrixArray::zero);

  py::class_<SubbandBeamformer, VectorComplexFeatureStream, SubbandBeamformerPtr>(m, "SubbandBeamformerPtr")
      .def("set_channel", &SubbandBeamformer::set_channel, py::arg("chan"))
      .def("clear_channel", &SubbandBeamformer::clear_channel)
      .def("fftLen", &SubbandBeamformer::fftLen).def("chanN", &SubbandBeamformer::chanN)
      .def("snapshot_array", &SubbandBeamformer::snapshot_array)
      .def("get_weights", &SubbandBeamformer::get_weights, py::arg("fbinX"))
      .def("set_chunk_blocks", &SubbandBeamformer::set_chunk_blocks, py::arg("blocks"))   // host-mirror extension: chunked realisation
      .def("chunk_blocks", &SubbandBeamformer::chunk_blocks);

  py::class_<SubbandDS, SubbandBeamformer, SubbandDSPtr>(m, "SubbandDSPtr")
      .def(py::init([](unsigned fftlen, bool hbs, const std::string& nm) { return std::make_shared<SubbandDS>(fftlen, hbs, nm); }), py::arg("fftlen") = 512,
           py::arg("half_band_shift") = false, py::arg("nm") = "SubbandDS")
      .def("calc_array_manifold_vectors", [](SubbandDS& s, double fs, py::array_t<double, py::array::c_style | py::array::forcecast> d) { s.calc_array_manifold_vectors(fs, vec_d(d)); },
           py::arg("samplerate"), py::arg("delays"))
      .def("calc_array_manifold_vectors_2", [](SubbandDS& s, double fs, py::array_t<double, py::array::c_style | py::array::forcecast> dT, py::array_t<double, py::array::c_style | py::array::forcecast> dJ) { s.calc_array_manifold_vectors_2(fs, vec_d(dT), vec_d(dJ)); },
           py::arg("samplerate"), py::arg("delays_t"), py::arg("delays_j"))
      .def("calc_array_manifold_vectors_n", [](SubbandDS& s, double fs, py::array_t<double, py::array::c_style | py::array::forcecast> dT, py::array_t<double, py::array::c_style | py::array::forcecast> dJ, unsigned NC) { s.calc_array_manifold_vectors_n(fs, vec_d(dT), vec_d(dJ), NC); },
           py::arg("samplerate"), py::arg("delays_t"), py::arg("delays_j"), py::arg("NC") = 2);

  py::class_<SubbandGSC, SubbandDS, SubbandGSCPtr>(m, "SubbandGSCPtr")
      .def(py::init([](unsigned fftlen, bool hbs, const std::string& nm) { return std::make_shared<SubbandGSC>(fftlen, hbs, nm); }), py::arg("fftlen") = 512,
           py::arg("half_band_shift") = false, py::arg("nm") = "SubbandGSC")
      .def("calc_gsc_weights", [](SubbandGSC& s, double fs, py::array_t<double, py::array::c_style | py::array::forcecast> d) { s.calc_gsc_weights(fs, vec_d(d)); },
           py::arg("samplerate"), py::arg("delaysT"))
      .def("calc_gsc_weights_n", [](SubbandGSC& s, double fs, py::array_t<double, py::array::c_style | py::array::forcecast> dT,
                                    py::array_t<double, py::array::c_style | py::array::forcecast> dJ, unsigned NC) { s.calc_gsc_weights_n(fs, vec_d(dT), vec_d(dJ), NC); },
           py::arg("samplerate"), py::arg("delays_t"), py::arg("delays_j"), py::arg("NC") = 2)
      .def("calc_gsc_weights_2", [](SubbandGSC& s, double fs, py::array_t<double, py::array::c_style | py::array::forcecast> dT,
                                    py::array_t<double, py::array::c_style | py::array::forcecast> dJ) { s.calc_gsc_weights_2(fs, vec_d(dT), vec_d(dJ)); },
           py::arg("samplerate"), py::arg("delays_t"), py::arg("delays_j"))
      .def("set_active_weights_f", [](SubbandGSC& s, unsigned f, py::array_t<double, py::array::c_style | py::array::forcecast> w) { s.set_active_weights_f(f, vec_d(w)); },
           py::arg("fbinX"), py::arg("packedWeight"))
      .def("zero_active_weights", &SubbandGSC::zero_active_weights)
      .def("normalize_weight", &SubbandGSC::normalize_weight, py::arg("flag"))
      .def("set_quiescent_weights_f", [](SubbandGSC& s, unsigned f, py::array_t<cplx, py::array::c_style | py::array::forcecast> w) {
             s.set_quiescent_weights_f(f, std::vector<cplx>(w.data(), w.data() + w.size())); }, py::arg("fbinX"), py::arg("srcWq"))
      .def("write_fir_coeff", &SubbandGSC::write_fir_coeff, py::arg("fn"), py::arg("winType") = 1);

  py::class_<LmsConfig>(m, "LmsConfig")
      .def(py::init<>())
      .def_readwrite("beta", &LmsConfig::beta).def_readwrite("gamma", &LmsConfig::gamma).def_readwrite("init_diagonal_load", &LmsConfig::init_diagonal_load)
      .def_readwrite("regularization_param", &LmsConfig::regularization_param).def_readwrite("energy_floor", &LmsConfig::energy_floor)
      .def_readwrite("sil_thresh", &LmsConfig::sil_thresh).def_readwrite("max_wa_l2norm", &LmsConfig::max_wa_l2norm)
      .def_readwrite("min_frames", &LmsConfig::min_frames).def_readwrite("slowdown_after", &LmsConfig::slowdown_after);

  py::class_<SubbandGSCRLS, SubbandGSC, SubbandGSCRLSPtr>(m, "SubbandGSCRLSPtr")   // beamformer.i:289-339
      .def(py::init([](unsigned fftlen, bool hbs, float myu, float sigma2, const std::string& nm) { return std::make_shared<SubbandGSCRLS>(fftlen, hbs, myu, sigma2, nm); }),
           py::arg("fftlen"), py::arg("half_band_shift") = false, py::arg("myu") = 0.9f, py::arg("sigma2") = 0.01f, py::arg("nm") = "SubbandGSCRLS")
      .def("init_precision_matrix", &SubbandGSCRLS::init_precision_matrix, py::arg("sigma2") = 0.01f)
      .def("update_active_weight_vecotrs", &SubbandGSCRLS::update_active_weight_vecotrs, py::arg("flag"))
      .def("set_quadratic_constraint", &SubbandGSCRLS::set_quadratic_constraint, py::arg("alpha"), py::arg("qctype") = 1);

  py::class_<SubbandGSCLMS, SubbandDS, SubbandGSCLMSPtr>(m, "SubbandGSCLMSPtr")
      .def(py::init([](unsigned fftlen, const LmsConfig& c, const std::string& nm) { return std::make_shared<SubbandGSCLMS>(fftlen, c, nm); }), py::arg("fftlen"),
           py::arg("config"), py::arg("nm") = "SubbandGSCLMS")
      .def("calc_beamformer_weights", [](SubbandGSCLMS& s, double fs, py::array_t<double, py::array::c_style | py::array::forcecast> d) { s.calc_beamformer_weights(fs, vec_d(d)); },
           py::arg("samplerate"), py::arg("delays"))
      .def("active_weights", [](SubbandGSCLMS& s) {
        auto w = s.active_weights();
        py::array_t<std::complex<float>> a({(size_t)(s.fftLen() / 2 + 1), (size_t)(s.chanN() - 1)});
        std::memcpy(a.mutable_data(), w.data(), sizeof(std::complex<float>) * w.size());
        return a; })
      .def("total_updates", &SubbandGSCLMS::total_updates);

  py::class_<RlsConfig>(m, "RlsConfig")
      .def(py::init<>())
      .def_readwrite("beta", &RlsConfig::beta).def_readwrite("gamma", &RlsConfig::gamma).def_readwrite("mu", &RlsConfig::mu)
      .def_readwrite("init_diagonal_load", &RlsConfig::init_diagonal_load).def_readwrite("regularization_param", &RlsConfig::regularization_param)
      .def_readwrite("sil_thresh", &RlsConfig::sil_thresh).def_readwrite("alpha2", &RlsConfig::alpha2).def_readwrite("max_wa_l2norm", &RlsConfig::max_wa_l2norm)
      .def_readwrite("constraint_option", &RlsConfig::constraint_option).def_readwrite("min_frames", &RlsConfig::min_frames);

  py::class_<SubbandGSCRLSNative, SubbandDS, SubbandGSCRLSNativePtr>(m, "SubbandGSCRLSNativePtr")
      .def(py::init([](unsigned fftlen, const RlsConfig& c, const std::string& nm) { return std::make_shared<SubbandGSCRLSNative>(fftlen, c, nm); }), py::arg("fftlen"),
           py::arg("config"), py::arg("nm") = "SubbandGSCRLSNative")
      .def("calc_beamformer_weights", [](SubbandGSCRLSNative& s, double fs, py::array_t<double, py::array::c_style | py::array::forcecast> d) { s.calc_beamformer_weights(fs, vec_d(d)); },
           py::arg("samplerate"), py::arg("delays"))
      .def("active_weights", [](SubbandGSCRLSNative& s) {
        auto w = s.active_weights();
        py::array_t<std::complex<float>> a({(size_t)(s.fftLen() / 2 + 1), (size_t)(s.chanN() - 1)});
        std::memcpy(a.mutable_data(), w.data(), sizeof(std::complex<float>) * w.size());
        return a; })
      .def("total_updates", &SubbandGSCRLSNative::total_updates);

  py::class_<SubbandSOSNative, SubbandDS, SubbandSOSNativePtr>(m, "SubbandSOSNativePtr")
      .def(py::init([](unsigned fftlen, const std::string& nm) { return std::make_shared<SubbandSOSNative>(fftlen, nm); }), py::arg("fftlen"), py::arg("nm") = "SubbandSOSNative")
      .def("reset_stats", &SubbandSOSNative::reset_stats)
      .def("accu_stats_from_label", [](SubbandSOSNative& s, double fs, py::array_t<double, py::array::c_style | py::array::forcecast> labs, double thr) {
        s.accu_stats_from_label(fs, vec_d(labs), thr); }, py::arg("samplerate"), py::arg("target_labs"), py::arg("energy_threshold") = 10.0)
      .def("accu_stats_from_tfmask", [](SubbandSOSNative& s, double fs, py::array_t<float, py::array::c_style | py::array::forcecast> mt,
                                        py::array_t<float, py::array::c_style | py::array::forcecast> mj, double thr) {
        if (mt.ndim() != 2 || mj.ndim() != 2) throw jdimension_error("TF masks must be 2-D [frames][subbands]");
        s.accu_stats_from_tfmask(fs, std::vector<float>(mt.data(), mt.data() + mt.size()), std::vector<float>(mj.data(), mj.data() + mj.size()),
                                 (unsigned)mt.shape(0), (unsigned)mt.shape(1), thr); },
           py::arg("samplerate"), py::arg("mask_t"), py::arg("mask_j"), py::arg("energy_threshold") = 10.0)
      .def("calc_weights", &SubbandSOSNative::calc_weights, py::arg("kind"), py::arg("gamma") = 1.0e-6, py::arg("ref_micx") = 0, py::arg("offset") = 0.0)
      .def("frame_counts", [](SubbandSOSNative& s) {
        auto c = s.frame_counts();
        py::array_t<double> out({(py::ssize_t)(c.size() / 2), (py::ssize_t)2});
        std::memcpy(out.mutable_data(), c.data(), c.size() * sizeof(double));
        return out; });

  py::class_<SubbandMVDR, SubbandDS, SubbandMVDRPtr>(m, "SubbandMVDRPtr")
      .def(py::init([](unsigned fftlen, bool hbs, const std::string& nm) { return std::make_shared<SubbandMVDR>(fftlen, hbs, nm); }), py::arg("fftlen") = 512,
           py::arg("half_band_shift") = false, py::arg("nm") = "SubbandMVDR")
      .def("calc_mvdr_weights", &SubbandMVDR::calc_mvdr_weights, py::arg("samplerate"), py::arg("dthreshold") = 1.0e-8, py::arg("calc_inverse_matrix") = true)
      .def("mvdr_weights", &SubbandMVDR::mvdr_weights, py::arg("fbinX"))
      .def("set_noise_spatial_spectral_matrix", [](SubbandMVDR& s, unsigned f, py::array_t<std::complex<double>, py::array::c_style | py::array::forcecast> R) {
        return s.set_noise_spatial_spectral_matrix(f, std::vector<cplx>(R.data(), R.data() + R.size())); }, py::arg("fbinX"), py::arg("Rnn"))
      .def("set_diffuse_noise_model", [](SubbandMVDR& s, py::array_t<double, py::array::c_style | py::array::forcecast> mpos, double fs, double c) {
        return s.set_diffuse_noise_model(vec_d(mpos), fs, c); }, py::arg("micPositions"), py::arg("samplerate"), py::arg("sspeed") = 343740.0)
      .def("set_all_diagonal_loading", &SubbandMVDR::set_all_diagonal_loading, py::arg("diagonalWeight"))
      .def("set_diagonal_looading", &SubbandMVDR::set_diagonal_looading, py::arg("fbinX"), py::arg("diagonalWeight"))
      .def("divide_all_nondiagonal_elements", &SubbandMVDR::divide_all_nondiagonal_elements, py::arg("mu"))
      .def("divide_nondiagonal_elements", &SubbandMVDR::divide_nondiagonal_elements, py::arg("fbinX"), py::arg("mu"))
      .def("accumulate_noise_covariance", &SubbandMVDR::accumulate_noise_covariance, py::arg("samplerate"), py::arg("label_start"), py::arg("label_end"),
           py::arg("energy_threshold"));

  py::class_<SubbandMVDRGSC, SubbandMVDR, SubbandMVDRGSCPtr>(m, "SubbandMVDRGSCPtr")
      .def(py::init([](unsigned fftlen, bool hbs, const std::string& nm) { return std::make_shared<SubbandMVDRGSC>(fftlen, hbs, nm); }), py::arg("fftlen") = 512,
           py::arg("half_band_shift") = false, py::arg("nm") = "SubbandMVDR")
      .def("set_active_weights_f", [](SubbandMVDRGSC& s, unsigned f, py::array_t<double, py::array::c_style | py::array::forcecast> w) { s.set_active_weights_f(f, vec_d(w)); },
           py::arg("fbinX"), py::arg("packedWeight"))
      .def("zero_active_weights", &SubbandMVDRGSC::zero_active_weights)
      .def("calc_blocking_matrix1", [](SubbandMVDRGSC& s, double fs, py::array_t<double, py::array::c_style | py::array::forcecast> d) { return s.calc_blocking_matrix1(fs, vec_d(d)); },
           py::arg("samplerate"), py::arg("delaysT"))
      .def("calc_blocking_matrix2", &SubbandMVDRGSC::calc_blocking_matrix2)
      .def("upgrade_blocking_matrix", &SubbandMVDRGSC::upgrade_blocking_matrix)
      .def("blocking_matrix_output", [](SubbandMVDRGSC& s, int outChanX) { const cplx* p = s.blocking_matrix_output(outChanX); return view<cplx>(s, p, s.size()); },
           py::arg("outChanX") = 0);

  // beamformer/beamformer.i:543-568
  py::class_<SubbandOrthogonalizer, VectorComplexFeatureStream, SubbandOrthogonalizerPtr>(m, "SubbandOrthogonalizerPtr")
      .def(py::init([](SubbandMVDRGSCPtr beamformer, int outChanX, const std::string& nm) { return std::make_shared<SubbandOrthogonalizer>(beamformer, outChanX, nm); }),
           py::arg("beamformer"), py::arg("outChanX") = 0, py::arg("nm") = "SubbandOrthogonalizer");

  py::class_<ZelinskiPostFilter, VectorComplexFeatureStream, ZelinskiPostFilterPtr>(m, "ZelinskiPostFilterPtr")
      .def(py::init<const VectorComplexFeatureStreamPtr&, unsigned, double, int, int, const std::string&>(), py::arg("output"), py::arg("M"), py::arg("alpha") = 0.6,
           py::arg("type") = 2, py::arg("min_frames") = 0, py::arg("nm") = "ZelinskPostFilter")
      .def("set_beamformer", &ZelinskiPostFilter::set_beamformer, py::arg("beamformer"))
      .def("set_snapshot_array", &ZelinskiPostFilter::set_snapshot_array, py::arg("snapShotArray"))
      .def("set_array_manifold_vector", [](ZelinskiPostFilter& s, unsigned fbinX, py::array_t<std::complex<double>, py::array::c_style | py::array::forcecast> v, bool hbs, unsigned NC) {
             s.set_array_manifold_vector(fbinX, std::vector<cplx>(v.data(), v.data() + v.size()), hbs, NC); },
           py::arg("fbinX"), py::arg("arrayManifoldVector"), py::arg("halfBandShift") = false, py::arg("NC") = 1)
      .def("postfilter_weights", &ZelinskiPostFilter::postfilter_weights);

  py::class_<McCowanPostFilter, ZelinskiPostFilter, McCowanPostFilterPtr>(m, "McCowanPostFilterPtr")
      .def(py::init<const VectorComplexFeatureStreamPtr&, unsigned, double, int, int, float, const std::string&>(), py::arg("output"), py::arg("fftlen"),
           py::arg("alpha") = 0.6, py::arg("type") = 2, py::arg("min_frames") = 0, py::arg("threshold") = 0.99f, py::arg("nm") = "McCowanPostFilterPtr")
      .def("noise_spatial_spectral_matrix", [](McCowanPostFilter& s, unsigned fbinX) {
             std::vector<cplx> R = s.noise_spatial_spectral_matrix(fbinX);
             py::array_t<cplx> out({(py::ssize_t)s.chanN(), (py::ssize_t)s.chanN()});
             std::memcpy(out.mutable_data(), R.data(), R.size() * sizeof(cplx));
             return out; }, py::arg("fbinX"))
      .def("set_noise_spatial_spectral_matrix", [](McCowanPostFilter& s, unsigned fbinX, py::array_t<cplx, py::array::c_style | py::array::forcecast> Rnn) {
             if (Rnn.ndim() != 2) throw jdimension_error("Rnn must be a matrix");
             std::vector<cplx> v(Rnn.data(), Rnn.data() + Rnn.size());
             return s.set_noise_spatial_spectral_matrix(fbinX, v, (unsigned)Rnn.shape(0), (unsigned)Rnn.shape(1)); }, py::arg("fbinX"), py::arg("Rnn"))
      .def("set_diffuse_noise_model", [](McCowanPostFilter& s, py::array_t<double, py::array::c_style | py::array::forcecast> mpos, double samplerate, double sspeed) {
             if (mpos.ndim() != 2) throw jdimension_error("micPositions must be a matrix");
             std::vector<double> v(mpos.data(), mpos.data() + mpos.size());
             return s.set_diffuse_noise_model(v, (unsigned)mpos.shape(0), (unsigned)mpos.shape(1), samplerate, sspeed); },
           py::arg("micPositions"), py::arg("sampleRate"), py::arg("sspeed") = 343740.0)
      .def("set_all_diagonal_loading", &McCowanPostFilter::set_all_diagonal_loading, py::arg("diagonalWeight"))
      .def("set_diagonal_looading", &McCowanPostFilter::set_diagonal_looading, py::arg("fbinX"), py::arg("diagonalWeight"))
      .def("divide_all_nondiagonal_elements", &McCowanPostFilter::divide_all_nondiagonal_elements, py::arg("mu"))
      .def("divide_nondiagonal_elements", &McCowanPostFilter::divide_nondiagonal_elements, py::arg("fbinX"), py::arg("mu"));

  py::class_<LefkimmiatisPostFilter, McCowanPostFilter, LefkimmiatisPostFilterPtr>(m, "LefkimmiatisPostFilterPtr")
      .def(py::init<const VectorComplexFeatureStreamPtr&, unsigned, double, unsigned, double, int, int, float, const std::string&>(), py::arg("output"), py::arg("fftlen"),
           py::arg("min_sv") = 1.0e-8, py::arg("fbin_no1") = 0, py::arg("alpha") = 0.6, py::arg("type") = 2, py::arg("min_frames") = 0, py::arg("threshold") = 0.99f,
           py::arg("nm") = "LefkimmiatisPostFilterPtr")
      .def("calc_inverse_noise_spatial_spectral_matrix", &LefkimmiatisPostFilter::calc_inverse_noise_spatial_spectral_matrix);

  m.def("calc_all_delays", [](double x, double y, double z, py::array_t<double, py::array::c_style | py::array::forcecast> mpos) { return calc_all_delays(x, y, z, vec_d(mpos)); },
        py::arg("x"), py::arg("y"), py::arg("z"), py::arg("mpos"));
}
