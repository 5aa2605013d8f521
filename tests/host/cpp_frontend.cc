// A C++ user of the host mirror, written the way the reference's C++ programs drive the stream graph (btk20_src/src/filterBankTest.cc:
// 189-196: `try { while (true) next(); } catch (jiterator_error&) {}`): 4 x SampleFeature -> OverSampledDFTAnalysisBank ->
// SubbandGSCLMS (the native body of pybeamformer.SubbandGSCLMSBeamformer) -> OverSampledDFTSynthesisBank, no Python anywhere.
// Built and driven by tests/test_zz_host_surface.py; test infrastructure only.
//   argv: prototype file (text: M m r, then m*M analysis taps, then m*M synthesis taps), samples file (text: C n, then C*n samples),
//         delays (C doubles on the command line)
//   stdout: "frames F blocks B energy E" on success; "j_error: <message>" when the library reports an error (e.g. no CUDA device).
#include <cstdio>
#include <cstdlib>
#include <memory>
#include <vector>
#include "../../distant_speech_recognition_b200/csrc/host/btk20_host.h"
using namespace btk20;

int main(int argc, char** argv) {
  if (argc < 3) return 2;
  FILE* fp = fopen(argv[1], "r");
  if (!fp) return 2;
  unsigned M, m, r;
  if (fscanf(fp, "%u %u %u", &M, &m, &r) != 3) return 2;
  std::vector<double> h((size_t)m * M), g((size_t)m * M);
  for (auto& v : h) if (fscanf(fp, "%lf", &v) != 1) return 2;
  for (auto& v : g) if (fscanf(fp, "%lf", &v) != 1) return 2;
  fclose(fp);
  fp = fopen(argv[2], "r");
  if (!fp) return 2;
  unsigned C, n;
  if (fscanf(fp, "%u %u", &C, &n) != 2) return 2;
  if (argc < 3 + (int)C) return 2;
  const unsigned D = M >> r;
  try {
    std::vector<SampleFeaturePtr> feats;
    std::vector<double> x(n);
    auto bf = std::make_shared<SubbandGSCLMS>(M, LmsConfig());
    for (unsigned c = 0; c < C; c++) {
      for (auto& v : x) if (fscanf(fp, "%lf", &v) != 1) return 2;
      auto sf = std::make_shared<SampleFeature>("", D, D, true);
      sf->set_samples(x.data(), n, 16000);
      feats.push_back(sf);
      bf->set_channel(std::make_shared<OverSampledDFTAnalysisBank>(sf, h, M, m, r, 2));
    }
    fclose(fp);
    std::vector<double> delays(C);
    for (unsigned c = 0; c < C; c++) delays[c] = atof(argv[3 + c]);
    bf->calc_beamformer_weights(16000.0, delays);
    OverSampledDFTSynthesisBank sfb(bf, g, M, m, r, 2);
    double energy = 0.0;
    int blocks = 0;
    try {
      while (true) {
        const float* b = sfb.next();
        for (unsigned i = 0; i < D; i++) energy += (double)b[i] * b[i];
        blocks++;
      }
    } catch (jiterator_error&) {
    }
    printf("frames %d blocks %d energy %.9e updates %d\n", bf->frames(), blocks, energy, bf->total_updates());
  } catch (j_error& e) {
    printf("j_error: %s\n", e.what());
    return 1;
  }
  return 0;
}
