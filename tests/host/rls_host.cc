// CPU execution of the RLS adaptation step the kernel k_perbin_rls compiles (csrc/btkb_nlms_math.cuh rls_core_step, scalar and packed
// forms) on one (utterance, bin) chain, so that tests/test_fft_packed_host.py can compare the fp32 recursion against an fp64 NumPy
// evaluation of the projector form (oracle/restate.py gsc_rls_projector; lib/pybeamformer.py:817-901).  Test infrastructure only.
//   argv: C T mu gamma reg load;  stdin: C complex weights v, then T x C complex snapshots (text, "re im" per value)
//   stdout: per step "yc.re yc.im e.re e.im" + the 2C components of u, scalar form; last line "packed_mismatches <n>"
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

template <class T> static inline T __ldg(const T* p) { return *p; }
static inline void sincospif(float x, float* s, float* c) { *s = (float)std::sin(M_PI * (double)x); *c = (float)std::cos(M_PI * (double)x); }

#include "../../distant_speech_recognition_b200/csrc/btkb_fft.cuh"
#include "../../distant_speech_recognition_b200/csrc/btkb_nlms_math.cuh"
using namespace btkb;

template <int C>
static void init_P(HermP<C>& P, const float2* w, float inv_load) {   // k_perbin_rls reset_P
  for (int i = 0; i < C; i++) {
    P.d[i] = (1.0f - (float)C * fmaf(w[i].x, w[i].x, w[i].y * w[i].y)) * inv_load;
    for (int j = 0; j < i; j++)
      P.o[HermP<C>::idx(i, j)] = make_float2(-(float)C * fmaf(w[i].x, w[j].x, w[i].y * w[j].y) * inv_load, -(float)C * fmaf(w[i].y, w[j].x, -w[i].x * w[j].y) * inv_load);
  }
}

template <int C>
static int run(int T, float mu, float gamma, float reg, float load) {
  float2 w[C], us[C], up[C];
  for (int c = 0; c < C; c++) { if (scanf("%f %f", &w[c].x, &w[c].y) != 2) return 2; us[c] = up[c] = make_float2(0.f, 0.f); }
  HermP<C> Ps, Pp;
  init_P<C>(Ps, w, 1.0f / load); Pp = Ps;
  int bad = 0;
  for (int t = 0; t < T; t++) {
    float2 x[C], ns[C], np_[C];
    for (int c = 0; c < C; c++) if (scanf("%f %f", &x[c].x, &x[c].y) != 2) return 2;
    const float2 ys = cdot<C, true, false>(x, w), yp = cdot<C, true, true>(x, w);
    rls_core_step<C, false>(x, w, us, ys, Ps, mu, 1.0f / mu, gamma, reg, ns);
    rls_core_step<C, true>(x, w, up, yp, Pp, mu, 1.0f / mu, gamma, reg, np_);
    for (int c = 0; c < C; c++) { bad += std::memcmp(&ns[c], &np_[c], sizeof(float2)) != 0; us[c] = ns[c]; up[c] = np_[c]; }
    const float2 e = csub(ys, cdot<C, false, false>(us, x));   // a-posteriori output of the frame (min_frames = 0)
    printf("%a %a %a %a", ys.x, ys.y, e.x, e.y);
    for (int c = 0; c < C; c++) printf(" %a %a", us[c].x, us[c].y);
    printf("\n");
  }
  printf("packed_mismatches %d\n", bad);
  return 0;
}

int main(int argc, char** argv) {
  if (argc < 7) return 2;
  const int C = atoi(argv[1]), T = atoi(argv[2]);
  const float mu = (float)atof(argv[3]), gamma = (float)atof(argv[4]), reg = (float)atof(argv[5]), load = (float)atof(argv[6]);
  if (C == 2) return run<2>(T, mu, gamma, reg, load);
  if (C == 4) return run<4>(T, mu, gamma, reg, load);
  if (C == 8) return run<8>(T, mu, gamma, reg, load);
  return 2;
}
