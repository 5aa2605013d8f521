// CPU execution of the kernels' FFT code (csrc/btkb_fft.cuh, the same source k_analysis_r1 / k_synthesis_fast compile for sm_100a):
// the NT = M/8 "threads" of one transform pair run as std::threads, the shared-memory buffers are plain arrays and the barrier is a
// std::barrier.  Both the scalar path (PK = false, the one measured on B200) and the packed 2 x fp32 path (PK = true, btkb_f2.cuh;
// on the host its primitives are plain scalar code with fmaf) are run on the same input; the driver (tests/test_fft_packed_host.py)
// requires them to agree BIT FOR BIT and checks both against a double-precision DFT.  Test infrastructure only.
//   argv: M SIGN seed;  stdout: M lines "re0 im0 re1 im1  re0p im0p re1p im1p" (hex floats) = spectra of the two transforms, scalar
//   then packed, followed by one line "fold <max abs difference of the packed vs scalar polyphase MAC> primitives <mismatches> regs_differ <0|1> nlms <mismatches>"
#include <barrier>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <thread>
#include <vector>

// device-only intrinsics the header uses in code paths exercised here
template <class T> static inline T __ldg(const T* p) { return *p; }
static inline void sincospif(float x, float* s, float* c) { *s = (float)std::sin(M_PI * (double)x); *c = (float)std::cos(M_PI * (double)x); }

#include "../../distant_speech_recognition_b200/csrc/btkb_fft.cuh"
#include "../../distant_speech_recognition_b200/csrc/btkb_nlms_math.cuh"
using namespace btkb;

// The NLMS recurrence of k_perbin<C, LMS> (btkb_nlms_math.cuh): 40 consecutive adaptation steps of one chain with the scalar and
// with the packed arithmetic, states carried separately; every intermediate (Yc, u.x, u, ||u||^2 partial sums) must agree bit for bit.
template <int C>
static int nlms_mismatches(unsigned seed) {
  std::mt19937 rng(seed);
  std::normal_distribution<float> nd(0.f, 1.f);
  auto ne = [](float a, float b) { unsigned x, y; std::memcpy(&x, &a, 4); std::memcpy(&y, &b, 4); return x != y; };
  int bad = 0;
  float2 w[C], us[C], up[C];
  for (int c = 0; c < C; c++) { const float ph = 3.f * nd(rng); w[c] = make_float2(std::cos(ph) / C, std::sin(ph) / C); us[c] = up[c] = make_float2(0.f, 0.f); }
  for (int t = 0; t < 40; t++) {
    float2 x[C];
    for (int c = 0; c < C; c++) x[c] = make_float2(3000.f * nd(rng), 3000.f * nd(rng));
    const float2 ys = cdot<C, true, false>(x, w), yp = cdot<C, true, true>(x, w);
    bad += ne(ys.x, yp.x) + ne(ys.y, yp.y);
    const float2 uxs = cdot<C, false, false>(us, x), uxp = cdot<C, false, true>(up, x);
    bad += ne(uxs.x, uxp.x) + ne(uxs.y, uxp.y);
    const float gamma = 0.01f, sub = 1.0e6f + 1.0e7f * std::fabs(nd(rng)), reg = (t & 1) ? 1.0e-4f : 0.f;
    float a0, a1, b0, b1;
    nlms_adapt_step<C, false>(x, w, us, ys, gamma, sub, reg, a0, a1);
    nlms_adapt_step<C, true>(x, w, up, yp, gamma, sub, reg, b0, b1);
    bad += ne(a0, b0) + ne(a1, b1);
    for (int c = 0; c < C; c++) bad += ne(us[c].x, up[c].x) + ne(us[c].y, up[c].y);
  }
  return bad;
}

// The Zelinski post-filter's per-frame statistics (btkb_nlms_math.cuh zelinski_csd_step): 30 frames, alpha = 0 for the first two,
// the last three past the end of the utterance (live = false); CSDs, PSDs and the two sums must agree bit for bit.
template <int C>
static int zelinski_mismatches(unsigned seed) {
  std::mt19937 rng(seed);
  std::normal_distribution<float> nd(0.f, 1.f);
  auto ne = [](float a, float b) { unsigned x, y; std::memcpy(&x, &a, 4); std::memcpy(&y, &b, 4); return x != y; };
  constexpr int NP = C * (C - 1) / 2;
  int bad = 0;
  float2 ta[C], cs[NP], cp[NP];
  float ps[C], pp[C];
  for (int c = 0; c < C; c++) { const float ph = 3.f * nd(rng); ta[c] = make_float2(std::cos(ph) / C, std::sin(ph) / C); ps[c] = pp[c] = 0.f; }
  for (int i = 0; i < NP; i++) cs[i] = cp[i] = make_float2(0.f, 0.f);
  for (int t = 0; t < 30; t++) {
    float2 x[C];
    for (int c = 0; c < C; c++) x[c] = make_float2(3000.f * nd(rng), 3000.f * nd(rng));
    const float al = (t >= 2) ? 0.7f : 0.f;
    const bool live = t < 27;
    float2 s0, s1; float d0, d1;
    zelinski_csd_step<C, false>(x, ta, cs, ps, al, live, s0, d0);
    zelinski_csd_step<C, true>(x, ta, cp, pp, al, live, s1, d1);
    bad += ne(s0.x, s1.x) + ne(s0.y, s1.y) + ne(d0, d1);
    for (int i = 0; i < NP; i++) bad += ne(cs[i].x, cp[i].x) + ne(cs[i].y, cp[i].y);
    for (int c = 0; c < C; c++) bad += ne(ps[c], pp[c]);
  }
  return bad;
}

// The RLS sidelobe canceller's adaptation step (btkb_nlms_math.cuh rls_core_step): 40 consecutive steps of one chain from the initial
// Pt = (I - C v v^H) / load, regularisation on every other step; Pt (diagonal and lower triangle) and the new u must agree bit for bit.
template <int C>
static int rls_mismatches(unsigned seed) {
  std::mt19937 rng(seed);
  std::normal_distribution<float> nd(0.f, 1.f);
  auto ne = [](float a, float b) { unsigned x, y; std::memcpy(&x, &a, 4); std::memcpy(&y, &b, 4); return x != y; };
  constexpr int NP = C * (C - 1) / 2;
  int bad = 0;
  float2 w[C], us[C], up[C];
  for (int c = 0; c < C; c++) { const float ph = 3.f * nd(rng); w[c] = make_float2(std::cos(ph) / C, std::sin(ph) / C); us[c] = up[c] = make_float2(0.f, 0.f); }
  HermP<C> Ps, Pp;
  const float inv_load = 1.0f / 1.0e6f;
  for (int i = 0; i < C; i++) {
    Ps.d[i] = (1.0f - (float)C * fmaf(w[i].x, w[i].x, w[i].y * w[i].y)) * inv_load;
    for (int j = 0; j < i; j++)
      Ps.o[HermP<C>::idx(i, j)] = make_float2(-(float)C * fmaf(w[i].x, w[j].x, w[i].y * w[j].y) * inv_load, -(float)C * fmaf(w[i].y, w[j].x, -w[i].x * w[j].y) * inv_load);
  }
  Pp = Ps;
  const float mu = 0.9999f, inv_mu = 1.0f / mu, gamma = 1.0f;
  for (int t = 0; t < 40; t++) {
    float2 x[C], ns[C], np[C];
    for (int c = 0; c < C; c++) x[c] = make_float2(3000.f * nd(rng), 3000.f * nd(rng));
    const float2 ys = cdot<C, true, false>(x, w), yp = cdot<C, true, true>(x, w);
    const float reg = (t & 1) ? 1.0e-2f : 0.f;
    rls_core_step<C, false>(x, w, us, ys, Ps, mu, inv_mu, gamma, reg, ns);
    rls_core_step<C, true>(x, w, up, yp, Pp, mu, inv_mu, gamma, reg, np);
    for (int c = 0; c < C; c++) { bad += ne(ns[c].x, np[c].x) + ne(ns[c].y, np[c].y) + ne(Ps.d[c], Pp.d[c]); us[c] = ns[c]; up[c] = np[c]; }
    for (int i = 0; i < NP; i++) bad += ne(Ps.o[i].x, Pp.o[i].x) + ne(Ps.o[i].y, Pp.o[i].y);
    for (int c = 0; c < C; c++) if (!std::isfinite(ns[c].x) || !std::isfinite(ns[c].y)) bad += 1000;   // a chain that blew up proves nothing
  }
  return bad;
}

static int regs_differ = 0;

template <int M, int SIGN, bool PK>
static void run_pair(const std::vector<float2>& in0, const std::vector<float2>& in1, const std::vector<float2>& tab, std::vector<float2>& out0,
                     std::vector<float2>& out1) {
  using Plan = FftPlan<M>;
  constexpr int NT = Plan::NT, R0 = Plan::R0, NB = 8 / R0;
  std::vector<float2> buf0(Plan::BUF), buf1(Plan::BUF);
  std::barrier bar(NT);
  out0.assign(M, make_float2(0, 0)); out1.assign(M, make_float2(0, 0));
  auto body = [&](int tg) {
    FftTwiddles<M, SIGN> tw;
    tw.init_from_table(tg, tab.data());
    float2 v0[8], v1[8];
    for (int b = 0; b < NB; b++)
      for (int r = 0; r < R0; r++) { v0[b * R0 + r] = in0[(tg + b * NT) + r * (M / R0)]; v1[b * R0 + r] = in1[(tg + b * NT) + r * (M / R0)]; }
    auto sync = [&] { bar.arrive_and_wait(); };
    fft_first_pass<M, SIGN, PK>(v0, buf0.data(), tg);
    fft_first_pass<M, SIGN, PK>(v1, buf1.data(), tg);
    sync();
    FftPassChain<M, SIGN, 0, decltype(sync), PK>::run(v0, v1, buf0.data(), buf1.data(), tg, tw, sync);
    // lower half of the spectrum stays in registers (v[r] = Z[tg + r NT], r < 4), the upper half is in the buffers in natural order
    for (int r = 0; r < 4; r++) { out0[tg + r * NT] = v0[r]; out1[tg + r * NT] = v1[r]; }
    for (int r = 4; r < 8; r++) { out0[tg + r * NT] = buf0[tg + r * NT]; out1[tg + r * NT] = buf1[tg + r * NT]; }
    // ... and is also still in the registers (k_synthesis_fast<PK> writes all eight values from v[])
    for (int r = 4; r < 8; r++)
      if (std::memcmp(&v0[r], &buf0[tg + r * NT], sizeof(float2)) || std::memcmp(&v1[r], &buf1[tg + r * NT], sizeof(float2))) regs_differ = 1;
  };
  std::vector<std::thread> th;
  for (int t = 0; t < NT; t++) th.emplace_back(body, t);
  for (auto& t : th) t.join();
}

static unsigned bits(float f) { unsigned u; std::memcpy(&u, &f, 4); return u; }

template <int M, int SIGN>
static int run(unsigned seed) {
  std::mt19937 rng(seed);
  std::normal_distribution<float> nd(0.f, 3000.f);
  std::vector<float2> in0(M), in1(M), tab(M);
  for (int i = 0; i < M; i++) { in0[i] = make_float2(nd(rng), nd(rng)); in1[i] = make_float2(nd(rng), nd(rng)); }
  if (seed < 3) {   // silent and almost silent frames: all (signed) zeros, or one sample: the zeros of the spectrum must keep their signs too
    for (int i = 0; i < M; i++) { in0[i] = make_float2(seed == 1 ? -0.f : 0.f, 0.f); in1[i] = make_float2(0.f, (i & 1) ? -0.f : 0.f); }
    if (seed == 2) { in0[3] = make_float2(1000.f, 0.f); in1[M - 1] = make_float2(0.f, -7.f); }
  }
  for (int i = 0; i < M; i++) { const double a = 2.0 * M_PI * i / M; tab[i] = make_float2((float)std::cos(a), (float)std::sin(a)); }   // btkb_api.cu:207-214
  std::vector<float2> s0, s1, p0, p1;
  run_pair<M, SIGN, false>(in0, in1, tab, s0, s1);
  run_pair<M, SIGN, true>(in0, in1, tab, p0, p1);
  for (int i = 0; i < M; i++) printf("%a %a %a %a %a %a %a %a\n", in0[i].x, in0[i].y, in1[i].x, in1[i].y, 0.0, 0.0, 0.0, 0.0);
  for (int i = 0; i < M; i++) printf("%a %a %a %a %a %a %a %a\n", s0[i].x, s0[i].y, s1[i].x, s1[i].y, p0[i].x, p0[i].y, p1[i].x, p1[i].y);
  // the polyphase MAC of k_analysis_r1: tap h times the (channel a, channel b) sample pair, accumulated (scalar vs f2_fma_s)
  float dmax = 0.f;
  for (int trial = 0; trial < 2000; trial++) {
    float2 acc_s = make_float2(0.f, 0.f), acc_p = acc_s;
    for (int k = 0; k < 4; k++) {
      const float h = nd(rng) * 1e-4f; const float2 s = make_float2(nd(rng), nd(rng));
      acc_s.x = fmaf(h, s.x, acc_s.x); acc_s.y = fmaf(h, s.y, acc_s.y);
      acc_p = f2_fma_s(s, h, acc_p);
    }
    dmax = std::fmax(dmax, std::fmax(std::fabs(acc_s.x - acc_p.x), std::fabs(acc_s.y - acc_p.y)));
  }
  // every primitive against the scalar expression it replaces
  int bad = 0;
  for (int trial = 0; trial < 20000; trial++) {
    float2 a = make_float2(nd(rng), nd(rng)), b = make_float2(nd(rng), nd(rng));
    float sc = nd(rng);
    if (trial < 4096) {   // signed zeros in every combination (silent / zero-padded frames): the packed forms must keep the SIGN of a zero too
      const float z[4] = {0.f, -0.f, 1.5f, -2.25f};
      a = make_float2(z[trial & 3], z[(trial >> 2) & 3]); b = make_float2(z[(trial >> 4) & 3], z[(trial >> 6) & 3]);
      sc = z[(trial >> 8) & 3] * ((trial >> 10) & 1 ? 1.f : 0.5f) + (((trial >> 11) & 1) ? 0.f : 0.f);
    }
    auto ne = [&](float2 x, float2 y) { return bits(x.x) != bits(y.x) || bits(x.y) != bits(y.y); };
    bad += ne(f2_add(a, b), cadd(a, b));
    bad += ne(f2_sub(a, b), csub(a, b));
    bad += ne(f2_cmul(a, b), cmul(a, b));
    bad += ne(f2_add_ib<+1>(a, b), cadd(a, mul_si<+1>(b)));
    bad += ne(f2_add_ib<-1>(a, b), cadd(a, mul_si<-1>(b)));
    bad += ne(f2_sub_ib<+1>(a, b), csub(a, mul_si<+1>(b)));
    bad += ne(f2_sub_ib<-1>(a, b), csub(a, mul_si<-1>(b)));
    bad += ne(f2_scale(a, sc), make_float2(a.x * sc, a.y * sc));
    bad += ne(f2_fma(a, b, make_float2(sc, -sc)), make_float2(fmaf(a.x, b.x, sc), fmaf(a.y, b.y, -sc)));
    bad += ne(f2_cmulc(a, b), cmulc(a, b));
    { float2 r1 = make_float2(sc, -sc), r2 = r1; cmac(r1, a, b); bad += ne(f2_cmac(r2, a, b), r1); }              // btkb_nlms_math.cuh cmac
    { float2 r1 = make_float2(sc, -sc), r2 = r1; cmac_conj(r1, a, b); bad += ne(f2_cmac_conj(r2, a, b), r1); }    // ... cmac_conj
    // the untangle of k_analysis_r1: A = (zk + conj zm)/2, B = (zk - conj zm)/(2i), exactly as the scalar kernel writes them
    bad += ne(f2_scale(f2_add_conj(a, b), 0.5f), make_float2(0.5f * (a.x + b.x), 0.5f * (a.y - b.y)));
    bad += ne(f2_scale_mi(f2_sub_conj(a, b), 0.5f), make_float2(0.5f * (a.y + b.y), -0.5f * (a.x - b.x)));
  }
  const int nl = nlms_mismatches<2>(seed) + nlms_mismatches<4>(seed + 1) + nlms_mismatches<8>(seed + 2) + zelinski_mismatches<2>(seed + 3) +
                 zelinski_mismatches<4>(seed + 4) + zelinski_mismatches<8>(seed + 5) + rls_mismatches<2>(seed + 6) + rls_mismatches<4>(seed + 7) +
                 rls_mismatches<8>(seed + 8);
  printf("fold %a primitives %d regs_differ %d nlms %d\n", dmax, bad, regs_differ, nl);
  return 0;
}

int main(int argc, char** argv) {
  if (argc < 4) return 2;
  const int M = atoi(argv[1]), sign = atoi(argv[2]);
  const unsigned seed = (unsigned)atoi(argv[3]);
#define CASE(MM) if (M == MM) return sign > 0 ? run<MM, +1>(seed) : run<MM, -1>(seed);
  CASE(256) CASE(512) CASE(1024) CASE(2048)
  return 2;
}
