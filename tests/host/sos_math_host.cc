// CPU harness around csrc/btkb_sos_math.cuh (the per-chain fp64 solve of the SOS beamformers, the same source the CUDA
// kernel k_sos_solve compiles): reads problems from stdin, writes weights to stdout.  Built and driven by
// tests/test_sos_math_host.py; test infrastructure only.
//   input : kind C n gamma ref_micx offset, then n x (Rt[C][C] re im, Rn[C][C] re im) as text doubles
//   output: n x (ok, w[C] re im)
#include <cstdio>
#include <vector>
#include "../../distant_speech_recognition_b200/csrc/btkb_sos_math.cuh"
using namespace btkb;

template <int C>
static void run(int kind, int n, double gamma, int ref, double offset) {
  for (int q = 0; q < n; q++) {
    zd Rt[C][C], Rn[C][C], w[C];
    for (int i = 0; i < C; i++) for (int j = 0; j < C; j++) if (scanf("%lf %lf", &Rt[i][j].x, &Rt[i][j].y) != 2) return;
    for (int i = 0; i < C; i++) for (int j = 0; j < C; j++) if (scanf("%lf %lf", &Rn[i][j].x, &Rn[i][j].y) != 2) return;
    for (int c = 0; c < C; c++) w[c] = zmk(0, 0);
    const bool ok = sos_solve_chain<C>(Rt, Rn, kind, gamma, ref, offset, w);
    printf("%d", ok ? 1 : 0);
    for (int c = 0; c < C; c++) printf(" %.17g %.17g", w[c].x, w[c].y);
    printf("\n");
  }
}

int main() {
  int kind, C, n, ref; double gamma, offset;
  if (scanf("%d %d %d %lf %d %lf", &kind, &C, &n, &gamma, &ref, &offset) != 6) return 2;
  if (C == 2) run<2>(kind, n, gamma, ref, offset);
  else if (C == 4) run<4>(kind, n, gamma, ref, offset);
  else if (C == 8) run<8>(kind, n, gamma, ref, offset);
  else return 3;
  return 0;
}
