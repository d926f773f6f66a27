// Host emulation of the four-step FFT core (csrc/fft2d_core.cuh + fft_reg.cuh): each lane task is executed in a
// loop; steps are separated where the kernel has __syncwarp().  stdin: "mode N M" + data; stdout: result.
#include <cstdio>
#include <cmath>
#include <vector>
#define __host__
#define __device__
#define __forceinline__ inline
#define __restrict__
#include "FFT2D_CORE_HOST"
using namespace msfno;

template <int P1, int P2, int NZ = 0>
static int run(int mode, int N, int M) {
  constexpr int H = P1 * P2, WP = WorkPitch<P2>::value;
  std::vector<cf> tw(H), tw2(M + 1), raw(H), work(P1 * WP), xs(H), out(H);
  for (int t = 0; t < H; ++t) { double a = -2.0 * M_PI * t / H; tw[t] = cf{(float)cos(a), (float)sin(a)}; }
  for (int m = 0; m <= M; ++m) { double a = -2.0 * M_PI * m / N; tw2[m] = cf{(float)cos(a), (float)sin(a)}; }
  const int mlim = M;
  if (mode == 0) {
    for (int i = 0; i < H; ++i) if (scanf("%f %f", &raw[i].x, &raw[i].y) != 2) return 1;
    for (int n2 = 0; n2 < P2; ++n2) fft2d_step1<P1, P2, -1>(raw.data(), P2, work.data(), tw.data(), n2);
    for (int k1 = 0; k1 < P1; ++k1) {
      cf v[P2];
      fft2d_step2<P1, P2, -1>(work.data(), k1, v);
      for (int k2 = 0; k2 < P2; ++k2) { int xi = xs_index(k1 + P1 * k2, H, mlim); if (xi >= 0) xs[xi] = v[k2]; }
    }
    for (int m = 0; m < M; ++m) { cf X = r2c_split_xs(xs.data(), tw2.data(), H, mlim, m); printf("%.9g %.9g\n", X.x, X.y); }
  } else {
    std::vector<cf> X(M);
    for (int i = 0; i < M; ++i) if (scanf("%f %f", &X[i].x, &X[i].y) != 2) return 1;
    auto Xh = [&](int q) -> cf { if (q >= mlim) return cf{0, 0}; cf v = X[q]; if (q == 0 || q == H) v.y = 0; return v; };
    for (auto& w : work) w = cf{NAN, NAN};   // rows the pruned variant must never read
    const int NBUILD = (NZ > 0) ? 2 * NZ * P2 : H;
    for (int kb = 0; kb < NBUILD; ++kb) {
      const int k = (NZ > 0 && kb >= NZ * P2) ? kb + (P1 - 2 * NZ) * P2 : kb;
      const int kk = H - k;
      cf o{0, 0};
      if (k < mlim || kk < mlim) {
        cf w;
        if (k <= mlim) w = tw2[k]; else { w = tw2[kk]; w.x = -w.x; }
        o = c2r_merge(Xh(k), Xh(kk), w);
      }
      work[(k / P2) * WP + (k % P2)] = o;   // Zt laid out as [n1][n2] with the padded pitch
    }
    for (int n2 = 0; n2 < P2; ++n2) fft2d_step1<P1, P2, +1, (NZ > 0 ? NZ : 0), (NZ > 0 ? P1 - NZ : 0)>(work.data(), WP, work.data(), tw.data(), n2);
    for (int k1 = 0; k1 < P1; ++k1) {
      cf v[P2];
      fft2d_step2<P1, P2, +1>(work.data(), k1, v);
      for (int k2 = 0; k2 < P2; ++k2) out[k1 + P1 * k2] = v[k2];
    }
    for (int i = 0; i < H; ++i) printf("%.9g\n%.9g\n", out[i].x, out[i].y);
  }
  return 0;
}

int main() {
  int mode, N, M;
  if (scanf("%d %d %d", &mode, &N, &M) != 3) return 1;
  switch (N / 2) {
    case 720: return (M <= 120) ? run<24, 30, 4>(mode, N, M) : run<24, 30>(mode, N, M);
    case 120: return run<15, 8>(mode, N, M);
    case 1440: return (M <= 240) ? run<48, 30, 8>(mode, N, M) : run<48, 30>(mode, N, M);
    case 24: return run<4, 6>(mode, N, M);
    case 36: return run<6, 6>(mode, N, M);
    default: printf("ERR size\n"); return 2;
  }
}
