// Host emulation of the device FFT core (csrc/fft_core.cuh): the 32 lanes of a warp are executed in a
// loop; stages are separated exactly where the kernel has __syncwarp().  Reads a problem from stdin,
// prints the result; tests/test_host_emul.py compares with numpy.fft.
#include <cstdio>
#include <cmath>
#include <vector>
#define __host__
#define __device__
#define __forceinline__ inline
#define __restrict__
#include "../../modulated-spherical-fourier-neural-operator_b200/csrc/fft_core_host.h"
using namespace msfno;

template <int SGN>
static int run_fft(std::vector<cf>& a, std::vector<cf>& b, const std::vector<cf>& tw, int H, const FftSchedule& s) {
  // emulate lane loop per stage (stockham_fft itself loops stages; call it per lane is wrong since sync
  // separates stages) -> re-implement the stage loop here with the same stage function.
  int Ns = 1, cur = 0;
  for (int st = 0; st < s.nstages; ++st) {
    const cf* src = cur ? b.data() : a.data();
    cf* dst = cur ? a.data() : b.data();
    for (int lane = 0; lane < 32; ++lane) {
      switch (s.radix[st]) {
        case 2: stockham_stage<2, SGN>(src, dst, tw.data(), H, Ns, lane, 32); break;
        case 3: stockham_stage<3, SGN>(src, dst, tw.data(), H, Ns, lane, 32); break;
        case 4: stockham_stage<4, SGN>(src, dst, tw.data(), H, Ns, lane, 32); break;
        default: stockham_stage<5, SGN>(src, dst, tw.data(), H, Ns, lane, 32); break;
      }
    }
    Ns *= s.radix[st];
    cur ^= 1;
  }
  return cur;
}

int main() {
  int mode, N, M;
  if (scanf("%d %d %d", &mode, &N, &M) != 3) return 1;
  const int H = N / 2;
  FftSchedule s;
  if (!make_schedule(H, &s)) { printf("ERR schedule\n"); return 2; }
  std::vector<cf> tw(H), tw2(M + 1), a(H), b(H);
  for (int t = 0; t < H; ++t) { double ang = -2.0 * M_PI * t / H; tw[t] = cf{(float)cos(ang), (float)sin(ang)}; }
  for (int m = 0; m <= M; ++m) { double ang = -2.0 * M_PI * m / N; tw2[m] = cf{(float)cos(ang), (float)sin(ang)}; }
  if (mode == 0) {  // forward: N reals in, M complex out
    for (int i = 0; i < H; ++i) { if (scanf("%f %f", &a[i].x, &a[i].y) != 2) return 1; }
    int res = run_fft<-1>(a, b, tw, H, s);
    const cf* Z = res ? b.data() : a.data();
    for (int m = 0; m < M; ++m) { cf X = r2c_split(Z, tw2.data(), H, m); printf("%.9g %.9g\n", X.x, X.y); }
  } else {  // inverse: M complex in, N reals out (kernel logic for building Zt)
    std::vector<cf> X(M);
    for (int i = 0; i < M; ++i) { if (scanf("%f %f", &X[i].x, &X[i].y) != 2) return 1; }
    const int mlim = M;
    auto Xh = [&](int q) -> cf { if (q >= mlim) return cf{0, 0}; cf v = X[q]; if (q == 0 || q == H) v.y = 0; return v; };
    for (int k = 0; k < H; ++k) {
      const int kk = H - k;
      cf out{0, 0};
      if (k < mlim || kk < mlim) {
        cf w;
        if (k <= mlim) w = tw2[k]; else { w = tw2[kk]; w.x = -w.x; }
        out = c2r_merge(Xh(k), Xh(kk), w);
      }
      a[k] = out;
    }
    int res = run_fft<+1>(a, b, tw, H, s);
    const cf* Z = res ? b.data() : a.data();
    for (int i = 0; i < H; ++i) printf("%.9g\n%.9g\n", Z[i].x, Z[i].y);
  }
  return 0;
}
