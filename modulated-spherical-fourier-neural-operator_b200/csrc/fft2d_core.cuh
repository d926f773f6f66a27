// Four-step length-H complex FFT for one row, H = P1 x P2, as two rounds of in-register DFTs (fft_reg.cuh).
//   n = P2*n1 + n2  (input index),   k = k1 + P1*k2  (output index)
//   step 1 (task n2 < P2): Y[k1][n2] = w_H^(n2 k1) * sum_n1 x[P2 n1 + n2] w_P1^(n1 k1)
//   step 2 (task k1 < P1): X[k1 + P1 k2] = sum_n2 Y[k1][n2] w_P2^(n2 k2)
// The intermediate Y lives in shared memory with an odd row pitch (P2 | 1) so that both the column writes of
// step 1 and the row reads of step 2 are bank-conflict free for 8-byte accesses.
// The functions are __host__ __device__: tests/host_emul runs them on the CPU, one "lane task" at a time.
#pragma once
#include "fft_reg.cuh"

namespace msfno {

template <int P2> struct WorkPitch { static constexpr int value = P2 | 1; };

// tw[t] = exp(-2 pi i t / H), t < H.
// Rows n1 in [ZLO, ZHI) of the input matrix are known to be zero (truncated spectrum of the inverse transform): they
// are neither stored by the caller nor loaded here.
template <int P1, int P2, int SGN, int ZLO = 0, int ZHI = 0>
MSFNO_HD void fft2d_step1(const cf* __restrict__ in, int in_n1_stride, cf* __restrict__ work, const cf* __restrict__ tw,
                          int n2) {
  constexpr int WP = WorkPitch<P2>::value;
  cf v[P1];
  static_for<0, P1>([&](auto c) {
    constexpr int n1 = decltype(c)::value;
    if constexpr (n1 >= ZLO && n1 < ZHI) v[n1] = cf{0.0f, 0.0f};
    else v[n1] = in[n1 * in_n1_stride + n2];
  });
  RegDft<P1, SGN>::run(v);
  static_for<0, P1>([&](auto c) {
    constexpr int k1 = decltype(c)::value;
    cf y = v[k1];
    if constexpr (k1 > 0) {
      cf w = tw[n2 * k1];  // n2*k1 < P2*P1 = H
      if (SGN > 0) w.y = -w.y;
      y = cmul(y, w);
    }
    work[k1 * WP + n2] = y;
  });
}

// Loads row k1 of Y, transforms it, returns X[k1 + P1*k2] in v[k2].
template <int P1, int P2, int SGN>
MSFNO_HD void fft2d_step2(const cf* __restrict__ work, int k1, cf* v) {
  constexpr int WP = WorkPitch<P2>::value;
  static_for<0, P2>([&](auto c) {
    constexpr int n2 = decltype(c)::value;
    v[n2] = work[k1 * WP + n2];
  });
  RegDft<P2, SGN>::run(v);
}

// Index of output bin k inside the compact store that keeps only the bins the real split needs:
// [0, mlim) and [H - mlim, H).  When 2*mlim >= H everything is kept.
MSFNO_HD int xs_index(int k, int H, int mlim) {
  if (2 * mlim >= H) return k;
  if (k < mlim) return k;
  if (k >= H - mlim) return k - (H - 2 * mlim);
  return -1;
}
MSFNO_HD int xs_size(int H, int mlim) { return (2 * mlim >= H) ? H : 2 * mlim; }

// Forward split on the compact store (see r2c_split in fft_core.cuh)
MSFNO_HD cf r2c_split_xs(const cf* Xs, const cf* tw2, int H, int mlim, int m) {
  const int i0 = m % H;
  int i1 = (H - m) % H;
  if (i1 < 0) i1 += H;
  cf a = Xs[xs_index(i0, H, mlim)];
  cf b = Xs[xs_index(i1, H, mlim)];
  b.y = -b.y;
  cf e = cf{0.5f * (a.x + b.x), 0.5f * (a.y + b.y)};
  cf d = csub(a, b);
  cf o = cf{0.5f * d.y, -0.5f * d.x};
  return cadd(e, cmul(tw2[m], o));
}

}  // namespace msfno
