// Fully unrolled in-register DFTs of small composite length (factors 2,3,4,5) with compile-time twiddles.
// They are the building block of the four-step longitude FFT (fft2d.cu): a length-H transform is split as
// H = P1 x P2, every lane runs whole length-P1 (then length-P2) DFTs in registers, and shared memory is touched
// once per step instead of once per radix stage.
//
// Part of the replacement of torch.fft.rfft / irfft inside torch_harmonics RealSHT / InverseRealSHT
// (call sites /root/reference MSFNO/Models/sfno/layers.py:405,421,629,638).
#pragma once
#include <type_traits>

#include "fft_core.cuh"

namespace msfno {

// ---- constexpr cos / sin of 2*pi*x (double, Taylor after octant reduction) ----------------------------
constexpr double cx_pi = 3.14159265358979323846264338327950288;
constexpr double cx_poly_sin(double x) {  // |x| <= pi/4
  double x2 = x * x, term = x, sum = x;
  for (int i = 1; i < 12; ++i) {
    term *= -x2 / ((2 * i) * (2 * i + 1));
    sum += term;
  }
  return sum;
}
constexpr double cx_poly_cos(double x) {
  double x2 = x * x, term = 1.0, sum = 1.0;
  for (int i = 1; i < 12; ++i) {
    term *= -x2 / ((2 * i - 1) * (2 * i));
    sum += term;
  }
  return sum;
}
// cos(2 pi num/den), sin(2 pi num/den) evaluated exactly on the rational argument (no large-angle error)
constexpr double cx_cos_turn(int num, int den) {
  num %= den;
  if (num < 0) num += den;
  // reduce to [0, 1/8] turn using symmetries; work in eighths: e = 8*num/den
  // angle a = 2 pi num/den
  if (8 * num <= den) return cx_poly_cos(2.0 * cx_pi * num / den);
  if (8 * num <= 3 * den) return cx_poly_sin(2.0 * cx_pi * (den - 4 * num) / (4.0 * den));      // cos a = sin(pi/2 - a)
  if (8 * num <= 5 * den) return -cx_poly_cos(2.0 * cx_pi * (2 * num - den) / (2.0 * den));     // cos a = -cos(a - pi)
  if (8 * num <= 7 * den) return cx_poly_sin(2.0 * cx_pi * (4 * num - 3 * den) / (4.0 * den));  // cos a = sin(a - 3pi/2)
  return cx_poly_cos(2.0 * cx_pi * (num - den) / den);
}
constexpr double cx_sin_turn(int num, int den) { return cx_cos_turn(4 * num - den, 4 * den); }  // sin a = cos(a - pi/2)

template <int NUM, int DEN>
struct TwC {
  static constexpr float c = (float)cx_cos_turn(NUM, DEN);
  static constexpr float s = (float)cx_sin_turn(NUM, DEN);
};

template <int I, int N, typename F>
MSFNO_HD void static_for(F&& f) {
  if constexpr (I < N) {
    f(std::integral_constant<int, I>{});
    static_for<I + 1, N>(f);
  }
}

// y[k] = sum_n v[n] exp(SGN 2 pi i n k / N), in place, natural order in and out
template <int N, int SGN>
struct RegDft {
  static constexpr int P = (N % 4 == 0) ? 4 : (N % 2 == 0) ? 2 : (N % 3 == 0) ? 3 : 5;
  static constexpr int Q = N / P;
  static_assert(N == P * Q && (N % 5 == 0 || N % 3 == 0 || N % 2 == 0), "length must factor into 2, 3, 5");

  static MSFNO_HD void run(cf* v) {
    if constexpr (Q == 1) {
      Butterfly<P, SGN>::run(v);
    } else {
      cf t[N];
      // P-point DFTs of the Q stride-Q subsequences, twiddled by w_N^(q kp): t[kp*Q + q]
      static_for<0, Q>([&](auto qc) {
        constexpr int q = decltype(qc)::value;
        cf a[P];
        static_for<0, P>([&](auto pc) {
          constexpr int p = decltype(pc)::value;
          a[p] = v[q + Q * p];
        });
        Butterfly<P, SGN>::run(a);
        static_for<0, P>([&](auto kc) {
          constexpr int kp = decltype(kc)::value;
          constexpr int e = (q * kp) % N;
          if constexpr (e == 0) {
            t[kp * Q + q] = a[kp];
          } else {
            const cf w{TwC<e, N>::c, (SGN > 0 ? 1.0f : -1.0f) * TwC<e, N>::s};
            t[kp * Q + q] = cmul(a[kp], w);
          }
        });
      });
      // Q-point DFT over q for every kp: X[kp + P kq]
      static_for<0, P>([&](auto kc) {
        constexpr int kp = decltype(kc)::value;
        RegDft<Q, SGN>::run(t + kp * Q);
        static_for<0, Q>([&](auto qc) {
          constexpr int kq = decltype(qc)::value;
          v[kp + P * kq] = t[kp * Q + kq];
        });
      });
    }
  }
};

template <int SGN> struct RegDft<1, SGN> {
  static MSFNO_HD void run(cf*) {}
};

}  // namespace msfno
