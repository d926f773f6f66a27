// Mixed-radix Stockham FFT building blocks shared by the forward (r2c, truncated) and
// inverse (c2r, zero-padded) longitude transforms.
//
// Replaces the cuFFT calls behind torch.fft.rfft / torch.fft.irfft in torch_harmonics'
// RealSHT.forward / InverseRealSHT.forward (call sites: /root/reference
// MSFNO/Models/sfno/layers.py:405,421,629,638; op order in SURVEY.md Appendix A.3).
//
// One warp owns one length-H complex FFT (H = nlon/2) held in shared memory; every stage
// is "each lane does butterflies j = lane, lane+32, ..." separated by __syncwarp().  The
// functions are __host__ __device__ so tests/host_emul can run the identical code on the
// CPU with the 32 lanes executed in a loop.
#pragma once
#include <cuda_runtime.h>

#if defined(__CUDACC__)
#define MSFNO_HD __host__ __device__ __forceinline__
#else
#define MSFNO_HD inline
#endif

namespace msfno {

struct cf { float x, y; };

MSFNO_HD cf cmul(cf a, cf b) { return cf{a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x}; }
MSFNO_HD cf cadd(cf a, cf b) { return cf{a.x + b.x, a.y + b.y}; }
MSFNO_HD cf csub(cf a, cf b) { return cf{a.x - b.x, a.y - b.y}; }
// multiply by (sgn * i)
template <int SGN> MSFNO_HD cf muli(cf a) { return SGN > 0 ? cf{-a.y, a.x} : cf{a.y, -a.x}; }

// In-register DFT of R points: y[p] = sum_q v[q] * exp(SGN * 2 pi i p q / R)
template <int R, int SGN> struct Butterfly;

template <int SGN> struct Butterfly<2, SGN> {
  static MSFNO_HD void run(cf* v) {
    cf a = v[0], b = v[1];
    v[0] = cadd(a, b);
    v[1] = csub(a, b);
  }
};

template <int SGN> struct Butterfly<3, SGN> {
  static MSFNO_HD void run(cf* v) {
    const float s60 = 0.86602540378443864676f;
    cf t1 = cadd(v[1], v[2]);
    cf t2 = cf{v[0].x - 0.5f * t1.x, v[0].y - 0.5f * t1.y};
    cf d = csub(v[1], v[2]);
    cf t3 = muli<SGN>(cf{s60 * d.x, s60 * d.y});
    v[0] = cadd(v[0], t1);
    v[1] = cadd(t2, t3);
    v[2] = csub(t2, t3);
  }
};

template <int SGN> struct Butterfly<4, SGN> {
  static MSFNO_HD void run(cf* v) {
    cf a = cadd(v[0], v[2]), b = csub(v[0], v[2]);
    cf c = cadd(v[1], v[3]), d = muli<SGN>(csub(v[1], v[3]));
    v[0] = cadd(a, c);
    v[1] = cadd(b, d);
    v[2] = csub(a, c);
    v[3] = csub(b, d);
  }
};

template <int SGN> struct Butterfly<5, SGN> {
  static MSFNO_HD void run(cf* v) {
    const float c1 = 0.30901699437494742410f, c2 = -0.80901699437494742410f;
    const float s1 = 0.95105651629515357212f, s2 = 0.58778525229247312917f;
    cf t1 = cadd(v[1], v[4]), t2 = cadd(v[2], v[3]);
    cf t3 = csub(v[1], v[4]), t4 = csub(v[2], v[3]);
    cf a1 = cf{v[0].x + c1 * t1.x + c2 * t2.x, v[0].y + c1 * t1.y + c2 * t2.y};
    cf a2 = cf{v[0].x + c2 * t1.x + c1 * t2.x, v[0].y + c2 * t1.y + c1 * t2.y};
    cf b1 = muli<SGN>(cf{s1 * t3.x + s2 * t4.x, s1 * t3.y + s2 * t4.y});
    cf b2 = muli<SGN>(cf{s2 * t3.x - s1 * t4.x, s2 * t3.y - s1 * t4.y});
    v[0] = cf{v[0].x + t1.x + t2.x, v[0].y + t1.y + t2.y};
    v[1] = cadd(a1, b1);
    v[4] = csub(a1, b1);
    v[2] = cadd(a2, b2);
    v[3] = csub(a2, b2);
  }
};

// One Stockham stage of radix R over a length-H sequence: src -> dst.
//   Ns  = product of the radices of the previous stages
//   tw  = table tw[t] = exp(-2 pi i t / H), t < H  (forward sign; conjugated when SGN > 0)
// Lane `lane` of `nlanes` handles butterflies j = lane, lane + nlanes, ...
template <int R, int SGN>
MSFNO_HD void stockham_stage(const cf* __restrict__ src, cf* __restrict__ dst, const cf* __restrict__ tw,
                             int H, int Ns, int lane, int nlanes) {
  const int T = H / R;
  const int tws = H / (Ns * R);  // twiddle stride
  for (int j = lane; j < T; j += nlanes) {
    const int k = j % Ns;
    cf v[R];
#pragma unroll
    for (int q = 0; q < R; ++q) v[q] = src[j + q * T];
    if (Ns > 1) {
#pragma unroll
      for (int q = 1; q < R; ++q) {
        cf w = tw[q * k * tws];
        if (SGN > 0) w.y = -w.y;
        v[q] = cmul(v[q], w);
      }
    }
    Butterfly<R, SGN>::run(v);
    const int j0 = (j / Ns) * Ns * R + k;
#pragma unroll
    for (int q = 0; q < R; ++q) dst[j0 + q * Ns] = v[q];
  }
}

// Radix schedule: factors of H drawn from {4,5,3,2}, at most 12 stages.
struct FftSchedule {
  int nstages;
  int radix[12];
};

inline bool make_schedule(int H, FftSchedule* s) {
  s->nstages = 0;
  const int cand[4] = {4, 5, 3, 2};
  for (int c = 0; c < 4; ++c)
    while (H % cand[c] == 0 && H > 1) {
      if (s->nstages >= 12) return false;
      s->radix[s->nstages++] = cand[c];
      H /= cand[c];
    }
  return H == 1;
}

struct WarpSync {
  MSFNO_HD void operator()() const {
#if defined(__CUDA_ARCH__)
    __syncwarp();
#endif
  }
};

// Full length-H complex FFT for one warp.  Data starts in `a`; stages ping-pong a -> b -> a ...
// Returns the buffer index (0 = a, 1 = b) holding the result.  `sync` is __syncwarp on device.
template <int SGN, typename Sync>
MSFNO_HD int stockham_fft(cf* a, cf* b, const cf* tw, int H, const FftSchedule& s, int lane, int nlanes, Sync sync) {
  int Ns = 1;
  int cur = 0;
  for (int st = 0; st < s.nstages; ++st) {
    const cf* src = cur ? b : a;
    cf* dst = cur ? a : b;
    switch (s.radix[st]) {
      case 2: stockham_stage<2, SGN>(src, dst, tw, H, Ns, lane, nlanes); break;
      case 3: stockham_stage<3, SGN>(src, dst, tw, H, Ns, lane, nlanes); break;
      case 4: stockham_stage<4, SGN>(src, dst, tw, H, Ns, lane, nlanes); break;
      default: stockham_stage<5, SGN>(src, dst, tw, H, Ns, lane, nlanes); break;
    }
    Ns *= s.radix[st];
    cur ^= 1;
    sync();
  }
  return cur;
}

// Forward split: from Z = FFT_H(x[2n] + i x[2n+1]) produce X[m] = sum_j x[j] exp(-2 pi i m j / N), N = 2H.
//   tw2[m] = exp(-2 pi i m / N)
MSFNO_HD cf r2c_split(const cf* Z, const cf* tw2, int H, int m) {
  int i0 = m % H;  // Z[H] == Z[0]
  int i1 = (H - m) % H;
  if (i1 < 0) i1 += H;
  cf a = Z[i0];
  cf b = Z[i1];
  b.y = -b.y;
  cf e = cf{0.5f * (a.x + b.x), 0.5f * (a.y + b.y)};
  cf d = csub(a, b);
  cf o = cf{0.5f * d.y, -0.5f * d.x};  // -i/2 * (a - b)
  return cadd(e, cmul(tw2[m], o));
}

// Inverse merge: Zt[k] = (Xh[k] + conj(Xh[H-k])) + i * exp(+2 pi i k / N) * (Xh[k] - conj(Xh[H-k])),
// where Xh is the Hermitian half-spectrum (bins >= M are zero).  The caller supplies a and b=Xh[H-k].
MSFNO_HD cf c2r_merge(cf a, cf bh, cf tw2k) {
  cf b = cf{bh.x, -bh.y};
  cf s = cadd(a, b);
  cf d = csub(a, b);
  cf w = cf{tw2k.x, -tw2k.y};  // exp(+2 pi i k / N)
  cf t = cmul(w, d);
  return cf{s.x - t.y, s.y + t.x};  // s + i t
}

}  // namespace msfno
