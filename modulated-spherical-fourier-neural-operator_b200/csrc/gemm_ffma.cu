// Grouped fp32 GEMM on the CUDA cores (FFMA, fp32 accumulate): the exact-arithmetic ("fp32 tier",
// <= 1e-5 rel-L2) engine behind the Legendre contractions and the spectral complex MLP, and the
// only engine used by the backward passes.
//
// replaces: the cuBLAS SGEMM / CGEMM bmm calls that torch.einsum dispatches to in
//   torch_harmonics RealSHT/InverseRealSHT.forward ('...km,mlk->...lm', '...lm,mlk->...km') and
//   /root/reference MSFNO/Models/sfno/contractions.py:132-137 (einsum "bixy,io->boxy").
//
// D[M][N] = opA(A) * opB(B)^T per group; both operands may be K-major (element (r,k) at r*ld+k)
// or MN-major (element (r,k) at k*ld+r), which covers every forward/adjoint combination without
// materialising a transpose.  128x128x16 tiles, 256 threads, 8x8 register micro-tiles,
// register-staged double buffering.
#include "common.cuh"
#include "plan.h"

namespace msfno {

static constexpr int BM = 128, BN = 128, BKT = 16, LDS = BM + 4;

struct TileRegs { float4 v[2]; };

// Load a 128 x 16 operand tile into registers (zero-filled outside [R) x [K)).
template <bool KMAJOR>
__device__ __forceinline__ void load_tile(TileRegs& t, const float* __restrict__ X, long long ld, int r0, int k0,
                                          int R, int K, bool vec, int tid) {
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int f = tid + h * 256;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (KMAJOR) {
      const int row = f >> 2, kq = f & 3;
      const int gr = r0 + row, gk = k0 + kq * 4;
      if (gr < R && gk < K) {
        const float* p = X + (long long)gr * ld + gk;
        if (vec && gk + 3 < K) v = *reinterpret_cast<const float4*>(p);
        else {
          v.x = p[0];
          if (gk + 1 < K) v.y = p[1];
          if (gk + 2 < K) v.z = p[2];
          if (gk + 3 < K) v.w = p[3];
        }
      }
    } else {
      const int kk = f >> 5, rq = f & 31;
      const int gk = k0 + kk, gr = r0 + rq * 4;
      if (gk < K && gr < R) {
        const float* p = X + (long long)gk * ld + gr;
        if (vec && gr + 3 < R) v = *reinterpret_cast<const float4*>(p);
        else {
          v.x = p[0];
          if (gr + 1 < R) v.y = p[1];
          if (gr + 2 < R) v.z = p[2];
          if (gr + 3 < R) v.w = p[3];
        }
      }
    }
    t.v[h] = v;
  }
}

template <bool KMAJOR>
__device__ __forceinline__ void store_tile(const TileRegs& t, float (*S)[LDS], int tid) {
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int f = tid + h * 256;
    if (KMAJOR) {
      const int row = f >> 2, kq = f & 3;
      S[kq * 4 + 0][row] = t.v[h].x;
      S[kq * 4 + 1][row] = t.v[h].y;
      S[kq * 4 + 2][row] = t.v[h].z;
      S[kq * 4 + 3][row] = t.v[h].w;
    } else {
      const int kk = f >> 5, rq = f & 31;
      *reinterpret_cast<float4*>(&S[kk][rq * 4]) = t.v[h];
    }
  }
}

template <bool A_KMAJOR, bool B_KMAJOR>
__global__ void __launch_bounds__(256, 2) gemm_ffma_kernel(GemmLaunch g, int tilesN) {
  __shared__ __align__(16) float As[2][BKT][LDS];
  __shared__ __align__(16) float Bs[2][BKT][LDS];

  GemmGroup grp;
  if (g.use_single) {
    grp = g.single;
    grp.a_off += blockIdx.y * g.sa;
    grp.b_off += blockIdx.y * g.sb;
    grp.d_off += blockIdx.y * g.sd;
  } else {
    grp = g.groups[blockIdx.y];
  }
  const int tm = blockIdx.x / tilesN, tn = blockIdx.x - tm * tilesN;
  const int m0 = tm * BM, n0 = tn * BN;
  if (m0 >= grp.M || n0 >= grp.N) return;
  const int M = grp.M, N = grp.N, K = grp.K;
  const float* __restrict__ A = g.A + grp.a_off;
  const float* __restrict__ B = g.B + grp.b_off;
  float* __restrict__ D = g.D + grp.d_off;
  const bool vecA = (((grp.a_off | g.lda) & 3) == 0) && ((reinterpret_cast<uintptr_t>(g.A) & 15) == 0);
  const bool vecB = (((grp.b_off | g.ldb) & 3) == 0) && ((reinterpret_cast<uintptr_t>(g.B) & 15) == 0);
  const bool vecD = (((grp.d_off | g.ldd) & 3) == 0) && ((reinterpret_cast<uintptr_t>(g.D) & 15) == 0);

  const int tid = threadIdx.x;
  const int ty = tid >> 4, tx = tid & 15;

  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

  TileRegs ra, rb;
  // k-tiles of the first operand pair followed by those of the optional second pair (same majorness)
  const int nk1 = (K + BKT - 1) / BKT;
  const int K2 = g.A2 ? g.K2 : 0;
  const int nk = nk1 + (K2 + BKT - 1) / BKT;
  const float* __restrict__ A2 = g.A2 ? g.A2 + blockIdx.y * g.sa2 : nullptr;
  const float* __restrict__ B2 = g.B2 ? g.B2 + blockIdx.y * g.sb2 : nullptr;
  const bool vecA2 = A2 && (((g.sa2 | g.lda2) & 3) == 0) && ((reinterpret_cast<uintptr_t>(g.A2) & 15) == 0);
  const bool vecB2 = B2 && (((g.sb2 | g.ldb2) & 3) == 0) && ((reinterpret_cast<uintptr_t>(g.B2) & 15) == 0);
  auto load_pair = [&](int kt) {
    if (kt < nk1) {
      load_tile<A_KMAJOR>(ra, A, g.lda, m0, kt * BKT, M, K, vecA, tid);
      load_tile<B_KMAJOR>(rb, B, g.ldb, n0, kt * BKT, N, K, vecB, tid);
    } else {
      load_tile<A_KMAJOR>(ra, A2, g.lda2, m0, (kt - nk1) * BKT, M, K2, vecA2, tid);
      load_tile<B_KMAJOR>(rb, B2, g.ldb2, n0, (kt - nk1) * BKT, N, K2, vecB2, tid);
    }
  };
  if (nk > 0) {
    load_pair(0);
    store_tile<A_KMAJOR>(ra, As[0], tid);
    store_tile<B_KMAJOR>(rb, Bs[0], tid);
  }
  __syncthreads();

  for (int kt = 0; kt < nk; ++kt) {
    const int cur = kt & 1;
    if (kt + 1 < nk) load_pair(kt + 1);
#pragma unroll
    for (int k = 0; k < BKT; ++k) {
      const float4 a0 = *reinterpret_cast<const float4*>(&As[cur][k][ty * 4]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[cur][k][64 + ty * 4]);
      const float4 b0 = *reinterpret_cast<const float4*>(&Bs[cur][k][tx * 4]);
      const float4 b1 = *reinterpret_cast<const float4*>(&Bs[cur][k][64 + tx * 4]);
      const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    if (kt + 1 < nk) {
      store_tile<A_KMAJOR>(ra, As[cur ^ 1], tid);
      store_tile<B_KMAJOR>(rb, Bs[cur ^ 1], tid);
    }
    __syncthreads();
  }

  // epilogue: row-major store with optional ReLU(even cols) / ReLU-mask / accumulate
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int gm = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
    if (gm >= M) continue;
#pragma unroll
    for (int hh = 0; hh < 2; ++hh) {
      const int gn = n0 + hh * 64 + tx * 4;
      if (gn >= N) continue;
      float v[4] = {acc[i][hh * 4 + 0], acc[i][hh * 4 + 1], acc[i][hh * 4 + 2], acc[i][hh * 4 + 3]};
      if (g.bias) {
        const float bv = g.bias[blockIdx.y * g.sbias + gm];
        v[0] += bv; v[1] += bv; v[2] += bv; v[3] += bv;
      }
      if (g.act_gelu == 1) { v[0] = gelu_erf(v[0]); v[1] = gelu_erf(v[1]); v[2] = gelu_erf(v[2]); v[3] = gelu_erf(v[3]); }
      if (g.add) {
        const float* ap = g.add + blockIdx.y * g.sadd + (long long)gm * g.ldadd + gn;
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (gn + j < N) v[j] = (g.act_gelu == 2) ? v[j] * gelu_erf_grad(ap[j]) : v[j] + ap[j];   // 2: activation adjoint
      }
      if (g.relu_even) {
        v[0] = fmaxf(v[0], 0.f);
        v[2] = fmaxf(v[2], 0.f);
      }
      if (g.mask) {
        const float* mk = g.mask + grp.d_off + (long long)gm * g.ldmask + gn;
        if (!(mk[0] > 0.f)) v[0] = 0.f;
        if (gn + 2 < N && !(mk[2] > 0.f)) v[2] = 0.f;
      }
      float* d = D + (long long)gm * g.ldd + gn;
      if (vecD && gn + 3 < N) {
        float4 o = make_float4(v[0], v[1], v[2], v[3]);
        if (g.accumulate) {
          const float4 old = *reinterpret_cast<const float4*>(d);
          o.x += old.x; o.y += old.y; o.z += old.z; o.w += old.w;
        }
        *reinterpret_cast<float4*>(d) = o;
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (gn + j < N) d[j] = g.accumulate ? d[j] + v[j] : v[j];
      }
    }
  }
}

int launch_gemm_ffma(const GemmLaunch& g, cudaStream_t st) {
  if (g.ngroups <= 0 || g.maxM <= 0 || g.maxN <= 0) return MSFNO_OK;
  const int tilesM = (g.maxM + BM - 1) / BM, tilesN = (g.maxN + BN - 1) / BN;
  dim3 grid(tilesM * tilesN, g.ngroups);
  if (g.a_kmajor && g.b_kmajor) gemm_ffma_kernel<true, true><<<grid, 256, 0, st>>>(g, tilesN);
  else if (g.a_kmajor && !g.b_kmajor) gemm_ffma_kernel<true, false><<<grid, 256, 0, st>>>(g, tilesN);
  else if (!g.a_kmajor && g.b_kmajor) gemm_ffma_kernel<false, true><<<grid, 256, 0, st>>>(g, tilesN);
  else gemm_ffma_kernel<false, false><<<grid, 256, 0, st>>>(g, tilesN);
  count_launch();
  MSFNO_CUDA_OK(cudaGetLastError());
  return MSFNO_OK;
}

// Single-problem convenience wrapper: the one group travels by value in the kernel arguments.
int launch_gemm_single(const float* A, long long lda, int a_kmajor, const float* B, long long ldb, int b_kmajor,
                       float* D, long long ldd, int M, int N, int K, int relu_even, const float* mask,
                       long long ldmask, int accumulate, cudaStream_t st) {
  GemmLaunch g{};
  g.A = A; g.B = B; g.D = D;
  g.lda = lda; g.ldb = ldb; g.ldd = ldd;
  g.a_kmajor = a_kmajor; g.b_kmajor = b_kmajor;
  g.relu_even = relu_even; g.mask = mask; g.ldmask = ldmask; g.accumulate = accumulate;
  g.groups = nullptr; g.ngroups = 1; g.maxM = M; g.maxN = N;
  g.use_single = 1;
  g.single = GemmGroup{0, 0, 0, M, N, K, 0};
  return launch_gemm_ffma(g, st);
}

}  // namespace msfno
