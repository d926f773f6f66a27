// SpectralConvS2 per-mode complex channel contraction and its two adjoints.
//
// replaces: compl_contract_fwd_c = einsum("bin,kin->bkn") on complex64
//   (/root/reference MSFNO/Models/sfno/contractions.py:37-41, called from layers.py:411) together
//   with the tril gather/scatter around it (layers.py:408-413).  cuBLAS handles that einsum by first
//   permuting the 3.8 GB weight to mode-major on EVERY call; here the weight is streamed exactly once
//   in its native [k][i][n][2] parameter layout (n contiguous -> fully coalesced 16-byte loads).
//
// The op is a batch of n = 7260 tiny (B x Ci) x (Ci x Co) complex products: HBM-bound on the weight
// for any realistic B, so it runs on the CUDA cores and is tuned for bytes in flight.
#include "common.cuh"
#include "plan.h"

namespace msfno {

static constexpr int SC_THREADS = 128;

__device__ __forceinline__ void cmac(float2& acc, const float2 a, const float2 w) {
  acc.x = fmaf(a.x, w.x, acc.x);
  acc.x = fmaf(-a.y, w.y, acc.x);
  acc.y = fmaf(a.x, w.y, acc.y);
  acc.y = fmaf(a.y, w.x, acc.y);
}
// acc += conj(w) * g
__device__ __forceinline__ void cmac_conj(float2& acc, const float2 w, const float2 g) {
  acc.x = fmaf(w.x, g.x, acc.x);
  acc.x = fmaf(w.y, g.y, acc.x);
  acc.y = fmaf(w.x, g.y, acc.y);
  acc.y = fmaf(-w.y, g.x, acc.y);
}

// out[b][k][n] = sum_i a[b][i][n] * w[k][i][n]
//   grid.x: chunks of SC_THREADS*NV modes, grid.y: tiles of KT output channels, grid.z: batch tiles of BT
template <int BT, int NV, int KT>
__global__ void __launch_bounds__(SC_THREADS) specconv_fwd_kernel(const float* __restrict__ a_pm,
                                                                  const float* __restrict__ w,
                                                                  float* __restrict__ out_cm,
                                                                  const int* __restrict__ n2p, int B, int Ci, int Co,
                                                                  int ntril, int P) {
  const int n0 = (blockIdx.x * SC_THREADS + threadIdx.x) * NV;
  const int k0 = blockIdx.y * KT;
  const int b0 = blockIdx.z * BT;
  if (n0 >= ntril) return;
  int p[NV];
  bool ok[NV];
#pragma unroll
  for (int v = 0; v < NV; ++v) {
    ok[v] = (n0 + v) < ntril;
    p[v] = n2p[ok[v] ? n0 + v : n0];
  }
  float2 acc[KT][BT][NV];
#pragma unroll
  for (int k = 0; k < KT; ++k)
#pragma unroll
    for (int b = 0; b < BT; ++b)
#pragma unroll
      for (int v = 0; v < NV; ++v) acc[k][b][v] = make_float2(0.f, 0.f);

  const size_t wstride_i = (size_t)ntril * 2;
  const size_t wstride_k = (size_t)Ci * wstride_i;
  const int C2 = 2 * Ci;
#pragma unroll 4
  for (int i = 0; i < Ci; ++i) {
    float2 av[BT][NV];
#pragma unroll
    for (int b = 0; b < BT; ++b)
#pragma unroll
      for (int v = 0; v < NV; ++v) {
        const int bb = (b0 + b < B) ? b0 + b : B - 1;
        av[b][v] = __ldg(reinterpret_cast<const float2*>(a_pm + ((size_t)bb * P + p[v]) * C2 + 2 * i));
      }
#pragma unroll
    for (int k = 0; k < KT; ++k) {
      if (k0 + k >= Co) break;
      const float* wp = w + (size_t)(k0 + k) * wstride_k + (size_t)i * wstride_i + (size_t)n0 * 2;
      float2 wv[NV];
      if (NV == 2 && ok[1]) {
        const float4 t = __ldcs(reinterpret_cast<const float4*>(wp));
        wv[0] = make_float2(t.x, t.y);
        wv[NV - 1] = make_float2(t.z, t.w);
      } else {
#pragma unroll
        for (int v = 0; v < NV; ++v)
          wv[v] = ok[v] ? __ldcs(reinterpret_cast<const float2*>(wp + 2 * v)) : make_float2(0.f, 0.f);
      }
#pragma unroll
      for (int b = 0; b < BT; ++b)
#pragma unroll
        for (int v = 0; v < NV; ++v) cmac(acc[k][b][v], av[b][v], wv[v]);
    }
  }
  const int Co2 = 2 * Co;
#pragma unroll
  for (int k = 0; k < KT; ++k) {
    if (k0 + k >= Co) break;
#pragma unroll
    for (int b = 0; b < BT; ++b) {
      if (b0 + b >= B) break;
#pragma unroll
      for (int v = 0; v < NV; ++v) {
        if (!ok[v]) continue;
        float* o = out_cm + ((size_t)(b0 + b) * Co2 + 2 * (k0 + k)) * P + p[v];
        o[0] = acc[k][b][v].x;
        o[P] = acc[k][b][v].y;
      }
    }
  }
}

// ga[b][i][n] = sum_k conj(w[k][i][n]) * g[b][k][n];  grid.y: tiles of IT input channels
template <int BT, int NV, int IT>
__global__ void __launch_bounds__(SC_THREADS) specconv_bwdx_kernel(const float* __restrict__ g_cm,
                                                                   const float* __restrict__ w,
                                                                   float* __restrict__ ga_pm,
                                                                   const int* __restrict__ n2p, int B, int Ci, int Co,
                                                                   int ntril, int P) {
  const int n0 = (blockIdx.x * SC_THREADS + threadIdx.x) * NV;
  const int i0 = blockIdx.y * IT;
  const int b0 = blockIdx.z * BT;
  if (n0 >= ntril) return;
  int p[NV];
  bool ok[NV];
#pragma unroll
  for (int v = 0; v < NV; ++v) {
    ok[v] = (n0 + v) < ntril;
    p[v] = n2p[ok[v] ? n0 + v : n0];
  }
  float2 acc[IT][BT][NV];
#pragma unroll
  for (int i = 0; i < IT; ++i)
#pragma unroll
    for (int b = 0; b < BT; ++b)
#pragma unroll
      for (int v = 0; v < NV; ++v) acc[i][b][v] = make_float2(0.f, 0.f);
  const size_t wstride_i = (size_t)ntril * 2;
  const size_t wstride_k = (size_t)Ci * wstride_i;
  const int Co2 = 2 * Co;
#pragma unroll 4
  for (int k = 0; k < Co; ++k) {
    float2 gv[BT][NV];
#pragma unroll
    for (int b = 0; b < BT; ++b)
#pragma unroll
      for (int v = 0; v < NV; ++v) {
        const int bb = (b0 + b < B) ? b0 + b : B - 1;
        const float* gp = g_cm + ((size_t)bb * Co2 + 2 * k) * P + p[v];
        gv[b][v] = make_float2(__ldg(gp), __ldg(gp + P));
      }
#pragma unroll
    for (int i = 0; i < IT; ++i) {
      if (i0 + i >= Ci) break;
      const float* wp = w + (size_t)k * wstride_k + (size_t)(i0 + i) * wstride_i + (size_t)n0 * 2;
      float2 wv[NV];
      if (NV == 2 && ok[1]) {
        const float4 t = __ldcs(reinterpret_cast<const float4*>(wp));
        wv[0] = make_float2(t.x, t.y);
        wv[NV - 1] = make_float2(t.z, t.w);
      } else {
#pragma unroll
        for (int v = 0; v < NV; ++v)
          wv[v] = ok[v] ? __ldcs(reinterpret_cast<const float2*>(wp + 2 * v)) : make_float2(0.f, 0.f);
      }
#pragma unroll
      for (int b = 0; b < BT; ++b)
#pragma unroll
        for (int v = 0; v < NV; ++v) cmac_conj(acc[i][b][v], wv[v], gv[b][v]);
    }
  }
  const int C2 = 2 * Ci;
#pragma unroll
  for (int b = 0; b < BT; ++b) {
    if (b0 + b >= B) break;
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      if (!ok[v]) continue;
      float* o = ga_pm + ((size_t)(b0 + b) * P + p[v]) * C2 + 2 * i0;
#pragma unroll
      for (int i = 0; i < IT; ++i) {
        if (i0 + i >= Ci) break;
        *reinterpret_cast<float2*>(o + 2 * i) = acc[i][b][v];
      }
    }
  }
}

// gw[k][i][n] = sum_b conj(a[b][i][n]) * g[b][k][n];  grid.y: k tiles (KT), grid.z: i tiles (IT)
template <int NV, int KT, int IT>
__global__ void __launch_bounds__(SC_THREADS) specconv_bwdw_kernel(const float* __restrict__ a_pm,
                                                                   const float* __restrict__ g_cm,
                                                                   float* __restrict__ gw,
                                                                   const int* __restrict__ n2p, int B, int Ci, int Co,
                                                                   int ntril, int P) {
  const int n0 = (blockIdx.x * SC_THREADS + threadIdx.x) * NV;
  const int k0 = blockIdx.y * KT;
  const int i0 = blockIdx.z * IT;
  if (n0 >= ntril) return;
  int p[NV];
  bool ok[NV];
#pragma unroll
  for (int v = 0; v < NV; ++v) {
    ok[v] = (n0 + v) < ntril;
    p[v] = n2p[ok[v] ? n0 + v : n0];
  }
  float2 acc[KT][IT][NV];
#pragma unroll
  for (int k = 0; k < KT; ++k)
#pragma unroll
    for (int i = 0; i < IT; ++i)
#pragma unroll
      for (int v = 0; v < NV; ++v) acc[k][i][v] = make_float2(0.f, 0.f);
  const int C2 = 2 * Ci, Co2 = 2 * Co;
  for (int b = 0; b < B; ++b) {
    float2 av[IT][NV], gv[KT][NV];
#pragma unroll
    for (int v = 0; v < NV; ++v) {
#pragma unroll
      for (int i = 0; i < IT; ++i) {
        const int ii = (i0 + i < Ci) ? i0 + i : Ci - 1;
        av[i][v] = __ldg(reinterpret_cast<const float2*>(a_pm + ((size_t)b * P + p[v]) * C2 + 2 * ii));
      }
#pragma unroll
      for (int k = 0; k < KT; ++k) {
        const int kk = (k0 + k < Co) ? k0 + k : Co - 1;
        const float* gp = g_cm + ((size_t)b * Co2 + 2 * kk) * P + p[v];
        gv[k][v] = make_float2(__ldg(gp), __ldg(gp + P));
      }
    }
#pragma unroll
    for (int k = 0; k < KT; ++k)
#pragma unroll
      for (int i = 0; i < IT; ++i)
#pragma unroll
        for (int v = 0; v < NV; ++v) cmac_conj(acc[k][i][v], av[i][v], gv[k][v]);
  }
  const size_t wstride_i = (size_t)ntril * 2;
  const size_t wstride_k = (size_t)Ci * wstride_i;
#pragma unroll
  for (int k = 0; k < KT; ++k) {
    if (k0 + k >= Co) break;
#pragma unroll
    for (int i = 0; i < IT; ++i) {
      if (i0 + i >= Ci) break;
      float* o = gw + (size_t)(k0 + k) * wstride_k + (size_t)(i0 + i) * wstride_i + (size_t)n0 * 2;
      if (NV == 2 && ok[1]) {
        __stcs(reinterpret_cast<float4*>(o), make_float4(acc[k][i][0].x, acc[k][i][0].y, acc[k][i][NV - 1].x, acc[k][i][NV - 1].y));
      } else {
#pragma unroll
        for (int v = 0; v < NV; ++v)
          if (ok[v]) __stcs(reinterpret_cast<float2*>(o + 2 * v), acc[k][i][v]);
      }
    }
  }
}

}  // namespace msfno

using namespace msfno;

extern "C" {

int msfno_specconv_fwd(const msfno_plan* p, const float* a_pm, const float* w, float* out_cm, int B, int Ci, int Co,
                       void* stream) {
  if (!p || !a_pm || !w || !out_cm || B < 1 || Ci < 1 || Co < 1) return record_error(MSFNO_ERR_BAD_SHAPE, "specconv_fwd: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  // pad slots of the CM layout must hold zeros (the synthesis GEMM multiplies them by zero table entries)
  MSFNO_CUDA_OK(cudaMemsetAsync(out_cm, 0, sizeof(float) * (size_t)B * 2 * Co * p->P, st));
  const bool vec = (p->ntril % 2 == 0) && ((reinterpret_cast<uintptr_t>(w) & 15) == 0);
  constexpr int KT = 4;
#define LAUNCH_FWD(BT, NV)                                                                                  \
  {                                                                                                         \
    dim3 grid((p->ntril + SC_THREADS * NV - 1) / (SC_THREADS * NV), (Co + KT - 1) / KT, (B + BT - 1) / BT); \
    specconv_fwd_kernel<BT, NV, KT><<<grid, SC_THREADS, 0, st>>>(a_pm, w, out_cm, p->d_n2p, B, Ci, Co, p->ntril, p->P); \
  }
  if (B == 1 && vec) LAUNCH_FWD(1, 2)
  else if (B <= 2 && vec) LAUNCH_FWD(2, 2)
  else if (B <= 4) LAUNCH_FWD(4, 1)
  else LAUNCH_FWD(8, 1)
#undef LAUNCH_FWD
  count_launch();
  MSFNO_CUDA_OK(cudaGetLastError());
  return MSFNO_OK;
}

int msfno_specconv_bwd_x(const msfno_plan* p, const float* g_cm, const float* w, float* ga_pm, int B, int Ci, int Co,
                         void* stream) {
  if (!p || !g_cm || !w || !ga_pm || B < 1 || Ci < 1 || Co < 1) return record_error(MSFNO_ERR_BAD_SHAPE, "specconv_bwd_x: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  MSFNO_CUDA_OK(cudaMemsetAsync(ga_pm, 0, sizeof(float) * (size_t)B * 2 * Ci * p->P, st));
  const bool vec = (p->ntril % 2 == 0) && ((reinterpret_cast<uintptr_t>(w) & 15) == 0);
  constexpr int IT = 4;
#define LAUNCH_BX(BT, NV)                                                                                   \
  {                                                                                                         \
    dim3 grid((p->ntril + SC_THREADS * NV - 1) / (SC_THREADS * NV), (Ci + IT - 1) / IT, (B + BT - 1) / BT); \
    specconv_bwdx_kernel<BT, NV, IT><<<grid, SC_THREADS, 0, st>>>(g_cm, w, ga_pm, p->d_n2p, B, Ci, Co, p->ntril, p->P); \
  }
  if (B == 1 && vec) LAUNCH_BX(1, 2)
  else if (B <= 2 && vec) LAUNCH_BX(2, 2)
  else if (B <= 4) LAUNCH_BX(4, 1)
  else LAUNCH_BX(8, 1)
#undef LAUNCH_BX
  count_launch();
  MSFNO_CUDA_OK(cudaGetLastError());
  return MSFNO_OK;
}

int msfno_specconv_bwd_w(const msfno_plan* p, const float* a_pm, const float* g_cm, float* gw, int B, int Ci, int Co,
                         void* stream) {
  if (!p || !a_pm || !g_cm || !gw || B < 1 || Ci < 1 || Co < 1) return record_error(MSFNO_ERR_BAD_SHAPE, "specconv_bwd_w: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  const bool vec = (p->ntril % 2 == 0) && ((reinterpret_cast<uintptr_t>(gw) & 15) == 0);
  constexpr int KT = 4, IT = 4;
  if (vec) {
    dim3 grid((p->ntril + SC_THREADS * 2 - 1) / (SC_THREADS * 2), (Co + KT - 1) / KT, (Ci + IT - 1) / IT);
    specconv_bwdw_kernel<2, KT, IT><<<grid, SC_THREADS, 0, st>>>(a_pm, g_cm, gw, p->d_n2p, B, Ci, Co, p->ntril, p->P);
  } else {
    dim3 grid((p->ntril + SC_THREADS - 1) / SC_THREADS, (Co + KT - 1) / KT, (Ci + IT - 1) / IT);
    specconv_bwdw_kernel<1, KT, IT><<<grid, SC_THREADS, 0, st>>>(a_pm, g_cm, gw, p->d_n2p, B, Ci, Co, p->ntril, p->P);
  }
  count_launch();
  MSFNO_CUDA_OK(cudaGetLastError());
  return MSFNO_OK;
}

}  // extern "C"
