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
//
// Layouts: activations in and out are position-major (PM: [b][p][2C], a mode's channels contiguous).
// A CTA owns NT = 128*NV consecutive modes n (reference tril order) and RT destination channels; it loops
// over the reduced channel in chunks of IC: the chunk of the source activations is gathered ROW-WISE
// (each packed row p(n) holds the chunk contiguously -> coalesced) into shared memory, transposed so that
// the inner loop reads it with conflict-free 8-byte loads.  (The first version gathered per lane from
// global memory: 32 distinct L1 lines per load instruction, 2.4 TB/s; profiles/r01_specconv_experiments.json.)
//
// One kernel serves the forward (reduce over i, weight w[d=k][s=i][n]) and the x-gradient
// (reduce over k, weight conj(w[s=k][d=i][n])).
//
// Main path (specconv_tma_kernel, needs the caller's workspace): the op is a pure stream, so it is written as
// one: a pre-pass gathers the activations into tril order ([b][s][n], n contiguous, 15 MB per sample), after
// which EVERY operand chunk is a contiguous 2 KB run that the TMA engine copies into a multi-stage shared-memory
// ring (3-D tensor-map cp.async.bulk.tensor + mbarrier, one producer thread, two requests per stage).  ~150 KB of loads stay in flight per SM without costing
// registers; 8 consumer warps multiply out of shared memory.  The grid is persistent (one CTA per SM) and the
// ring keeps running across tiles.  specconv_kernel below (register loads, shared-memory staging of the
// gathered activations) is the fallback when no workspace is given or ntril is odd.
#include "common.cuh"
#include "plan.h"
#include "tc_common.cuh"

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

// dst[b][p(n)][d] = sum_s src[b][p(n)][s] * W(d, s, n)
//   ADJ = false: W(d, s, n) = w[d][s][n]            (forward: d = k out, s = i in)
//   ADJ = true : W(d, s, n) = conj(w[s][d][n])      (x-gradient: d = i, s = k)
// grid.x: mode chunks of NT, grid.y: tiles of RT destination channels, grid.z: batch tiles of BT
template <int BT, int NV, int RT, int IC, bool ADJ>
__global__ void __launch_bounds__(SC_THREADS) specconv_kernel(const float* __restrict__ src_pm,
                                                              const float* __restrict__ w, float* __restrict__ dst_pm,
                                                              const int* __restrict__ n2p, int B, int Cs, int Cd, int Ci,
                                                              int ntril, int P) {
  constexpr int NT = SC_THREADS * NV;
  constexpr int PITCH = NT + 1;                      // float2 elements per staged row (odd: conflict-free transposed stores)
  extern __shared__ float2 a_s[];                    // [BT][IC][PITCH]
  __shared__ int p_s[NT];

  const int nbase = blockIdx.x * NT;
  const int d0 = blockIdx.y * RT;
  const int b0 = blockIdx.z * BT;
  for (int t = threadIdx.x; t < NT; t += SC_THREADS) p_s[t] = n2p[min(nbase + t, ntril - 1)];

  const int n0 = nbase + threadIdx.x * NV;           // first mode of this thread
  bool ok[NV];
#pragma unroll
  for (int v = 0; v < NV; ++v) ok[v] = (n0 + v) < ntril;

  float2 acc[RT][BT][NV];
#pragma unroll
  for (int r = 0; r < RT; ++r)
#pragma unroll
    for (int b = 0; b < BT; ++b)
#pragma unroll
      for (int v = 0; v < NV; ++v) acc[r][b][v] = make_float2(0.f, 0.f);

  const size_t wrow = (size_t)ntril * 2;             // floats per (k, i) weight row
  const int Cs2 = 2 * Cs;
  // staging: half-warps read IC consecutive complex values (IC*8 bytes) of one packed row
  constexpr int LPR = (IC >= 16) ? 16 : IC;          // lanes per row
  constexpr int RPW = 32 / LPR;                      // rows per warp-instruction
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int l_i = lane % LPR, l_r = lane / LPR;
  __syncthreads();

  for (int s0 = 0; s0 < Cs; s0 += IC) {
    // ---- stage src[b][p(n)][s0 .. s0+IC) for the CTA's NT modes: a_s[b][i][n]
#pragma unroll
    for (int b = 0; b < BT; ++b) {
      const int bb = min(b0 + b, B - 1);
      for (int row = warp * RPW + l_r; row < NT; row += (SC_THREADS / 32) * RPW) {
#pragma unroll
        for (int ii = l_i; ii < IC; ii += LPR) {
          float2 v = make_float2(0.f, 0.f);
          if (s0 + ii < Cs) v = __ldg(reinterpret_cast<const float2*>(src_pm + ((size_t)bb * P + p_s[row]) * Cs2 + 2 * (s0 + ii)));
          a_s[(b * IC + ii) * PITCH + row] = v;
        }
      }
    }
    __syncthreads();
    // ---- stream the weights of this chunk: RT x IC rows of NT modes
#pragma unroll 4
    for (int ii = 0; ii < IC; ++ii) {
      const int s = s0 + ii;
      if (s >= Cs) break;
      float2 av[BT][NV];
#pragma unroll
      for (int b = 0; b < BT; ++b)
#pragma unroll
        for (int v = 0; v < NV; ++v) av[b][v] = a_s[(b * IC + ii) * PITCH + threadIdx.x * NV + v];
#pragma unroll
      for (int r = 0; r < RT; ++r) {
        if (d0 + r >= Cd) break;
        const size_t rowidx = ADJ ? ((size_t)s * Ci + (d0 + r)) : ((size_t)(d0 + r) * Ci + s);
        const float* wp = w + rowidx * wrow + (size_t)n0 * 2;
        float2 wv[NV];
        if (NV == 2 && ok[NV - 1]) {
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
          for (int v = 0; v < NV; ++v) {
            if (ADJ) cmac_conj(acc[r][b][v], wv[v], av[b][v]);
            else cmac(acc[r][b][v], av[b][v], wv[v]);
          }
      }
    }
    __syncthreads();
  }

  // ---- store: RT complex values (8*RT bytes) contiguous per (b, mode) in the PM layout
  const int Cd2 = 2 * Cd;
#pragma unroll
  for (int b = 0; b < BT; ++b) {
    if (b0 + b >= B) break;
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      if (!ok[v]) continue;
      float* o = dst_pm + ((size_t)(b0 + b) * P + p_s[threadIdx.x * NV + v]) * Cd2 + 2 * d0;
      if (RT == 4 && d0 + 3 < Cd && ((Cd2 | (2 * d0)) & 3) == 0) {
        *reinterpret_cast<float4*>(o) = make_float4(acc[0][b][v].x, acc[0][b][v].y, acc[1][b][v].x, acc[1][b][v].y);
        *reinterpret_cast<float4*>(o + 4) = make_float4(acc[2][b][v].x, acc[2][b][v].y, acc[RT - 1][b][v].x, acc[RT - 1][b][v].y);
      } else {
#pragma unroll
        for (int r = 0; r < RT; ++r)
          if (d0 + r < Cd) *reinterpret_cast<float2*>(o + 2 * r) = acc[r][b][v];
      }
    }
  }
}


// ------------------------------------------------------------------------------------------ TMA-fed stream
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;\n" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;\n" : "=l"(pol));
  return pol;
}
// a_t[b][s][n] (complex) = src_pm[b][p(n)][s]: tile transpose through shared memory, coalesced on both sides
__global__ void __launch_bounds__(256) specconv_gather_kernel(const float* __restrict__ src_pm, float* __restrict__ a_t,
                                                              const int* __restrict__ n2p, int Cs, int ntril, int P) {
  __shared__ float2 tile[32][33];
  const int n0 = blockIdx.x * 32, s0 = blockIdx.y * 32, b = blockIdx.z;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int n = n0 + ty + 8 * j, s = s0 + tx;
    float2 v = make_float2(0.f, 0.f);
    if (n < ntril && s < Cs) v = __ldg(reinterpret_cast<const float2*>(src_pm + ((size_t)b * P + n2p[n]) * (2 * Cs) + 2 * s));
    tile[ty + 8 * j][tx] = v;
  }
  __syncthreads();
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int s = s0 + ty + 8 * j, n = n0 + tx;
    if (n < ntril && s < Cs) reinterpret_cast<float2*>(a_t)[((size_t)b * Cs + s) * ntril + n] = tile[tx][ty + 8 * j];
  }
}

static constexpr int ST_NT = 256;        // modes per tile = consumer threads
static constexpr int ST_RT = 4;          // destination channels per tile
static constexpr int ST_CONS_WARPS = ST_NT / 32;

__device__ __forceinline__ void tma_load_3d_hint(void* smem_dst, const CUtensorMap* tm, uint64_t* bar, int c0, int c1, int c2,
                                                 uint64_t pol) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%3, %4, %5}], "
      "[%2], %6;\n" ::"r"(smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(tm)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "l"(pol)
      : "memory");
}

// dst[b][p(n)][d] = sum_s a_t[b][s][n] * W(d, s, n); persistent grid, tiles = (mode chunk, d tile, batch tile).
// A stage of the ring is two TMA boxes of complex (8-byte) elements: the weights [RT][IC][NT] (forward) or
// [IC][RT][NT] (adjoint: the same tensor map geometry [Co][Ci][n] read with the channel roles swapped) and the
// activations [BT][IC][NT].  Out-of-range modes / channels / samples are zero-filled by the TMA unit, so the
// consumers run without predicates and only the final store is masked.
template <int BT, int IC, int NS, bool ADJ>
__global__ void __launch_bounds__(ST_NT + 32, 1) specconv_tma_kernel(const __grid_constant__ CUtensorMap tm_w,
                                                                     const __grid_constant__ CUtensorMap tm_a,
                                                                     float* __restrict__ dst_pm, const int* __restrict__ n2p,
                                                                     int B, int Cs, int Cd, int ntril, int P) {
  constexpr int W_ROWS = ST_RT * IC, A_ROWS = BT * IC;
  constexpr int STAGE_F2 = (W_ROWS + A_ROWS) * ST_NT;   // float2 elements per stage
  constexpr uint32_t STAGE_BYTES = STAGE_F2 * 8u;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  float2* ring = reinterpret_cast<float2*>(smem_raw);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw + sizeof(float2) * (size_t)STAGE_F2 * NS);
  uint64_t* empty = full + NS;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < NS; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], ST_CONS_WARPS); }
    fence_mbar_init();
  }
  __syncthreads();

  const int nchunks = (ntril + ST_NT - 1) / ST_NT;
  const int ndt = (Cd + ST_RT - 1) / ST_RT;
  const int nbt = (B + BT - 1) / BT;
  const int ntiles = nchunks * ndt * nbt;
  const int nsteps = (Cs + IC - 1) / IC;

  if (warp == ST_CONS_WARPS) {
    // ------------------------------------------------------------------ producer: one thread feeds the ring
    if (lane == 0) {
      const uint64_t pol_w = l2_policy_evict_first(), pol_a = l2_policy_evict_last();
      uint32_t it = 0;
      for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int nbase = (tile % nchunks) * ST_NT;
        const int d0 = ((tile / nchunks) % ndt) * ST_RT;
        const int b0 = (tile / (nchunks * ndt)) * BT;
        for (int step = 0; step < nsteps; ++step, ++it) {
          const int stage = it % NS;
          const uint32_t phase = (it / NS) & 1u;
          const int s0 = step * IC;
          mbar_wait_bounded(&empty[stage], phase ^ 1u);
          mbar_arrive_expect_tx(&full[stage], STAGE_BYTES);
          float2* ws = ring + (size_t)stage * STAGE_F2;
          // tensor w[k][i][n]: forward reads k = d, i = s; adjoint reads k = s, i = d
          if (ADJ) tma_load_3d_hint(ws, &tm_w, &full[stage], nbase, d0, s0, pol_w);
          else tma_load_3d_hint(ws, &tm_w, &full[stage], nbase, s0, d0, pol_w);
          tma_load_3d_hint(ws + W_ROWS * ST_NT, &tm_a, &full[stage], nbase, s0, b0, pol_a);
        }
      }
    }
    return;
  }

  // -------------------------------------------------------------------- consumers: thread <-> mode
  const int t = threadIdx.x;
  uint32_t it = 0;
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int nbase = (tile % nchunks) * ST_NT;
    const int d0 = ((tile / nchunks) % ndt) * ST_RT;
    const int b0 = (tile / (nchunks * ndt)) * BT;
    const int rt_n = min(ST_RT, Cd - d0), bt_n = min(BT, B - b0);
    const bool okn = (nbase + t) < ntril;
    const int p = okn ? n2p[nbase + t] : 0;

    float2 acc[ST_RT][BT];
#pragma unroll
    for (int r = 0; r < ST_RT; ++r)
#pragma unroll
      for (int b = 0; b < BT; ++b) acc[r][b] = make_float2(0.f, 0.f);

    for (int step = 0; step < nsteps; ++step, ++it) {
      const int stage = it % NS;
      const uint32_t phase = (it / NS) & 1u;
      mbar_wait_bounded(&full[stage], phase);
      const float2* ws = ring + (size_t)stage * STAGE_F2 + t;
      const float2* as = ws + W_ROWS * ST_NT;
#pragma unroll
      for (int ii = 0; ii < IC; ++ii) {
        float2 av[BT];
#pragma unroll
        for (int b = 0; b < BT; ++b) av[b] = as[(b * IC + ii) * ST_NT];
#pragma unroll
        for (int r = 0; r < ST_RT; ++r) {
          const float2 wv = ws[(ADJ ? (ii * ST_RT + r) : (r * IC + ii)) * ST_NT];
#pragma unroll
          for (int b = 0; b < BT; ++b) {
            if (ADJ) cmac_conj(acc[r][b], wv, av[b]);
            else cmac(acc[r][b], av[b], wv);
          }
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty[stage]);
    }

    if (okn) {
      const int Cd2 = 2 * Cd;
#pragma unroll
      for (int b = 0; b < BT; ++b) {
        if (b < bt_n) {
          float* o = dst_pm + ((size_t)(b0 + b) * P + p) * Cd2 + 2 * d0;
          if (rt_n == ST_RT && ((Cd2 | (2 * d0)) & 3) == 0) {
            *reinterpret_cast<float4*>(o) = make_float4(acc[0][b].x, acc[0][b].y, acc[1][b].x, acc[1][b].y);
            *reinterpret_cast<float4*>(o + 4) = make_float4(acc[2][b].x, acc[2][b].y, acc[3][b].x, acc[3][b].y);
          } else {
#pragma unroll
            for (int r = 0; r < ST_RT; ++r)
              if (r < rt_n) *reinterpret_cast<float2*>(o + 2 * r) = acc[r][b];
          }
        }
      }
    }
  }
}

// 3-D map over complex (8-byte) elements: dims (innermost first) [d0][d1][d2], dense; box [ST_NT][b1][b2]
static int make_map_c3(CUtensorMap* tm, const float* base, long long d0, long long d1, long long d2, int b1, int b2) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return record_error(MSFNO_ERR_CUDA, "cuTensorMapEncodeTiled entry point unavailable");
  cuuint64_t dims[3] = {(cuuint64_t)d0, (cuuint64_t)d1, (cuuint64_t)d2};
  cuuint64_t strides[2] = {(cuuint64_t)d0 * 8, (cuuint64_t)d0 * (cuuint64_t)d1 * 8};
  cuuint32_t box[3] = {(cuuint32_t)ST_NT, (cuuint32_t)b1, (cuuint32_t)b2};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_UINT64, 3, const_cast<float*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return record_error(MSFNO_ERR_CUDA, "cuTensorMapEncodeTiled failed (specconv)");
  return MSFNO_OK;
}

template <int BT, int IC, int NS, bool ADJ>
static int launch_specconv_tma(const msfno_plan* p, const float* a_t, const float* w, float* dst, int B, int Cs, int Cd, int Ci,
                               int Co, cudaStream_t st) {
  constexpr size_t smem = sizeof(float2) * (size_t)(ST_RT * IC + BT * IC) * ST_NT * NS + 2 * NS * sizeof(uint64_t);
  static_assert(smem <= 227 * 1024, "ring does not fit in shared memory");
  CUtensorMap tm_w, tm_a;
  int rc = make_map_c3(&tm_w, w, p->ntril, Ci, Co, ADJ ? ST_RT : IC, ADJ ? IC : ST_RT);
  if (rc) return rc;
  rc = make_map_c3(&tm_a, a_t, p->ntril, Cs, B, IC, BT);
  if (rc) return rc;
  auto kern = specconv_tma_kernel<BT, IC, NS, ADJ>;
  MSFNO_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int dev = 0, sms = 0;
  MSFNO_CUDA_OK(cudaGetDevice(&dev));
  MSFNO_CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const int ntiles = ((p->ntril + ST_NT - 1) / ST_NT) * ((Cd + ST_RT - 1) / ST_RT) * ((B + BT - 1) / BT);
  kern<<<ntiles < sms ? ntiles : sms, ST_NT + 32, smem, st>>>(tm_w, tm_a, dst, p->d_n2p, B, Cs, Cd, p->ntril, p->P);
  count_launch();
  MSFNO_CUDA_OK(cudaGetLastError());
  return MSFNO_OK;
}

// gw[k][i][n] = sum_b conj(a[b][i][n]) * g[b][k][n];  grid.y: k tiles (KT), grid.z: i tiles (IT); a, g in PM layout
template <int NV, int KT, int IT>
__global__ void __launch_bounds__(SC_THREADS) specconv_bwdw_kernel(const float* __restrict__ a_pm,
                                                                   const float* __restrict__ g_pm,
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
        gv[k][v] = __ldg(reinterpret_cast<const float2*>(g_pm + ((size_t)b * P + p[v]) * Co2 + 2 * kk));
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
      if (NV == 2 && ok[NV - 1]) {
        __stcs(reinterpret_cast<float4*>(o), make_float4(acc[k][i][0].x, acc[k][i][0].y, acc[k][i][NV - 1].x, acc[k][i][NV - 1].y));
      } else {
#pragma unroll
        for (int v = 0; v < NV; ++v)
          if (ok[v]) __stcs(reinterpret_cast<float2*>(o + 2 * v), acc[k][i][v]);
      }
    }
  }
}

// Same product on tril-ordered operands (a_t [b][i][n], g_t [b][k][n], see specconv_gather_kernel): every load and
// store is lane-contiguous.  The 3.8 GB gradient write is the floor; operand re-reads (L2) shrink with the tile.
template <int KT, int IT>
__global__ void __launch_bounds__(256) specconv_bwdw_t_kernel(const float2* __restrict__ a_t, const float2* __restrict__ g_t,
                                                              float2* __restrict__ gw, int B, int Ci, int Co, int ntril) {
  const int n = blockIdx.x * 256 + threadIdx.x;
  const int k0 = blockIdx.y * KT, i0 = blockIdx.z * IT;
  if (n >= ntril) return;
  float2 acc[KT][IT];
#pragma unroll
  for (int k = 0; k < KT; ++k)
#pragma unroll
    for (int i = 0; i < IT; ++i) acc[k][i] = make_float2(0.f, 0.f);
#pragma unroll 2
  for (int b = 0; b < B; ++b) {
    float2 av[IT], gv[KT];
#pragma unroll
    for (int i = 0; i < IT; ++i) av[i] = __ldg(a_t + ((size_t)b * Ci + min(i0 + i, Ci - 1)) * ntril + n);
#pragma unroll
    for (int k = 0; k < KT; ++k) gv[k] = __ldg(g_t + ((size_t)b * Co + min(k0 + k, Co - 1)) * ntril + n);
#pragma unroll
    for (int k = 0; k < KT; ++k)
#pragma unroll
      for (int i = 0; i < IT; ++i) cmac_conj(acc[k][i], av[i], gv[k]);
  }
#pragma unroll
  for (int k = 0; k < KT; ++k) {
    if (k0 + k >= Co) break;
#pragma unroll
    for (int i = 0; i < IT; ++i)
      if (i0 + i < Ci) __stcs(gw + ((size_t)(k0 + k) * Ci + (i0 + i)) * ntril + n, acc[k][i]);
  }
}

template <int BT, int NV, int IC, bool ADJ>
static int launch_specconv(const msfno_plan* p, const float* src, const float* w, float* dst, int B, int Cs, int Cd, int Ci,
                           cudaStream_t st) {
  constexpr int RT = 4, NT = SC_THREADS * NV;
  const size_t smem = sizeof(float2) * BT * IC * (NT + 1);
  auto kern = specconv_kernel<BT, NV, RT, IC, ADJ>;
  MSFNO_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid((p->ntril + NT - 1) / NT, (Cd + RT - 1) / RT, (B + BT - 1) / BT);
  kern<<<grid, SC_THREADS, smem, st>>>(src, w, dst, p->d_n2p, B, Cs, Cd, Ci, p->ntril, p->P);
  count_launch();
  MSFNO_CUDA_OK(cudaGetLastError());
  return MSFNO_OK;
}

template <bool ADJ>
static int dispatch_specconv(const msfno_plan* p, const float* src, const float* w, float* dst, float* ws, int B, int Cs,
                             int Cd, int Ci, int Co, cudaStream_t st) {
  // pad slots of the destination (PM layout) must hold zeros: later GEMMs multiply them by zero table entries
  MSFNO_CUDA_OK(cudaMemsetAsync(dst, 0, sizeof(float) * (size_t)B * 2 * Cd * p->P, st));
  const bool aligned = (p->ntril % 2 == 0) && ((reinterpret_cast<uintptr_t>(w) & 15) == 0) &&
                       ((reinterpret_cast<uintptr_t>(ws) & 15) == 0);
  if (ws && aligned) {
    dim3 g((p->ntril + 31) / 32, (Cs + 31) / 32, B);
    specconv_gather_kernel<<<g, 256, 0, st>>>(src, ws, p->d_n2p, Cs, p->ntril, p->P);
    count_launch();
    MSFNO_CUDA_OK(cudaGetLastError());
    if (B == 1) return launch_specconv_tma<1, 4, 5, ADJ>(p, ws, w, dst, B, Cs, Cd, Ci, Co, st);
    if (B == 2) return launch_specconv_tma<2, 4, 4, ADJ>(p, ws, w, dst, B, Cs, Cd, Ci, Co, st);
    if (B <= 4) return launch_specconv_tma<4, 4, 3, ADJ>(p, ws, w, dst, B, Cs, Cd, Ci, Co, st);
    return launch_specconv_tma<8, 2, 4, ADJ>(p, ws, w, dst, B, Cs, Cd, Ci, Co, st);
  }
  const bool vec = (p->ntril % 2 == 0) && ((reinterpret_cast<uintptr_t>(w) & 15) == 0);
  if (B == 1 && vec) return launch_specconv<1, 2, 16, ADJ>(p, src, w, dst, B, Cs, Cd, Ci, st);
  if (B <= 2 && vec) return launch_specconv<2, 2, 8, ADJ>(p, src, w, dst, B, Cs, Cd, Ci, st);
  if (B <= 4) return launch_specconv<4, 1, 16, ADJ>(p, src, w, dst, B, Cs, Cd, Ci, st);
  return launch_specconv<8, 1, 8, ADJ>(p, src, w, dst, B, Cs, Cd, Ci, st);
}

}  // namespace msfno

using namespace msfno;

extern "C" {

size_t msfno_specconv_ws_floats(const msfno_plan* p, int B, int Ci, int Co) {
  if (!p || B < 1 || Ci < 1 || Co < 1) return 0;
  return (size_t)B * (size_t)(Ci + Co) * (size_t)p->ntril * 2;   // bwd_w stages both operands
}

int msfno_specconv_fwd(const msfno_plan* p, const float* a_pm, const float* w, float* out_pm, float* ws, int B, int Ci, int Co,
                       void* stream) {
  if (!p || !a_pm || !w || !out_pm || B < 1 || Ci < 1 || Co < 1) return record_error(MSFNO_ERR_BAD_SHAPE, "specconv_fwd: bad argument");
  return dispatch_specconv<false>(p, a_pm, w, out_pm, ws, B, Ci, Co, Ci, Co, (cudaStream_t)stream);
}

int msfno_specconv_bwd_x(const msfno_plan* p, const float* g_pm, const float* w, float* ga_pm, float* ws, int B, int Ci,
                         int Co, void* stream) {
  if (!p || !g_pm || !w || !ga_pm || B < 1 || Ci < 1 || Co < 1) return record_error(MSFNO_ERR_BAD_SHAPE, "specconv_bwd_x: bad argument");
  return dispatch_specconv<true>(p, g_pm, w, ga_pm, ws, B, Co, Ci, Ci, Co, (cudaStream_t)stream);
}

int msfno_specconv_bwd_w(const msfno_plan* p, const float* a_pm, const float* g_pm, float* gw, float* ws, int B, int Ci,
                         int Co, void* stream) {
  if (!p || !a_pm || !g_pm || !gw || B < 1 || Ci < 1 || Co < 1) return record_error(MSFNO_ERR_BAD_SHAPE, "specconv_bwd_w: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  if (ws && (reinterpret_cast<uintptr_t>(ws) & 7) == 0 && (reinterpret_cast<uintptr_t>(gw) & 7) == 0) {
    float* a_t = ws;
    float* g_t = ws + (size_t)B * Ci * p->ntril * 2;
    specconv_gather_kernel<<<dim3((p->ntril + 31) / 32, (Ci + 31) / 32, B), 256, 0, st>>>(a_pm, a_t, p->d_n2p, Ci, p->ntril, p->P);
    specconv_gather_kernel<<<dim3((p->ntril + 31) / 32, (Co + 31) / 32, B), 256, 0, st>>>(g_pm, g_t, p->d_n2p, Co, p->ntril, p->P);
    constexpr int KT2 = 8, IT2 = 4;
    dim3 grid((p->ntril + 255) / 256, (Co + KT2 - 1) / KT2, (Ci + IT2 - 1) / IT2);
    specconv_bwdw_t_kernel<KT2, IT2><<<grid, 256, 0, st>>>(reinterpret_cast<const float2*>(a_t), reinterpret_cast<const float2*>(g_t),
                                                          reinterpret_cast<float2*>(gw), B, Ci, Co, p->ntril);
    count_launch(3);
    MSFNO_CUDA_OK(cudaGetLastError());
    return MSFNO_OK;
  }
  const bool vec = (p->ntril % 2 == 0) && ((reinterpret_cast<uintptr_t>(gw) & 15) == 0);
  constexpr int KT = 4, IT = 4;
  if (vec) {
    dim3 grid((p->ntril + SC_THREADS * 2 - 1) / (SC_THREADS * 2), (Co + KT - 1) / KT, (Ci + IT - 1) / IT);
    specconv_bwdw_kernel<2, KT, IT><<<grid, SC_THREADS, 0, st>>>(a_pm, g_pm, gw, p->d_n2p, B, Ci, Co, p->ntril, p->P);
  } else {
    dim3 grid((p->ntril + SC_THREADS - 1) / SC_THREADS, (Co + KT - 1) / KT, (Ci + IT - 1) / IT);
    specconv_bwdw_kernel<1, KT, IT><<<grid, SC_THREADS, 0, st>>>(a_pm, g_pm, gw, p->d_n2p, B, Ci, Co, p->ntril, p->P);
  }
  count_launch();
  MSFNO_CUDA_OK(cudaGetLastError());
  return MSFNO_OK;
}

}  // extern "C"
