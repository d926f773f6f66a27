// Longitude transforms of the tensor-core tier as TF32 GEMMs against precomputed DFT matrices.
//
// replaces (tensor-core tier, forward direction only): the rfft / irfft stage of RealSHT / InverseRealSHT
//   (torch_harmonics: x = 2 pi rfft(x, norm="forward") truncated to mmax; irfft(x, n=nlon, norm="forward")), i.e.
//   launch_rfft2d / launch_irfft2d of the fp32 tier.  The truncated spectrum keeps only mlim <= 120 orders, so the
//   transform is a [rows x nlon] x [nlon x 2 mlim] product: 0.5-1 kflop per input byte -- tensor-core work.  On this
//   part the FP32 pipe limits the FFT kernels (52 / 60 us for a 30 MB inner grid, 0.55 / 0.74 ms at 721x1440); the
//   same contraction on tcgen05 is bound by memory again.
//
//   forward : Xt[b][m][2c+ri][lat] = tf32( sc[b,c] * sum_j x[b,c,lat,j] F[2m+ri][j] + [m=0,ri=0] 2 pi sh[b,c] )
//             A = x rows (K-major), B = F [2 mlim][nlon] (K-major); lat sits on the TMEM lanes, so the transposed
//             store into the [m][2C][kpad] layout of the Legendre stage is a coalesced 128-byte access per column
//   inverse : y[b,c,lat,j] = act( sum_n Yt[b][m][2c+ri][lat] G[j][n] + skip ), plane statistics on the way out
//             A = Yt read in place as an MN-major operand through a 5-D tensor map (lat contiguous, n = (m, ri) rows),
//             B = G [nlon][2 mlim] (K-major); the output tile is staged through shared memory (the pipeline stages are
//             idle by then) so that every global access of y / skip is a contiguous row segment
//
// CTA tile: MT x 128 rows (lat of one (b, c) plane) by 256 columns, K blocks of 32; MT = 2 halves the re-reads of the
// DFT matrix at 721x1440 (two accumulators = all 512 TMEM columns).  Warp roles as in gemm_tc.cu.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "plan.h"
#include "tc_common.cuh"

namespace msfno {

static constexpr int DF_BN = 256;
static constexpr int DF_A_BYTES = 128 * TC_BK * 4;      // 16 KB: 128 rows x 32 k
static constexpr int DF_B_BYTES = DF_BN * TC_BK * 4;    // 32 KB: 256 rows x 32 k

struct DftParams {
  float* out;
  const float* sc; const float* sh;     // forward: per-plane scale / shift (InstanceNorm affine) or null
  const float* skip;                    // inverse: tensor added before the activation, or null
  double* stats;                        // inverse: per-plane (sum, sum of squares) accumulators or null
  int nlat, nlon, mlim, kpad, C;
  int nkb;                              // k blocks
  int flags;                            // forward: bit0 round to TF32; inverse: bit0 GELU, bit1 round to TF32
  float dc;                             // forward: factor of the shift on the (m = 0, re) bin = 2 pi
  long long* trace;                     // debug (MSFNO_DFT_TRACE): per-CTA clock stamps, or null
  int skip_tma;                         // inverse: skip is accumulated on the tensor cores (tmS / tmI), the epilogue does not load it
  int out_tma;                          // inverse: output leaves through per-warp TMA tensor stores (tmO)
};

// TMA tensor store shared -> global of one 3-D box (bulk async group of the issuing thread)
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* tm, const void* smem_src, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];\n" ::"l"(
                   reinterpret_cast<uint64_t>(tm)),
               "r"(smem_u32(smem_src)), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;\n" ::"n"(N) : "memory"); }

__device__ __forceinline__ void tma_load_5d(void* smem_dst, const CUtensorMap* tm, uint64_t* bar, int c0, int c1, int c2, int c3,
                                            int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];\n" ::"r"(
          smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(tm)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}

#define MSFNO_DFT_LD32(r, taddr)                                                                                         \
  asm volatile(                                                                                                          \
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "                                                                          \
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "                                          \
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"                        \
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),      \
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),           \
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),          \
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])                        \
      : "r"(taddr))

// grid: (N tiles of 256 [inverse only], row tiles of MT*128, B*C planes)
template <bool INV, int MT, int NS>
__global__ void __launch_bounds__(256, 1)
dft_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const __grid_constant__ CUtensorMap tmS,
              const __grid_constant__ CUtensorMap tmI, const __grid_constant__ CUtensorMap tmO, DftParams p) {
  constexpr int STAGE = MT * DF_A_BYTES + DF_B_BYTES;
  extern __shared__ uint8_t smem_raw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int bc = blockIdx.z, b = bc / p.C, c = bc - b * p.C;
  const int lat0 = blockIdx.y * (MT * 128);
  const int n0 = blockIdx.x * DF_BN;

  pdl_trigger();
  const long long t_start = clock64();
  const int cta_lin = (blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* tiles = smem_raw + (base - smem_u32(smem_raw));
  uint8_t* idblk = tiles + NS * STAGE;                               // 32 x 32 identity, K-major swizzled (skip path)
  uint64_t* bars = reinterpret_cast<uint64_t*>(idblk + 4096);
  uint64_t* full = bars;
  uint64_t* empty = bars + NS;
  uint64_t* tmem_full = bars + 2 * NS;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * NS + 1);
  uint64_t* id_full = bars + 2 * NS + 2;
  double* red = reinterpret_cast<double*>(bars + 2 * NS + 4);   // [16]

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < NS; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    mbar_init(tmem_full, 1);
    mbar_init(id_full, 1);
    fence_mbar_init();
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(tmem_slot)), "r"(MT * DF_BN));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();
  if (p.trace && threadIdx.x == 0 && cta_lin < 512) p.trace[cta_lin * 8 + 0] = clock64() - t_start;

  if (warp == 0 && lane == 0) {
    // ---------------- TMA producer ----------------
    for (int kb = 0; kb < p.nkb; ++kb) {
      const int s = kb % NS;
      mbar_wait_bounded(&empty[s], (uint32_t)(((kb / NS) & 1) ^ 1));
      mbar_arrive_expect_tx(&full[s], STAGE);
      uint8_t* sa = tiles + (size_t)s * STAGE;
      if (!INV) {
        // A: rows (plane, lat) of x, 32 longitudes per block
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) tma_load_2d(sa + mt * DF_A_BYTES, &tmA, &full[s], kb * TC_BK, bc * p.nlat + lat0 + mt * 128);
      } else {
        // A: Yt[b][m][2c+ri][lat]: box = 32 lat x (ri 2) x (c 1) x (m 16) x (b 1) = 32 k-rows of 128 bytes
#pragma unroll
        for (int mt = 0; mt < MT; ++mt)
#pragma unroll
          for (int j = 0; j < 4; ++j)
            tma_load_5d(sa + mt * DF_A_BYTES + j * (TC_BK * 128), &tmA, &full[s], lat0 + mt * 128 + 32 * j, 0, c, kb * (TC_BK / 2), b);
      }
      tma_load_2d(sa + MT * DF_A_BYTES, &tmB, &full[s], kb * TC_BK, n0);
    }
    if (INV && p.skip_tma) {
      // skip tile: [MT x 128 rows (lat)] x 32 longitudes per block, K-major, through the A slots of the same ring; it is
      // accumulated by N = 32 MMAs against the identity (plain loads of it were the inverse kernel's top stall)
      mbar_arrive_expect_tx(id_full, 4096);
      tma_load_2d(idblk, &tmI, id_full, 0, 0);
      // a stage holds SPS skip blocks here (the B part of the stage is free in this phase): more bytes in flight
      constexpr int SPS = STAGE / (MT * DF_A_BYTES);
      const int nsk = (min(DF_BN, p.nlon - n0) + TC_BK - 1) / TC_BK;
      for (int j0 = 0, u = 0; j0 < nsk; j0 += SPS, ++u) {
        const int kc = p.nkb + u, s = kc % NS;
        const int nb = min(SPS, nsk - j0);
        mbar_wait_bounded(&empty[s], (uint32_t)(((kc / NS) & 1) ^ 1));
        mbar_arrive_expect_tx(&full[s], (uint32_t)(nb * MT * DF_A_BYTES));
        uint8_t* sa = tiles + (size_t)s * STAGE;
        for (int i = 0; i < nb; ++i)
#pragma unroll
          for (int mt = 0; mt < MT; ++mt)
            tma_load_2d(sa + (i * MT + mt) * DF_A_BYTES, &tmS, &full[s], n0 + (j0 + i) * TC_BK, bc * p.nlat + lat0 + mt * 128);
      }
    }
  } else if (warp == 1 && lane == 0) {
    // ---------------- MMA issuer ----------------
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((INV ? 1u : 0u) << 15) | ((uint32_t)(DF_BN >> 3) << 17) |
                           ((uint32_t)(128 >> 4) << 24);
    for (int kb = 0; kb < p.nkb; ++kb) {
      const int s = kb % NS;
      mbar_wait_bounded(&full[s], (uint32_t)((kb / NS) & 1));
      tc_fence_after();
      const uint32_t sa = base + (uint32_t)s * STAGE;
      const uint32_t sb = sa + MT * DF_A_BYTES;
#pragma unroll
      for (int mt = 0; mt < MT; ++mt)
#pragma unroll
        for (int k = 0; k < TC_BK / 8; ++k) {
          const uint64_t adesc = INV ? make_smem_desc(sa + mt * DF_A_BYTES + 1024 * k, TC_BK * 128, 512, 1)
                                     : make_smem_desc(sa + mt * DF_A_BYTES + 32 * k, 16, 1024);
          tc_mma_tf32(tmem_base + mt * DF_BN, adesc, make_smem_desc(sb + 32 * k, 16, 1024), idesc, (kb | k) ? 1u : 0u);
        }
      tc_commit(&empty[s]);
    }
    if (INV && p.skip_tma) {
      const uint32_t idesc_s = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(32 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
      const uint32_t id_addr = base + NS * STAGE;
      mbar_wait_bounded(id_full, 0);
      constexpr int SPS = STAGE / (MT * DF_A_BYTES);
      const int nsk = (min(DF_BN, p.nlon - n0) + TC_BK - 1) / TC_BK;
      for (int j0 = 0, u = 0; j0 < nsk; j0 += SPS, ++u) {
        const int kc = p.nkb + u, s = kc % NS;
        const int nb = min(SPS, nsk - j0);
        mbar_wait_bounded(&full[s], (uint32_t)((kc / NS) & 1));
        tc_fence_after();
        const uint32_t sa = base + (uint32_t)s * STAGE;
        for (int i = 0; i < nb; ++i)
#pragma unroll
          for (int mt = 0; mt < MT; ++mt)
#pragma unroll
            for (int k = 0; k < TC_BK / 8; ++k)
              tc_mma_tf32(tmem_base + mt * DF_BN + (j0 + i) * TC_BK, make_smem_desc(sa + (i * MT + mt) * DF_A_BYTES + 32 * k, 16, 1024),
                          make_smem_desc(id_addr + 32 * k, 16, 1024), idesc_s, 1u);
        tc_commit(&empty[s]);
      }
    }
    tc_commit(tmem_full);
  }
  __syncwarp();

  // ---------------- epilogue: all 8 warps; lane quarter q, column half ----------------
  const int q = warp & 3, chalf = warp >> 2;
  mbar_wait_bounded(tmem_full, 0);
  tc_fence_after();
  if (p.trace && threadIdx.x == 0 && cta_lin < 512) p.trace[cta_lin * 8 + 1] = clock64() - t_start;
  if (!INV) {
    const float sc = p.sc ? p.sc[bc] : 1.0f;
    const float dcv = p.sh ? p.dc * p.sh[bc] : 0.0f;
    const int twoC = 2 * p.C;
#pragma unroll 1
    for (int mt = 0; mt < MT; ++mt) {
      const int lat = lat0 + mt * 128 + q * 32 + lane;
      const bool in_rows = lat < p.nlat, in_pad = lat < p.kpad;
#pragma unroll 1
      for (int c0 = chalf * 128; c0 < chalf * 128 + 128; c0 += 32) {
        if (c0 >= 2 * p.mlim) break;
        uint32_t r[32];
        MSFNO_DFT_LD32(r, tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(mt * DF_BN + c0));
        asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
        // values of this lane's latitude for 32 spectral columns -> private [col][lat] tile (pitch 36) -> four latitudes
        // (16 bytes) per lane: a store instruction then writes 128 contiguous bytes of FOUR columns instead of one
        // (the epilogue cost 25 of 95 kclk per CTA with one 4-byte store per lane and column)
        float* wt = reinterpret_cast<float*>(tiles) + warp * (32 * 36);
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          float v = __uint_as_float(r[j]) * sc;
          if (c0 + j == 0) v += dcv;
          if (p.flags & 1) v = round_to_tf32(v);
          if (!in_rows) v = 0.0f;
          wt[j * 36 + lane] = v;
        }
        __syncwarp();
        {
          const int cc = lane >> 3, l4 = lane & 7;
          const int latq = lat0 + mt * 128 + q * 32 + 4 * l4;          // first of this lane's four latitudes
          if (latq < p.kpad) {                                         // kpad % 4 == 0: all four or none
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const int j = cc + 4 * i, n = c0 + j;                    // column n = 2m + ri -> row (m * 2C + 2c + ri)
              if (n < 2 * p.mlim) {
                const float4 t4 = *reinterpret_cast<const float4*>(wt + j * 36 + 4 * l4);
                *reinterpret_cast<float4*>(p.out + (((size_t)b * p.mlim + (n >> 1)) * twoC + 2 * c + (n & 1)) * p.kpad + latq) = t4;
              }
            }
          }
        }
        __syncwarp();
      }
    }
    tc_fence_before();
  } else if (p.out_tma) {
    // Output through per-warp TMA tensor stores.  A warp owns 32 latitudes (its TMEM lane quarter) and four 32-column
    // blocks of each accumulator: TMEM -> registers (activation, rounding, plane statistics: sums do not care which lane
    // holds which element) -> a private 4 KB tile in the 128-byte-swizzle layout (conflict-free 16-byte writes, row =
    // lane) -> one 32 x 32 box store by lane 0.  Two tiles per warp alternate, so the next block is read and processed
    // while the TMA engine drains the previous one; no CTA-wide barrier and no second pass over shared memory (the
    // staged row-walk epilogue below spent 15 kclk per 128 x 128 block, most of this kernel's time).  The tensor map
    // [B C][nlat][nlon] clips the ragged last latitude tile and the ragged last column block.
    float lsum = 0.0f, lsq = 0.0f;
    uint8_t* wbuf = tiles + warp * 8192;                  // pipeline stages are idle: every MMA has completed
    int it = 0;
#pragma unroll 1
    for (int mt = 0; mt < MT; ++mt) {
      const int row0 = lat0 + mt * 128 + q * 32;
      if (row0 >= p.nlat) break;
      const bool row_ok = row0 + lane < p.nlat;
#pragma unroll 1
      for (int c0 = chalf * 128; c0 < chalf * 128 + 128; c0 += 32) {
        if (n0 + c0 >= p.nlon) break;
        uint32_t r[32];
        MSFNO_DFT_LD32(r, tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(mt * DF_BN + c0));
        asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
        if (it >= 2) {                                    // the store issued two blocks ago has read this tile
          if (lane == 0) bulk_wait_read<1>();
          __syncwarp();
        }
        uint8_t* buf = wbuf + (it & 1) * 4096;
        float v[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          float t = __uint_as_float(r[j]);
          if (p.flags & 1) t = gelu_tanh3(t);
          if (p.flags & 2) t = round_to_tf32(t);
          v[j] = t;
          const float ts = row_ok ? t : 0.0f;             // columns beyond nlon are exact zeros (zero-filled G rows)
          lsum += ts;
          lsq = fmaf(ts, ts, lsq);
        }
#pragma unroll
        for (int jj = 0; jj < 8; ++jj)
          *reinterpret_cast<float4*>(buf + lane * 128 + ((jj ^ (lane & 7)) << 4)) = make_float4(v[4 * jj], v[4 * jj + 1], v[4 * jj + 2], v[4 * jj + 3]);
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) {
          tma_store_3d(&tmO, buf, n0 + c0, row0, bc);
          bulk_commit();
        }
        ++it;
      }
    }
    if (lane == 0) bulk_wait_read<0>();                   // shared memory must outlive the engine's reads
    if (p.trace && threadIdx.x == 0 && cta_lin < 512) p.trace[cta_lin * 8 + 4] = clock64() - t_start;
    tc_fence_before();
    if (p.stats) {
      double ds = (double)lsum, dq = (double)lsq;
      for (int o = 16; o > 0; o >>= 1) {
        ds += __shfl_xor_sync(0xffffffffu, ds, o);
        dq += __shfl_xor_sync(0xffffffffu, dq, o);
      }
      if (lane == 0) { red[2 * warp] = ds; red[2 * warp + 1] = dq; }
      __syncthreads();
      if (threadIdx.x == 0) {
        double s = 0.0, sq = 0.0;
        for (int w = 0; w < 8; ++w) { s += red[2 * w]; sq += red[2 * w + 1]; }
        atomicAdd(&p.stats[2 * bc], s);
        atomicAdd(&p.stats[2 * bc + 1], sq);
      }
    }
  } else {
    // Staging tile in the (now idle) pipeline stages: all 256 columns when they fit, otherwise two passes of 128
    // (small grids run with a 2-stage ring so that two CTAs share an SM and overlap each other's phases).
    constexpr int SW = (NS * STAGE >= 128 * (DF_BN + 4) * 4) ? DF_BN : DF_BN / 2;
    constexpr int SP = SW + 4;   // pitch = 4 (mod 32) words: 16-byte writes by row (lane = row) and 16-byte reads along a row are both conflict-free
    static_assert(NS * STAGE >= 128 * SP * 4, "staging tile does not fit in the pipeline stages");
    float* stage_t = reinterpret_cast<float*>(tiles);     // [128][SP]
    float lsum = 0.0f, lsq = 0.0f;
    const size_t plane = (size_t)bc * p.nlat;
    const int ncol = min(DF_BN, p.nlon - n0);             // multiple of 4 (nlon % 4 == 0)
    const bool has_skip = p.skip != nullptr && !p.skip_tma;
#pragma unroll 1
    for (int pass = 0; pass < MT * (DF_BN / SW); ++pass) {
      const int mt = pass / (DF_BN / SW), hc = pass % (DF_BN / SW);
      const int cbase = hc * SW;                          // first column of this pass
      if (pass) __syncthreads();
      // TMEM -> staging tile
#pragma unroll 1
      for (int c0 = chalf * 128; c0 < chalf * 128 + 128; c0 += 32) {
        if (c0 < cbase || c0 >= cbase + SW) continue;
        uint32_t r[32];
        MSFNO_DFT_LD32(r, tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(mt * DF_BN + c0));
        asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
        float4* trow = reinterpret_cast<float4*>(stage_t + (size_t)(q * 32 + lane) * SP + (c0 - cbase));
#pragma unroll
        for (int j = 0; j < 8; ++j)
          trow[j] = make_float4(__uint_as_float(r[4 * j]), __uint_as_float(r[4 * j + 1]), __uint_as_float(r[4 * j + 2]), __uint_as_float(r[4 * j + 3]));
      }
      __syncthreads();
      if (p.trace && threadIdx.x == 0 && cta_lin < 512 && pass == 0) p.trace[cta_lin * 8 + 3] = clock64() - t_start;
      // staging tile -> global: a warp walks rows; a lane owns float4 groups (512 contiguous bytes per warp access).
      // The skip values are requested three rows ahead (register ring; the loop is unrolled by 3 so its indices are
      // compile-time): a dependent DRAM / L2 round trip per row was most of this phase.
      auto load_skip = [&](int row, float4 (&dst)[SW / 128]) {
        const int lat = lat0 + mt * 128 + row;
        const size_t goff = (plane + lat) * p.nlon + n0 + cbase;
#pragma unroll
        for (int i = 0; i < SW / 128; ++i) {
          const int col = 4 * (lane + 32 * i);
          dst[i] = (has_skip && row < 128 && lat < p.nlat && cbase + col < ncol) ? __ldg(reinterpret_cast<const float4*>(p.skip + goff + col))
                                                                                : make_float4(0.f, 0.f, 0.f, 0.f);
        }
      };
      float4 skr[3][SW / 128];
      load_skip(warp, skr[0]);
      load_skip(warp + 8, skr[1]);
      load_skip(warp + 16, skr[2]);
#pragma unroll 3
      for (int row = warp, u = 0; row < 128; row += 8, ++u) {
        const int lat = lat0 + mt * 128 + row;
        if (lat >= p.nlat) break;
        const size_t goff = (plane + lat) * p.nlon + n0 + cbase;
        const float* trow = stage_t + (size_t)row * SP;
        float4 sk[SW / 128];
#pragma unroll
        for (int i = 0; i < SW / 128; ++i) sk[i] = skr[u % 3][i];
        load_skip(row + 24, skr[u % 3]);
#pragma unroll
        for (int i = 0; i < SW / 128; ++i) {
          const int col = 4 * (lane + 32 * i);
          if (cbase + col < ncol) {
            const float4 t4 = *reinterpret_cast<const float4*>(trow + col);
            float v[4] = {t4.x + sk[i].x, t4.y + sk[i].y, t4.z + sk[i].z, t4.w + sk[i].w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              if (p.flags & 1) v[e] = gelu_tanh3(v[e]);
              if (p.flags & 2) v[e] = round_to_tf32(v[e]);
              lsum += v[e];
              lsq = fmaf(v[e], v[e], lsq);
            }
            *reinterpret_cast<float4*>(p.out + goff + col) = make_float4(v[0], v[1], v[2], v[3]);
          }
        }
      }
      if (p.trace && threadIdx.x == 0 && cta_lin < 512 && pass == 0) p.trace[cta_lin * 8 + 4] = clock64() - t_start;
    }
    tc_fence_before();
    if (p.stats) {
      double ds = (double)lsum, dq = (double)lsq;
      for (int o = 16; o > 0; o >>= 1) {
        ds += __shfl_xor_sync(0xffffffffu, ds, o);
        dq += __shfl_xor_sync(0xffffffffu, dq, o);
      }
      if (lane == 0) { red[2 * warp] = ds; red[2 * warp + 1] = dq; }
      __syncthreads();
      if (threadIdx.x == 0) {
        double s = 0.0, sq = 0.0;
        for (int w = 0; w < 8; ++w) { s += red[2 * w]; sq += red[2 * w + 1]; }
        atomicAdd(&p.stats[2 * bc], s);
        atomicAdd(&p.stats[2 * bc + 1], sq);
      }
    }
  }
  __syncthreads();
  if (p.trace && threadIdx.x == 0 && cta_lin < 512) p.trace[cta_lin * 8 + 2] = clock64() - t_start;
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem_base), "r"(MT * DF_BN));
  }
}


// ------------------------------------------------------------------------------------------ full-grid inverse, m-parity split
// y[lat, j]            = E[lat, j] + O[lat, j]
// y[lat, j + nlon / 2] = E[lat, j] - O[lat, j],   0 <= j < nlon / 2,
// E / O = the contributions of the even / odd orders m (cos and sin of 2 pi m (j + nlon/2) / nlon pick up (-1)^m).  Half
// the MMAs and half the DFT-matrix bytes of the plain product.  The plain kernel at 721 x 1440 moved 384 KB of operands
// through L2 per 128 KB of output and ran at the L2 throughput (3.5 GB + 1.06 GB at ~12 TB/s); here a CTA keeps the
// A operand (128 latitudes x all orders of one plane, 128 KB) resident in shared memory for ALL column tiles and streams
// only the two half matrices: 1.2 bytes of operand per byte of output.
//
// Persistent CTAs (one per SM) walk (plane, latitude tile) items.  Warp 0: TMA producer (A blocks once per item, B ring),
// warp 1: MMA issuer, warps 2-5: epilogue.  Two accumulator pairs (E, O: 2 x 128 columns each, all 512 TMEM columns)
// alternate between column tiles, so the epilogue of tile t (TMEM -> registers -> E+O / E-O -> swizzled 4 KB tile -> TMA
// tensor store) runs under the MMAs of tile t+1.
static constexpr int EO_BN = 128;                 // j columns per tile (two output blocks of 128 columns)
static constexpr int EO_KB = 4;                   // k blocks of 32 per parity (K padded to 128)
static constexpr int EO_BLK = 128 * TC_BK * 4;    // 16 KB: operand block, 128 rows x 32 k
static constexpr int EO_NSB = 4;                  // B ring slots
static constexpr int EO_EPI_WARPS = 8;
static constexpr int EO_THREADS = 64 + 32 * EO_EPI_WARPS;
static constexpr int EO_SMEM = 2 * EO_KB * EO_BLK + EO_NSB * EO_BLK + EO_EPI_WARPS * 4096 + 1024 + 512;

struct EoParams {
  double* stats;
  int nlat, half, C;          // half = nlon / 2
  int ntr, ntn, nitems;       // latitude tiles per plane, column tiles, planes * ntr
  int flags;                  // bit0 GELU, bit1 round to TF32
  long long* trace;           // debug (MSFNO_DFT_TRACE): per-CTA wait-time totals of each role, or null
};

// clock64-bracketed wait: adds the cycles spent to `acc` when tracing
#define EO_TIMED(acc, stmt)                          \
  do {                                               \
    if (p.trace) {                                   \
      const long long _t0 = clock64();               \
      stmt;                                          \
      acc += clock64() - _t0;                        \
    } else {                                         \
      stmt;                                          \
    }                                                \
  } while (0)

template <int FLAGS>
__global__ void __launch_bounds__(EO_THREADS, 1)
idft_eo_kernel(const __grid_constant__ CUtensorMap tmAe, const __grid_constant__ CUtensorMap tmAo, const __grid_constant__ CUtensorMap tmBe,
               const __grid_constant__ CUtensorMap tmBo, const __grid_constant__ CUtensorMap tmO1, const __grid_constant__ CUtensorMap tmO2,
               EoParams p) {
  extern __shared__ uint8_t smem_raw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  pdl_trigger();
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* tiles = smem_raw + (base - smem_u32(smem_raw));
  uint8_t* a_blk = tiles;                                         // [2 EO_KB] blocks: even k blocks, then odd
  uint8_t* b_ring = tiles + 2 * EO_KB * EO_BLK;
  uint8_t* stage = b_ring + EO_NSB * EO_BLK;                      // per epilogue warp: one 4 KB swizzled tile
  uint64_t* bars = reinterpret_cast<uint64_t*>(stage + EO_EPI_WARPS * 4096);
  uint64_t* a_full = bars;                  // [8]
  uint64_t* a_empty = bars + 8;             // [8]
  uint64_t* b_full = bars + 16;             // [EO_NSB]
  uint64_t* b_empty = bars + 16 + EO_NSB;   // [EO_NSB]
  uint64_t* t_full = bars + 16 + 2 * EO_NSB;       // [2]
  uint64_t* t_empty = bars + 18 + 2 * EO_NSB;      // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 20 + 2 * EO_NSB);

  if (warp == 0 && lane == 0) {
    for (int i = 0; i < 2 * EO_KB; ++i) { mbar_init(&a_full[i], 1); mbar_init(&a_empty[i], 1); }
    for (int i = 0; i < EO_NSB; ++i) { mbar_init(&b_full[i], 1); mbar_init(&b_empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&t_full[i], 1); mbar_init(&t_empty[i], EO_EPI_WARPS); }
    fence_mbar_init();
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(tmem_slot)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();

  if (warp == 0 && lane == 0) {
    // ---------------- TMA producer ----------------
    uint32_t cb = 0;
    int it = 0;
    long long w_be = 0, w_ae = 0;
    for (int item = blockIdx.x; item < p.nitems; item += gridDim.x, ++it) {
      const int bc = item / p.ntr, lat0 = (item - bc * p.ntr) * 128;
      const int b = bc / p.C, c = bc - b * p.C;
      for (int nt = 0; nt < p.ntn; ++nt)
        for (int kb = 0; kb < 2 * EO_KB; ++kb) {
          const bool odd = kb >= EO_KB;
          const int kl = odd ? kb - EO_KB : kb;
          if (nt == 0) {
            // A block kb of this item: 128 latitudes x 16 orders of one parity x (re, im); free once the previous item's
            // last column tile has consumed it
            EO_TIMED(w_ae, mbar_wait_bounded(&a_empty[kb], (uint32_t)((it & 1) ^ 1)));
            mbar_arrive_expect_tx(&a_full[kb], EO_BLK);
#pragma unroll
            for (int j = 0; j < 4; ++j)
              tma_load_5d(a_blk + kb * EO_BLK + j * (TC_BK * 128), odd ? &tmAo : &tmAe, &a_full[kb], lat0 + 32 * j, 0, c, kl * (TC_BK / 2), b);
          }
          const uint32_t s = cb % EO_NSB;
          EO_TIMED(w_be, mbar_wait_bounded(&b_empty[s], ((cb / EO_NSB) & 1u) ^ 1u));
          mbar_arrive_expect_tx(&b_full[s], EO_BLK);
          tma_load_2d(b_ring + s * EO_BLK, odd ? &tmBo : &tmBe, &b_full[s], kl * TC_BK, nt * EO_BN);
          ++cb;
        }
    }
    if (p.trace) { p.trace[blockIdx.x * 8 + 6] = w_be; p.trace[blockIdx.x * 8 + 7] = w_ae; }
  } else if (warp == 1 && lane == 0) {
    // ---------------- MMA issuer ----------------
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | ((uint32_t)(EO_BN >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    uint32_t cb = 0, t = 0;
    int it = 0;
    long long w_te = 0, w_bf = 0, w_af = 0;
    const long long t_begin = clock64();
    for (int item = blockIdx.x; item < p.nitems; item += gridDim.x, ++it) {
      for (int nt = 0; nt < p.ntn; ++nt, ++t) {
        const uint32_t acc = t & 1u;
        EO_TIMED(w_te, mbar_wait_bounded(&t_empty[acc], ((t >> 1) & 1u) ^ 1u));   // the epilogue has drained this accumulator pair
        tc_fence_after();
        for (int kb = 0; kb < 2 * EO_KB; ++kb) {
          const bool odd = kb >= EO_KB;
          if (nt == 0) EO_TIMED(w_af, mbar_wait_bounded(&a_full[kb], (uint32_t)(it & 1)));
          const uint32_t s = cb % EO_NSB;
          EO_TIMED(w_bf, mbar_wait_bounded(&b_full[s], (cb / EO_NSB) & 1u));
          tc_fence_after();
          const uint32_t sa = base + (uint32_t)kb * EO_BLK;
          const uint32_t sb = base + (uint32_t)(2 * EO_KB + s) * EO_BLK;
          const uint32_t d = tmem_base + acc * 256u + (odd ? 128u : 0u);
#pragma unroll
          for (int k = 0; k < TC_BK / 8; ++k)
            tc_mma_tf32(d, make_smem_desc(sa + 1024 * k, TC_BK * 128, 512, 1), make_smem_desc(sb + 32 * k, 16, 1024), idesc,
                        ((kb & (EO_KB - 1)) | k) ? 1u : 0u);
          tc_commit(&b_empty[s]);
          if (nt == p.ntn - 1) tc_commit(&a_empty[kb]);
          ++cb;
        }
        tc_commit(&t_full[acc]);
      }
    }
    if (p.trace) {
      p.trace[blockIdx.x * 8 + 0] = clock64() - t_begin;
      p.trace[blockIdx.x * 8 + 1] = w_te; p.trace[blockIdx.x * 8 + 2] = w_bf; p.trace[blockIdx.x * 8 + 3] = w_af;
    }
  } else if (warp >= 2) {
    // ---------------- epilogue warps: TMEM lane quarter q, column half ch of every tile ----------------
    const int q = warp & 3, ch = (warp - 2) >> 2;
    uint8_t* buf = stage + (warp - 2) * 4096;
    uint32_t t = 0;
    bool pending = false;                                          // a store of this warp may still be reading buf
    long long w_tf = 0, w_rd = 0;
    for (int item = blockIdx.x; item < p.nitems; item += gridDim.x) {
      const int bc = item / p.ntr, lat0 = (item - bc * p.ntr) * 128;
      const int row0 = lat0 + q * 32;
      const bool warp_rows = row0 < p.nlat, row_ok = row0 + lane < p.nlat, all_rows = row0 + 32 <= p.nlat;
      float ls[4] = {0.f, 0.f, 0.f, 0.f}, lq[4] = {0.f, 0.f, 0.f, 0.f};
      for (int nt = 0; nt < p.ntn; ++nt, ++t) {
        const uint32_t acc = t & 1u;
        EO_TIMED(w_tf, mbar_wait_bounded(&t_full[acc], (t >> 1) & 1u));
        tc_fence_after();
        const int j0 = nt * EO_BN;
        bool released = false;
        if (warp_rows) {
#pragma unroll 1
          for (int cc = 0; cc < 2; ++cc) {
            const int c0 = ch * (EO_BN / 2) + cc * 32;
            if (j0 + c0 >= p.half) break;
            uint32_t e[32], o[32];
            const uint32_t ta = tmem_base + ((uint32_t)(q * 32) << 16) + acc * 256u + (uint32_t)c0;
            MSFNO_DFT_LD32(e, ta);
            MSFNO_DFT_LD32(o, ta + 128u);
            asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
            if (cc == 1 || j0 + c0 + 32 >= p.half) {               // last read of this accumulator pair: hand it back to the MMA warp now
              tc_fence_before();
              __syncwarp();
              if (lane == 0) mbar_arrive(&t_empty[acc]);
              released = true;
            }
#pragma unroll
            for (int hf = 0; hf < 2; ++hf) {
              float v[32];
#pragma unroll
              for (int j = 0; j < 32; ++j) {
                const float ev = __uint_as_float(e[j]), ov = __uint_as_float(o[j]);
                float tv = hf ? ev - ov : ev + ov;
                if (FLAGS & 1) tv = gelu_tanh3(tv);
                if (FLAGS & 2) tv = round_to_tf32(tv);
                v[j] = tv;
              }
              // plane statistics (four independent chains); columns beyond nlon / 2 are exact zeros (zero-filled matrix
              // rows), latitudes beyond nlat hold whatever the workspace held: masked unless the whole warp is inside
              if (all_rows) {
#pragma unroll
                for (int j = 0; j < 32; ++j) { ls[j & 3] += v[j]; lq[j & 3] = fmaf(v[j], v[j], lq[j & 3]); }
              } else {
#pragma unroll
                for (int j = 0; j < 32; ++j) { const float ts = row_ok ? v[j] : 0.0f; ls[j & 3] += ts; lq[j & 3] = fmaf(ts, ts, lq[j & 3]); }
              }
              if (pending) {                                       // the previous store of this warp has read the tile
                EO_TIMED(w_rd, if (lane == 0) bulk_wait_read<0>(); __syncwarp());
              }
#pragma unroll
              for (int jj = 0; jj < 8; ++jj)
                *reinterpret_cast<float4*>(buf + lane * 128 + ((jj ^ (lane & 7)) << 4)) = make_float4(v[4 * jj], v[4 * jj + 1], v[4 * jj + 2], v[4 * jj + 3]);
              fence_proxy_async();
              __syncwarp();
              if (lane == 0) {
                tma_store_3d(hf ? &tmO2 : &tmO1, buf, j0 + c0, row0, bc);
                bulk_commit();
              }
              pending = true;
            }
          }
        }
        if (!released) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&t_empty[acc]);
        }
      }
      if (p.stats && warp_rows) {
        double ds = (double)((ls[0] + ls[1]) + (ls[2] + ls[3])), dq = (double)((lq[0] + lq[1]) + (lq[2] + lq[3]));
        for (int of = 16; of > 0; of >>= 1) {
          ds += __shfl_xor_sync(0xffffffffu, ds, of);
          dq += __shfl_xor_sync(0xffffffffu, dq, of);
        }
        if (lane == 0) {
          atomicAdd(&p.stats[2 * bc], ds);
          atomicAdd(&p.stats[2 * bc + 1], dq);
        }
      }
    }
    if (lane == 0) bulk_wait_read<0>();                            // shared memory must outlive the engine's reads
    if (p.trace && warp == 2 && lane == 0) { p.trace[blockIdx.x * 8 + 4] = w_tf; p.trace[blockIdx.x * 8 + 5] = w_rd; }
  }
  __syncwarp();
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem_base), "r"(512));
  }
}

// ------------------------------------------------------------------------------------------ host side
static float host_round_tf32(float x) {
  uint32_t u;
  memcpy(&u, &x, 4);
  u = (u + 0x1000u) & 0xffffe000u;
  memcpy(&x, &u, 4);
  return x;
}

bool dft_tc_supported(const msfno_plan* p) {
  static const bool off = dbg_env("MSFNO_DFT_FFT");
  return !off && p->nlon % 4 == 0 && 2 * p->mlim <= DF_BN && (2 * p->mlim) % 4 == 0 && p->mlim <= p->nlon / 2 && get_encode() != nullptr;
}

// F [2 mlim][nlon]: (2 pi / nlon) (cos, -sin)(2 pi m j / nlon);  G [nlon][2 mlim]: c_m (cos, -sin), c_0 = 1 (and the
// Nyquist order), 2 otherwise; the imaginary parts of the DC and Nyquist bins do not contribute (irfft semantics)
int dft_tc_build(msfno_plan* p) {
  if (p->d_dft_fwd) return MSFNO_OK;
  const int nlon = p->nlon, N = 2 * p->mlim;
  std::vector<float> F((size_t)N * nlon), G((size_t)nlon * N);
  for (int m = 0; m < p->mlim; ++m) {
    const double cm = (m == 0 || 2 * m == nlon) ? 1.0 : 2.0;
    const bool real_only = (m == 0 || 2 * m == nlon);
    for (int j = 0; j < nlon; ++j) {
      const double a = 2.0 * M_PI * (double)(((long long)m * j) % nlon) / nlon;
      F[(size_t)(2 * m) * nlon + j] = host_round_tf32((float)(2.0 * M_PI / nlon * cos(a)));
      F[(size_t)(2 * m + 1) * nlon + j] = host_round_tf32((float)(-2.0 * M_PI / nlon * sin(a)));
      G[(size_t)j * N + 2 * m] = host_round_tf32((float)(cm * cos(a)));
      G[(size_t)j * N + 2 * m + 1] = real_only ? 0.0f : host_round_tf32((float)(-cm * sin(a)));
    }
  }
  MSFNO_CUDA_OK(cudaMalloc(&p->d_dft_fwd, F.size() * sizeof(float)));
  MSFNO_CUDA_OK(cudaMalloc(&p->d_dft_inv, G.size() * sizeof(float)));
  MSFNO_CUDA_OK(cudaMemcpy(p->d_dft_fwd, F.data(), F.size() * sizeof(float), cudaMemcpyHostToDevice));
  MSFNO_CUDA_OK(cudaMemcpy(p->d_dft_inv, G.data(), G.size() * sizeof(float), cudaMemcpyHostToDevice));
  return MSFNO_OK;
}

static int make_map_5d(CUtensorMap* tm, const float* base, const cuuint64_t dims[5], const cuuint64_t strides_bytes[4],
                       const cuuint32_t box[5]) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return record_error(MSFNO_ERR_CUDA, "cuTensorMapEncodeTiled entry point unavailable");
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 5, const_cast<float*>(base), dims, strides_bytes, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return record_error(MSFNO_ERR_CUDA, "cuTensorMapEncodeTiled failed (dft 5-D map)");
  return MSFNO_OK;
}

template <bool INV, int MT, int NS>
static int launch_dft(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmS, const CUtensorMap& tmI, const CUtensorMap& tmO,
                      const DftParams& prm, dim3 grid, cudaStream_t st) {
  constexpr int smem = NS * (MT * DF_A_BYTES + DF_B_BYTES) + 4096 + 1024 + 512;
  auto kern = dft_tc_kernel<INV, MT, NS>;
  MSFNO_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  static const bool trace_on = dbg_env("MSFNO_DFT_TRACE");
  static long long* d_trace = nullptr;
  DftParams prm2 = prm;
  if (trace_on) {
    if (!d_trace) MSFNO_CUDA_OK(cudaMalloc(&d_trace, 512 * 8 * sizeof(long long)));
    MSFNO_CUDA_OK(cudaMemsetAsync(d_trace, 0, 512 * 8 * sizeof(long long), st));
    prm2.trace = d_trace;
  }
  MSFNO_CUDA_OK(launch_pdl(kern, grid, dim3(256), smem, st, tmA, tmB, tmS, tmI, tmO, prm2));
  if (trace_on) {
    static long long h[512 * 8];
    MSFNO_CUDA_OK(cudaStreamSynchronize(st));
    MSFNO_CUDA_OK(cudaMemcpy(h, d_trace, sizeof(h), cudaMemcpyDeviceToHost));
    double a0 = 0, a1 = 0, a2 = 0, a3 = 0, a4 = 0; int n = 0;
    for (int i = 0; i < 512; ++i) if (h[i * 8 + 2]) { a0 += h[i * 8]; a1 += h[i * 8 + 1]; a2 += h[i * 8 + 2]; a3 += h[i * 8 + 3]; a4 += h[i * 8 + 4]; ++n; }
    if (n) fprintf(stderr, "dft_tc trace INV=%d MT=%d grid=(%u,%u,%u): mean clk since CTA start: prologue %.0f  mainloop_done %.0f  [staged %.0f  pass0_done %.0f]  cta_done %.0f  (n=%d)\n",
                   (int)INV, MT, grid.x, grid.y, grid.z, a0 / n, a1 / n, a3 / n, a4 / n, a2 / n, n);
  }
  count_launch();
  MSFNO_CUDA_OK(cudaGetLastError());
  return MSFNO_OK;
}

int launch_dft_fwd(msfno_plan* p, const float* x, float* Xt, const float* in_scale, const float* in_shift, int B, int C,
                   cudaStream_t st) {
  int rc = dft_tc_build(p);
  if (rc) return rc;
  CUtensorMap tmA, tmB;
  rc = make_map(&tmA, x, (long long)B * C * p->nlat, p->nlon, p->nlon, 128);
  if (rc) return rc;
  rc = make_map(&tmB, p->d_dft_fwd, 2 * p->mlim, p->nlon, p->nlon, DF_BN);
  if (rc) return rc;
  DftParams prm{};
  prm.out = Xt; prm.sc = in_scale; prm.sh = in_shift;
  prm.nlat = p->nlat; prm.nlon = p->nlon; prm.mlim = p->mlim; prm.kpad = p->kpad; prm.C = C;
  prm.nkb = (p->nlon + TC_BK - 1) / TC_BK;
  prm.flags = 1;
  prm.dc = (float)(2.0 * M_PI);
  static const int fwd_mt = dbg_env_int("MSFNO_DFT_FWD_MT", 2);
  if (p->kpad > 128 && fwd_mt == 2) return launch_dft<false, 2, 3>(tmA, tmB, tmA, tmB, tmA, prm, dim3(1, (p->kpad + 255) / 256, B * C), st);
  if (p->kpad > 128) return launch_dft<false, 1, 2>(tmA, tmB, tmA, tmB, tmA, prm, dim3(1, (p->kpad + 127) / 128, B * C), st);
  return launch_dft<false, 1, 2>(tmA, tmB, tmA, tmB, tmA, prm, dim3(1, 1, B * C), st);
}


// Even / odd halves of the inverse DFT matrix: Geo[par][j][2 mh + ri] = c_m (cos, -sin)(2 pi m j / nlon), m = 2 mh + par,
// 0 <= j < nlon / 2, K padded to 128 with zeros
static int dft_eo_build(msfno_plan* p) {
  if (p->d_dft_inv_eo) return MSFNO_OK;
  const int nlon = p->nlon, half = nlon / 2;
  std::vector<float> G((size_t)2 * half * 128, 0.0f);
  for (int m = 0; m < p->mlim; ++m) {
    const double cm = (m == 0 || 2 * m == nlon) ? 1.0 : 2.0;
    const bool real_only = (m == 0 || 2 * m == nlon);
    const int par = m & 1, mh = m >> 1;
    for (int j = 0; j < half; ++j) {
      const double a = 2.0 * M_PI * (double)(((long long)m * j) % nlon) / nlon;
      float* row = G.data() + ((size_t)par * half + j) * 128;
      row[2 * mh] = host_round_tf32((float)(cm * cos(a)));
      row[2 * mh + 1] = real_only ? 0.0f : host_round_tf32((float)(-cm * sin(a)));
    }
  }
  MSFNO_CUDA_OK(cudaMalloc(&p->d_dft_inv_eo, G.size() * sizeof(float)));
  MSFNO_CUDA_OK(cudaMemcpy(p->d_dft_inv_eo, G.data(), G.size() * sizeof(float), cudaMemcpyHostToDevice));
  return MSFNO_OK;
}

static bool dft_eo_supported(const msfno_plan* p, const float* y) {
  static const bool off = dbg_env("MSFNO_DFT_NO_EO");
  return !off && p->nlat > 128 && p->nlon % 8 == 0 && p->mlim + 1 <= 128 && (reinterpret_cast<uintptr_t>(y) & 15) == 0;
}

static int launch_idft_eo(msfno_plan* p, const float* Yt, float* y, int act_flags, double* stats, int B, int C, cudaStream_t st) {
  int rc = dft_eo_build(p);
  if (rc) return rc;
  EncodeTiledFn enc = get_encode();
  const int half = p->nlon / 2;
  const int me = (p->mlim + 1) / 2, mo = p->mlim / 2;             // even / odd orders
  CUtensorMap tmAe, tmAo, tmBe, tmBo, tmO1, tmO2;
  const cuuint64_t kp = (cuuint64_t)p->kpad;
  const cuuint64_t strides[4] = {kp * 4, 2 * kp * 4, (cuuint64_t)2 * (2 * C) * kp * 4, (cuuint64_t)p->mlim * 2 * C * kp * 4};
  const cuuint32_t box[5] = {32, 2, 1, TC_BK / 2, 1};
  const cuuint64_t dims_e[5] = {kp, 2, (cuuint64_t)C, (cuuint64_t)me, (cuuint64_t)B};
  const cuuint64_t dims_o[5] = {kp, 2, (cuuint64_t)C, (cuuint64_t)(mo > 0 ? mo : 1), (cuuint64_t)B};
  rc = make_map_5d(&tmAe, Yt, dims_e, strides, box);
  if (rc) return rc;
  rc = make_map_5d(&tmAo, Yt + (size_t)2 * C * p->kpad, dims_o, strides, box);
  if (rc) return rc;
  rc = make_map(&tmBe, p->d_dft_inv_eo, half, 128, 128, EO_BN);
  if (rc) return rc;
  rc = make_map(&tmBo, p->d_dft_inv_eo + (size_t)half * 128, half, 128, 128, EO_BN);
  if (rc) return rc;
  const cuuint64_t odims[3] = {(cuuint64_t)half, (cuuint64_t)p->nlat, (cuuint64_t)B * C};
  const cuuint64_t ostr[2] = {(cuuint64_t)p->nlon * 4, (cuuint64_t)p->nlat * p->nlon * 4};
  const cuuint32_t obox[3] = {32, 32, 1};
  const cuuint32_t oes[3] = {1, 1, 1};
  for (int h = 0; h < 2; ++h)
    if (enc(h ? &tmO2 : &tmO1, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, y + (size_t)h * half, odims, ostr, obox, oes, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      return record_error(MSFNO_ERR_CUDA, "cuTensorMapEncodeTiled failed (idft_eo output map)");
  EoParams prm{};
  prm.stats = stats;
  prm.nlat = p->nlat; prm.half = half; prm.C = C;
  prm.ntr = (p->nlat + 127) / 128;
  prm.ntn = (half + EO_BN - 1) / EO_BN;
  prm.nitems = B * C * prm.ntr;
  prm.flags = act_flags;
  static int n_sm = 0;
  if (!n_sm) {
    int dev = 0;
    MSFNO_CUDA_OK(cudaGetDevice(&dev));
    MSFNO_CUDA_OK(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev));
  }
  auto kern = (act_flags & 3) == 0 ? idft_eo_kernel<0> : (act_flags & 3) == 1 ? idft_eo_kernel<1> : (act_flags & 3) == 2 ? idft_eo_kernel<2> : idft_eo_kernel<3>;
  MSFNO_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, EO_SMEM));
  const int grid = prm.nitems < n_sm ? prm.nitems : n_sm;
  static const bool trace_on = dbg_env("MSFNO_DFT_TRACE");
  static long long* d_trace = nullptr;
  if (trace_on) {
    if (!d_trace) MSFNO_CUDA_OK(cudaMalloc(&d_trace, 512 * 8 * sizeof(long long)));
    MSFNO_CUDA_OK(cudaMemsetAsync(d_trace, 0, 512 * 8 * sizeof(long long), st));
    prm.trace = d_trace;
  }
  MSFNO_CUDA_OK(launch_pdl(kern, dim3(grid), dim3(EO_THREADS), EO_SMEM, st, tmAe, tmAo, tmBe, tmBo, tmO1, tmO2, prm));
  if (trace_on) {
    static long long h[512 * 8];
    MSFNO_CUDA_OK(cudaStreamSynchronize(st));
    MSFNO_CUDA_OK(cudaMemcpy(h, d_trace, sizeof(h), cudaMemcpyDeviceToHost));
    double a[8] = {0};
    const int n = grid < 512 ? grid : 512;
    for (int i = 0; i < n; ++i) for (int k = 0; k < 8; ++k) a[k] += (double)h[i * 8 + k];
    fprintf(stderr, "idft_eo trace grid=%d items=%d tiles/item=%d: mean clk per CTA: mma_total %.0f | mma waits: t_empty %.0f b_full %.0f a_full %.0f | "
                    "epilogue(warp 2) waits: t_full %.0f store_read %.0f | producer waits: b_empty %.0f a_empty %.0f\n",
            grid, prm.nitems, prm.ntn, a[0] / n, a[1] / n, a[2] / n, a[3] / n, a[4] / n, a[5] / n, a[6] / n, a[7] / n);
  }
  count_launch();
  MSFNO_CUDA_OK(cudaGetLastError());
  return MSFNO_OK;
}

int launch_dft_inv(msfno_plan* p, const float* Yt, float* y, const float* skip, int act_flags, double* stats, int B, int C,
                   cudaStream_t st) {
  if (skip == nullptr && dft_eo_supported(p, y)) return launch_idft_eo(p, Yt, y, act_flags, stats, B, C, st);
  int rc = dft_tc_build(p);
  if (rc) return rc;
  CUtensorMap tmA, tmB;
  const cuuint64_t kp = (cuuint64_t)p->kpad;
  const cuuint64_t dims[5] = {kp, 2, (cuuint64_t)C, (cuuint64_t)p->mlim, (cuuint64_t)B};
  const cuuint64_t strides[4] = {kp * 4, 2 * kp * 4, (cuuint64_t)2 * C * kp * 4, (cuuint64_t)p->mlim * 2 * C * kp * 4};
  const cuuint32_t box[5] = {32, 2, 1, TC_BK / 2, 1};
  rc = make_map_5d(&tmA, Yt, dims, strides, box);
  if (rc) return rc;
  rc = make_map(&tmB, p->d_dft_inv, p->nlon, 2 * p->mlim, 2 * p->mlim, DF_BN);
  if (rc) return rc;
  DftParams prm{};
  prm.out = y; prm.skip = skip; prm.stats = stats;
  prm.nlat = p->nlat; prm.nlon = p->nlon; prm.mlim = p->mlim; prm.kpad = p->kpad; prm.C = C;
  prm.nkb = (2 * p->mlim + TC_BK - 1) / TC_BK;
  prm.flags = act_flags;
  const int tilesN = (p->nlon + DF_BN - 1) / DF_BN;
  static const int inv_mt = dbg_env_int("MSFNO_DFT_INV_MT", 1);   // 1: 128-row tiles, two CTAs per SM overlap each other's load / MMA / store phases (faster than one 256-row tile per SM)
  // skip operand on the tensor cores: TMA-able tensor (16-byte aligned rows) and the shared identity block
  CUtensorMap tmS = tmB, tmI = tmB;
  const float* d_ident = identity32_device();
  static const bool skip_lsu = dbg_env("MSFNO_DFT_SKIP_LSU");
  if (skip && !skip_lsu && d_ident != nullptr && (reinterpret_cast<uintptr_t>(skip) & 15) == 0) {
    rc = make_map(&tmS, skip, (long long)B * C * p->nlat, p->nlon, p->nlon, 128);
    if (rc) return rc;
    rc = make_map(&tmI, d_ident, 32, 32, 32, 32);
    if (rc) return rc;
    prm.skip_tma = 1;
  }
  // output through TMA tensor stores: [B C][nlat][nlon] map, 32 x 32 boxes in the 128-byte-swizzle layout
  CUtensorMap tmO = tmB;
  static const bool out_lsu = dbg_env("MSFNO_DFT_OUT_LSU");
  if (!out_lsu && (skip == nullptr || prm.skip_tma) && (reinterpret_cast<uintptr_t>(y) & 15) == 0) {
    EncodeTiledFn enc = get_encode();
    const cuuint64_t odims[3] = {(cuuint64_t)p->nlon, (cuuint64_t)p->nlat, (cuuint64_t)B * C};
    const cuuint64_t ostr[2] = {(cuuint64_t)p->nlon * 4, (cuuint64_t)p->nlat * p->nlon * 4};
    const cuuint32_t obox[3] = {32, 32, 1};
    const cuuint32_t oes[3] = {1, 1, 1};
    if (enc(&tmO, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, y, odims, ostr, obox, oes, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      return record_error(MSFNO_ERR_CUDA, "cuTensorMapEncodeTiled failed (dft output map)");
    prm.out_tma = 1;
  }
  if (p->nlat > 128 && inv_mt == 2) return launch_dft<true, 2, 3>(tmA, tmB, tmS, tmI, tmO, prm, dim3(tilesN, (p->nlat + 255) / 256, B * C), st);
  if (p->nlat > 128) return launch_dft<true, 1, 2>(tmA, tmB, tmS, tmI, tmO, prm, dim3(tilesN, (p->nlat + 127) / 128, B * C), st);
  return launch_dft<true, 1, 2>(tmA, tmB, tmS, tmI, tmO, prm, dim3(tilesN, 1, B * C), st);
}

}  // namespace msfno
