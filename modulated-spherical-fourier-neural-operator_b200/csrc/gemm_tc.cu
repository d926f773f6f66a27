// Grouped TF32 GEMM on the 5th-generation tensor cores: tcgen05.mma (kind::tf32) issued by one thread,
// operands staged in shared memory by TMA (cp.async.bulk.tensor, 128-byte swizzle), fp32 accumulators in
// TMEM, read back with tcgen05.ld for the epilogue.  This is the "tensor-core tier" (<= 2e-3 rel-L2) engine of
// the spectral complex MLP, of the Legendre contractions and of the 1x1-conv MLPs either side of the path.
//
// replaces: the cuBLAS / cuDNN GEMMs behind torch.einsum and nn.Conv2d(1x1) in /root/reference
//   MSFNO/Models/sfno/contractions.py:132-137 ("bixy,io->boxy"), torch_harmonics' Legendre einsums and
//   MSFNO/Models/sfno/layers.py:161-168 (MLP.fwd).
//
// D[M][N] = A[M][K] * op(B) (+ A2 * op(B2)),  D = act(D + bias[row]) + add
//   A  : K-major  [M][lda]                      (weights / tables)
//   B  : K-major  [N][ldb]   (B_MN = false)     or   MN-major [K][ldb], N contiguous (B_MN = true: NCHW activations)
// Operands are described to TMA as plain 2-D tensors; a group selects its sub-problem by (row, column)
// coordinates, so one tensor map per operand serves all groups of a launch.  Out-of-range boxes are zero-filled.
//
// Warp roles (256 threads): warp 0 = TMA producer, warp 1 = MMA issuer, warp 2 = TMEM allocator; afterwards all
// 8 warps drain the accumulator (TMEM lane quarter q = warp % 4, column half = warp / 4).  BLOCK_M = 128, BLOCK_N = 128, BLOCK_K = 32 fp32
// (= one 128-byte swizzle span = 4 UMMA K-steps of 8).
#include <cstdio>
#include <cstdlib>

#include "plan.h"
#include "tc_common.cuh"

namespace msfno {

static constexpr int TC_BM = 128, TC_BN = 128, TC_STAGES = 3;  // 3 x 32 KB stages -> two CTAs per SM (one's epilogue overlaps the other's main loop)
static constexpr int TC_A_BYTES = TC_BM * TC_BK * 4;  // 16 KB
static constexpr int TC_B_BYTES = TC_BN * TC_BK * 4;  // 16 KB
static constexpr int TC_STAGE_BYTES = TC_A_BYTES + TC_B_BYTES;
static constexpr int TC_SMEM_BYTES = TC_STAGES * TC_STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;

struct TcParams {
  float* D;
  long long* trace;   // debug (MSFNO_GEMM_TRACE): clock stamps of every CTA of the pair kernel, or null
  long long lda, ldb, ldd;
  const GemmGroup* groups;
  GemmGroup single;
  long long sa, sb, sd;
  int use_single;
  int relu_even;
  int round_tf32;
  int tilesN, tilesM;
  const float* bias; long long sbias;
  const float* add; long long ldadd, sadd;
  int act_gelu;
  long long lda2, ldb2, sa2, sb2;
  int K2;
};

template <bool B_MN>
__global__ void __launch_bounds__(256, 2)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               const __grid_constant__ CUtensorMap tmA2, const __grid_constant__ CUtensorMap tmB2, TcParams p) {
  extern __shared__ uint8_t smem_raw[];
  GemmGroup grp;
  if (p.use_single) {
    grp = p.single;
    grp.a_off += blockIdx.y * p.sa;
    grp.b_off += blockIdx.y * p.sb;
    grp.d_off += blockIdx.y * p.sd;
  } else {
    grp = p.groups[blockIdx.y];
  }
  // M-tiles of one N-tile are adjacent in launch order, so the (large) B operand of that N-tile is re-read from L2
  const int tn = blockIdx.x / p.tilesM, tm = blockIdx.x - tn * p.tilesM;
  const int m0 = tm * TC_BM, n0 = tn * TC_BN;
  if (m0 >= grp.M || n0 >= grp.N) return;  // uniform for the CTA, before any barrier / allocation
  pdl_trigger();
  const long long t_start = clock64();
  const int cta_lin = blockIdx.y * gridDim.x + blockIdx.x;

  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;  // swizzle-128B tiles need 1024-byte alignment
  uint8_t* tiles = smem_raw + (base - smem_u32(smem_raw));
  uint64_t* bars = reinterpret_cast<uint64_t*>(tiles + TC_STAGES * TC_STAGE_BYTES);
  uint64_t* full = bars;
  uint64_t* empty = bars + TC_STAGES;
  uint64_t* tmem_full = bars + 2 * TC_STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * TC_STAGES + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nkb1 = (grp.K + TC_BK - 1) / TC_BK;
  const int nkb = nkb1 + (p.K2 + TC_BK - 1) / TC_BK;

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < TC_STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    mbar_init(tmem_full, 1);
    fence_mbar_init();
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(tmem_slot)), "r"(TC_BN));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();   // the prologue above overlapped the previous kernel; its outputs are visible from here on
  if (p.trace && threadIdx.x == 0 && cta_lin < 1024) p.trace[cta_lin * 8 + 1] = clock64() - t_start;

  if (nkb > 0) {
    if (warp == 0 && lane == 0) {
      // ---------------- TMA producer ----------------
      // first pair: offsets come from the group; second pair: plain strided batch (blockIdx.y * s?2)
      const long long a_off2 = blockIdx.y * p.sa2, b_off2 = blockIdx.y * p.sb2;
      for (int kb = 0; kb < nkb; ++kb) {
        const int s = kb % TC_STAGES;
        const uint32_t ph = (uint32_t)((kb / TC_STAGES) & 1);
        const bool second = kb >= nkb1;
        const int kk = (second ? kb - nkb1 : kb) * TC_BK;
        const long long aoff = second ? a_off2 : grp.a_off, boff = second ? b_off2 : grp.b_off;
        const long long lda = second ? p.lda2 : p.lda, ldb = second ? p.ldb2 : p.ldb;
        const CUtensorMap* ma = second ? &tmA2 : &tmA;
        const CUtensorMap* mb = second ? &tmB2 : &tmB;
        mbar_wait_bounded(&empty[s], ph ^ 1u);
        mbar_arrive_expect_tx(&full[s], TC_STAGE_BYTES);
        uint8_t* sa = tiles + s * TC_STAGE_BYTES;
        tma_load_2d(sa, ma, &full[s], (int)(aoff % lda) + kk, (int)(aoff / lda) + m0);
        if (!B_MN) {
          tma_load_2d(sa + TC_A_BYTES, mb, &full[s], (int)(boff % ldb) + kk, (int)(boff / ldb) + n0);
        } else {
          // [K][N] operand: four boxes of 32 (n) x 32 (k); box j holds n in [n0 + 32 j, +32)
          const int ncol = (int)(boff % ldb) + n0, krow = (int)(boff / ldb) + kk;
#pragma unroll
          for (int j = 0; j < TC_BN / 32; ++j)
            tma_load_2d(sa + TC_A_BYTES + j * (TC_BK * 128), mb, &full[s], ncol + 32 * j, krow);
        }
      }
    } else if (warp == 1 && lane == 0) {
      // ---------------- MMA issuer ----------------
      // instruction descriptor: D=f32 (bit 4), A=B=tf32 (bits 7,10), B major (bit 16), N>>3 at bit 17, M>>4 at bit 24
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((B_MN ? 1u : 0u) << 16) |
                             ((uint32_t)(TC_BN >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);
      for (int kb = 0; kb < nkb; ++kb) {
        const int s = kb % TC_STAGES;
        const uint32_t ph = (uint32_t)((kb / TC_STAGES) & 1);
        mbar_wait_bounded(&full[s], ph);
        tc_fence_after();
        const uint32_t sa = base + s * TC_STAGE_BYTES;
        const uint32_t sb = sa + TC_A_BYTES;
#pragma unroll
        for (int k = 0; k < TC_BK / 8; ++k) {
          // A (K-major): advance 8 tf32 = 32 bytes inside the swizzle span
          const uint64_t adesc = make_smem_desc(sa + 32 * k, 16, 1024);
          // B K-major: same; B MN-major (SW128_BASE32B): 8 k-rows = two 512-byte atoms (SBO), N blocks 4096 bytes apart (LBO)
          const uint64_t bdesc = B_MN ? make_smem_desc(sb + 1024 * k, TC_BK * 128, 512, 1) : make_smem_desc(sb + 32 * k, 16, 1024);
          tc_mma_tf32(tmem_base, adesc, bdesc, idesc, (kb | k) ? 1u : 0u);
        }
        tc_commit(&empty[s]);  // frees the smem slot once these MMAs have read it
      }
      tc_commit(tmem_full);    // accumulator complete
    }
  }

  __syncwarp();  // producer / MMA lanes rejoin their warps: every warp takes part in the epilogue
  {
    // ---------------- epilogue: TMEM -> registers -> global ----------------
    // warp w drains TMEM lane quarter q = w % 4 (hardware restriction) and column half w / 4
    const int q = warp & 3;
    const int chalf = warp >> 2;
    const int row = m0 + q * 32 + lane;
    if (nkb > 0) {
      mbar_wait_bounded(tmem_full, 0);
      tc_fence_after();
      if (p.trace && threadIdx.x == 0 && cta_lin < 1024) p.trace[cta_lin * 8 + 2] = clock64() - t_start;
    }
    float* drow = p.D + grp.d_off + (long long)row * p.ldd;
    const bool vec = (((grp.d_off | p.ldd) & 3) == 0) && ((reinterpret_cast<uintptr_t>(p.D) & 15) == 0);
    const float bv = (p.bias && row < grp.M) ? p.bias[blockIdx.y * p.sbias + row] : 0.0f;
    const float* arow = p.add ? p.add + blockIdx.y * p.sadd + (long long)row * p.ldadd : nullptr;
#pragma unroll 1
    for (int c0 = chalf * (TC_BN / 2); c0 < (chalf + 1) * (TC_BN / 2); c0 += 32) {
      uint32_t r[32];
      if (nkb > 0) {
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0;
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
            "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
            "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
            : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
              "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
              "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
              "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
            : "r"(taddr));
        asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) r[j] = 0u;
      }
      const int gn = n0 + c0;
      if (gn < grp.N) {                                     // warp-uniform
        const bool full_vec = vec && gn + 31 < grp.N;       // warp-uniform
        const bool row_ok = row < grp.M;
        float v[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          float t = __uint_as_float(r[j]) + bv;
          if (p.act_gelu) t = gelu_tanh3(t);
          v[j] = t;
        }
        if (arow && row_ok) {
          if (full_vec && ((p.ldadd | p.sadd) & 3) == 0 && (reinterpret_cast<uintptr_t>(p.add) & 15) == 0) {
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              const float4 a4 = *reinterpret_cast<const float4*>(arow + gn + j);
              v[j] += a4.x; v[j + 1] += a4.y; v[j + 2] += a4.z; v[j + 3] += a4.w;
            }
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (gn + j < grp.N) v[j] += arow[gn + j];
          }
        }
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          if (p.relu_even && !(j & 1)) v[j] = fmaxf(v[j], 0.f);
          if (p.round_tf32) v[j] = round_to_tf32(v[j]);
        }
        if (vec) {   // all lanes take part: rows beyond M and columns beyond N are masked inside
          store_block_transposed(v, reinterpret_cast<float*>(tiles) + warp * (32 * 36),
                                 p.D + grp.d_off + (long long)(m0 + q * 32) * p.ldd + gn, p.ldd, grp.M - (m0 + q * 32), lane, grp.N - gn);
        } else if (row_ok) {
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (gn + j < grp.N) drow[gn + j] = v[j];
        }
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (p.trace && threadIdx.x == 0 && cta_lin < 1024) p.trace[cta_lin * 8 + 3] = clock64() - t_start;
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem_base), "r"(TC_BN));
  }
}

// ---------------------------------------------------------------------------------------------
// CTA-pair variant (cta_group::2): two CTAs of a cluster (the two SMs of a TPC) compute one 256 x 256 tile.
// Each CTA stages ITS 128 rows of A and ITS 128 rows of B (= 128 of the 256 output columns); the leader's single
// thread issues tcgen05.mma.cta_group::2 (M = 256, N = 256), which reads both CTAs' shared memory and writes each
// CTA's 128 x 256 accumulator slice into that CTA's TMEM.  With 4-byte operands a single-CTA 128 x 128 tile needs
// 128 B/clk of smem reads for the MMA plus 128 B/clk of TMA writes, twice what an SM's shared memory delivers; the
// pair halves both (each operand byte is staged once per pair), which is what lets the TF32 MLP GEMMs leave the
// 40 % plateau.  Barriers: full[s] lives in the leader (both producers' TMA bytes complete on it), empty[s] and
// tmem_full exist in both CTAs and are signalled by multicast commits.  K-major B, single operand pair only.
static constexpr int TC2_BN = 256;
static constexpr int TC2_STAGE_BYTES = TC_A_BYTES + TC_A_BYTES;   // 128 rows of A + 128 rows of B per CTA
static constexpr int TC2_SMEM_BYTES = TC_STAGES * TC2_STAGE_BYTES + 1024 + 256;

__device__ __forceinline__ uint32_t __smid() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%smid;\n" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;\n" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;\n" ::: "memory");
}
// TMA load whose completion bytes land on the LEADER CTA's barrier (peer bit of the barrier address cleared)
__device__ __forceinline__ void tma_load_2d_pair(void* smem_dst, const CUtensorMap* tm, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];\n" ::"r"(
          smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(tm)), "r"(smem_u32(bar) & 0xFEFFFFFFu), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tc_mma_tf32_pair(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive on the barrier at this shared-memory offset in BOTH CTAs once the MMAs issued so far have completed
__device__ __forceinline__ void tc_commit_pair(uint64_t* bar) {
  const uint16_t mask = 3;
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;\n" ::"r"(smem_u32(bar)),
               "h"(mask)
               : "memory");
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(256, 2)
gemm_tc2_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, TcParams p) {
  extern __shared__ uint8_t smem_raw[];
  GemmGroup grp;
  if (p.use_single) {
    grp = p.single;
    grp.a_off += blockIdx.y * p.sa;
    grp.b_off += blockIdx.y * p.sb;
    grp.d_off += blockIdx.y * p.sd;
  } else {
    grp = p.groups[blockIdx.y];
  }
  const uint32_t rank = cluster_ctarank();
  const int pair = blockIdx.x >> 1;
  const int tn = pair / p.tilesM, tm = pair - tn * p.tilesM;   // tilesM counts 256-row pair tiles here
  const int m0 = tm * 256 + (int)rank * TC_BM;                 // this CTA's 128 rows
  const int n0 = tn * TC2_BN;
  if (tm * 256 >= grp.M || n0 >= grp.N) return;                // uniform for the PAIR
  pdl_trigger();
  long long t_start = 0;
  if (p.trace && threadIdx.x == 0) { t_start = clock64(); p.trace[blockIdx.x * 8 + 0] = (long long)(__smid()); }

  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* tiles = smem_raw + (base - smem_u32(smem_raw));
  uint64_t* bars = reinterpret_cast<uint64_t*>(tiles + TC_STAGES * TC2_STAGE_BYTES);
  uint64_t* full = bars;
  uint64_t* empty = bars + TC_STAGES;
  uint64_t* tmem_full = bars + 2 * TC_STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * TC_STAGES + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nkb = (grp.K + TC_BK - 1) / TC_BK;

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < TC_STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    mbar_init(tmem_full, 1);
    fence_mbar_init();
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(tmem_slot)), "r"(TC2_BN));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;\n");
  }
  tc_fence_before();
  cluster_sync_all();   // both CTAs' barriers are initialised and both TMEM slices allocated before any signal crosses
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();
  if (p.trace && threadIdx.x == 0) p.trace[blockIdx.x * 8 + 1] = clock64() - t_start;

  if (nkb > 0) {
    if (warp == 0 && lane == 0) {
      // ---------------- TMA producer (both CTAs) ----------------
      for (int kb = 0; kb < nkb; ++kb) {
        const int s = kb % TC_STAGES;
        const uint32_t ph = (uint32_t)((kb / TC_STAGES) & 1);
        const int kk = kb * TC_BK;
        mbar_wait_bounded(&empty[s], ph ^ 1u);
        if (rank == 0) mbar_arrive_expect_tx(&full[s], 2 * TC2_STAGE_BYTES);   // bytes of both CTAs
        uint8_t* sa = tiles + s * TC2_STAGE_BYTES;
        tma_load_2d_pair(sa, &tmA, &full[s], (int)(grp.a_off % p.lda) + kk, (int)(grp.a_off / p.lda) + m0);
        tma_load_2d_pair(sa + TC_A_BYTES, &tmB, &full[s], (int)(grp.b_off % p.ldb) + kk,
                         (int)(grp.b_off / p.ldb) + n0 + (int)rank * TC_BM);
      }
    } else if (warp == 1 && lane == 0 && rank == 0) {
      // ---------------- MMA issuer (leader CTA only) ----------------
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(TC2_BN >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
      for (int kb = 0; kb < nkb; ++kb) {
        const int s = kb % TC_STAGES;
        const uint32_t ph = (uint32_t)((kb / TC_STAGES) & 1);
        mbar_wait_bounded(&full[s], ph);
        tc_fence_after();
        const uint32_t sa = base + s * TC2_STAGE_BYTES;
        const uint32_t sb = sa + TC_A_BYTES;
#pragma unroll
        for (int k = 0; k < TC_BK / 8; ++k)
          tc_mma_tf32_pair(tmem_base, make_smem_desc(sa + 32 * k, 16, 1024), make_smem_desc(sb + 32 * k, 16, 1024), idesc,
                           (kb | k) ? 1u : 0u);
        tc_commit_pair(&empty[s]);
      }
      tc_commit_pair(tmem_full);
    }
  }

  __syncwarp();
  {
    // ---------------- epilogue: this CTA's 128 rows x 256 columns ----------------
    const int q = warp & 3;
    const int chalf = warp >> 2;
    const int row = m0 + q * 32 + lane;
    if (nkb > 0) {
      mbar_wait_bounded(tmem_full, 0);
      tc_fence_after();
      if (p.trace && threadIdx.x == 0) p.trace[blockIdx.x * 8 + 2] = clock64() - t_start;
    }
    float* drow = p.D + grp.d_off + (long long)row * p.ldd;
    const bool vec = (((grp.d_off | p.ldd) & 3) == 0) && ((reinterpret_cast<uintptr_t>(p.D) & 15) == 0);
    const float bv = (p.bias && row < grp.M) ? p.bias[blockIdx.y * p.sbias + row] : 0.0f;
#pragma unroll 1
    for (int c0 = chalf * (TC2_BN / 2); c0 < (chalf + 1) * (TC2_BN / 2); c0 += 32) {
      uint32_t r[32];
      if (nkb > 0) {
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0;
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
            "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
            "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
            : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
              "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
              "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
              "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
            : "r"(taddr));
        asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) r[j] = 0u;
      }
      const int gn = n0 + c0;
      if (gn < grp.N) {                                     // warp-uniform
        const bool full_vec = vec && gn + 31 < grp.N;       // warp-uniform
        float v[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          float t = __uint_as_float(r[j]) + bv;
          if (p.act_gelu) t = gelu_tanh3(t);
          if (p.relu_even && !(j & 1)) t = fmaxf(t, 0.f);
          if (p.round_tf32) t = round_to_tf32(t);
          v[j] = t;
        }
        if (vec) {   // all lanes take part: rows beyond M and columns beyond N are masked inside
          store_block_transposed(v, reinterpret_cast<float*>(tiles) + warp * (32 * 36),
                                 p.D + grp.d_off + (long long)(m0 + q * 32) * p.ldd + gn, p.ldd, grp.M - (m0 + q * 32), lane, grp.N - gn);
        } else if (row < grp.M) {
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (gn + j < grp.N) drow[gn + j] = v[j];
        }
      }
    }
    tc_fence_before();
  }
  if (p.trace && threadIdx.x == 0) { p.trace[blockIdx.x * 8 + 3] = clock64() - t_start; p.trace[blockIdx.x * 8 + 4] = t_start; }
  cluster_sync_all();   // neither CTA may retire (or free TMEM) while its peer can still touch it
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;\n" ::"r"(tmem_base), "r"(TC2_BN));
  }
}

// ---------------------------------------------------------------------------------------------
static bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

bool gemm_tc_supported(const GemmLaunch& g) {
  if (!g.a_kmajor || g.mask || g.accumulate) return false;
  if ((g.lda & 3) || (g.ldb & 3) || !aligned16(g.A) || !aligned16(g.B)) return false;
  if (g.A2 && ((g.lda2 & 3) || (g.ldb2 & 3) || !aligned16(g.A2) || !aligned16(g.B2))) return false;
  return get_encode() != nullptr;
}

// a_rows / b_rows: rows of the underlying 2-D buffers (TMA zero-fills beyond them); a_cols / b_cols: valid columns.
// For an MN-major B (g.b_kmajor == 0) the buffer is [K rows][N cols].
int launch_gemm_tc(const GemmLaunch& g, long long a_rows, long long a_cols, long long b_rows, long long b_cols,
                   int round_tf32, cudaStream_t st, long long a2_rows, long long a2_cols, long long b2_rows,
                   long long b2_cols) {
  if (g.ngroups <= 0 || g.maxM <= 0 || g.maxN <= 0) return MSFNO_OK;
  const bool bmn = !g.b_kmajor;
  static const bool pair_off = getenv("MSFNO_GEMM_NO_PAIR") != nullptr;
  // 256-column pair tiles: with N = 512 only 60 pairs exist for the 7440-row MLP (< 74 TPCs) and the single-CTA
  // kernel's finer tiles win (tools/gemm_bench.py); MSFNO_GEMM_PAIR_MIN_N overrides for experiments
  static const int pair_min_n = getenv("MSFNO_GEMM_PAIR_MIN_N") ? atoi(getenv("MSFNO_GEMM_PAIR_MIN_N")) : 768;
  if (!bmn && !g.A2 && !g.add && g.maxM >= 1024 && g.maxN >= pair_min_n && !pair_off) {
    CUtensorMap tmA, tmB;
    int rc = make_map(&tmA, g.A, a_rows, a_cols, g.lda, TC_BM);
    if (rc) return rc;
    rc = make_map(&tmB, g.B, b_rows, b_cols, g.ldb, TC_BM);
    if (rc) return rc;
    static std::once_flag once2;
    static cudaError_t attr_err2 = cudaSuccess;
    std::call_once(once2, [] { attr_err2 = cudaFuncSetAttribute(gemm_tc2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TC2_SMEM_BYTES); });
    MSFNO_CUDA_OK(attr_err2);
    TcParams p{};
    p.D = g.D; p.lda = g.lda; p.ldb = g.ldb; p.ldd = g.ldd;
    p.groups = g.groups; p.single = g.single; p.sa = g.sa; p.sb = g.sb; p.sd = g.sd; p.use_single = g.use_single;
    p.relu_even = g.relu_even; p.round_tf32 = round_tf32;
    p.bias = g.bias; p.sbias = g.sbias; p.act_gelu = g.act_gelu;
    p.tilesM = (g.maxM + 255) / 256;
    p.tilesN = (g.maxN + TC2_BN - 1) / TC2_BN;
    dim3 grid(2 * p.tilesM * p.tilesN, g.ngroups);
    static const bool trace_on = getenv("MSFNO_GEMM_TRACE") != nullptr;
    static long long* d_trace = nullptr;
    if (trace_on) {
      if (!d_trace) MSFNO_CUDA_OK(cudaMalloc(&d_trace, 1024 * 8 * sizeof(long long)));
      MSFNO_CUDA_OK(cudaMemsetAsync(d_trace, 0, 1024 * 8 * sizeof(long long), st));
      if (grid.x <= 1024 && grid.y == 1) p.trace = d_trace;
    }
    MSFNO_CUDA_OK(launch_pdl(gemm_tc2_kernel, grid, dim3(256), TC2_SMEM_BYTES, st, tmA, tmB, p));
    if (trace_on && p.trace) {
      static long long h[1024 * 8];
      MSFNO_CUDA_OK(cudaStreamSynchronize(st));
      MSFNO_CUDA_OK(cudaMemcpy(h, d_trace, sizeof(h), cudaMemcpyDeviceToHost));
      long long t0 = h[4];
      for (unsigned i = 0; i < grid.x; ++i) if (h[i * 8 + 4] && h[i * 8 + 4] < t0) t0 = h[i * 8 + 4];
      fprintf(stderr, "gemm_tc2 trace: cta sm | start(rel) prologue_done mainloop_done epilogue_done (clk since CTA start)\n");
      for (unsigned i = 0; i < grid.x; i += (grid.x > 64 ? 7 : 1))
        fprintf(stderr, "%4u %3lld | %8lld %8lld %8lld %8lld\n", i, h[i * 8], h[i * 8 + 4] - t0, h[i * 8 + 1], h[i * 8 + 2], h[i * 8 + 3]);
    }
    count_launch();
    MSFNO_CUDA_OK(cudaGetLastError());
    return MSFNO_OK;
  }
  CUtensorMap tmA, tmB, tmA2, tmB2;
  int rc = make_map(&tmA, g.A, a_rows, a_cols, g.lda, TC_BM);
  if (rc) return rc;
  rc = make_map(&tmB, g.B, b_rows, b_cols, g.ldb, bmn ? TC_BK : TC_BN, bmn);
  if (rc) return rc;
  if (g.A2) {
    rc = make_map(&tmA2, g.A2, a2_rows, a2_cols, g.lda2, TC_BM);
    if (rc) return rc;
    rc = make_map(&tmB2, g.B2, b2_rows, b2_cols, g.ldb2, bmn ? TC_BK : TC_BN, bmn);
    if (rc) return rc;
  } else {
    tmA2 = tmA;
    tmB2 = tmB;
  }
  static std::once_flag once;
  static cudaError_t attr_err = cudaSuccess;
  std::call_once(once, [] {
    attr_err = cudaFuncSetAttribute(gemm_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_BYTES);
    if (attr_err == cudaSuccess)
      attr_err = cudaFuncSetAttribute(gemm_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_BYTES);
  });
  MSFNO_CUDA_OK(attr_err);
  TcParams p{};
  p.D = g.D; p.lda = g.lda; p.ldb = g.ldb; p.ldd = g.ldd;
  p.groups = g.groups; p.single = g.single; p.sa = g.sa; p.sb = g.sb; p.sd = g.sd; p.use_single = g.use_single;
  p.relu_even = g.relu_even; p.round_tf32 = round_tf32;
  p.bias = g.bias; p.sbias = g.sbias; p.add = g.add; p.ldadd = g.ldadd; p.sadd = g.sadd; p.act_gelu = g.act_gelu;
  p.lda2 = g.A2 ? g.lda2 : 4; p.ldb2 = g.A2 ? g.ldb2 : 4; p.sa2 = g.sa2; p.sb2 = g.sb2; p.K2 = g.A2 ? g.K2 : 0;
  const int tilesM = (g.maxM + TC_BM - 1) / TC_BM;
  p.tilesN = (g.maxN + TC_BN - 1) / TC_BN;
  p.tilesM = tilesM;
  dim3 grid(tilesM * p.tilesN, g.ngroups);
  static const bool trace1_on = getenv("MSFNO_GEMM_TRACE") != nullptr;
  static long long* d_trace1 = nullptr;
  if (trace1_on && (long long)grid.x * grid.y <= 1024) {
    if (!d_trace1) MSFNO_CUDA_OK(cudaMalloc(&d_trace1, 1024 * 8 * sizeof(long long)));
    MSFNO_CUDA_OK(cudaMemsetAsync(d_trace1, 0, 1024 * 8 * sizeof(long long), st));
    p.trace = d_trace1;
  }
  if (bmn) MSFNO_CUDA_OK(launch_pdl(gemm_tc_kernel<true>, grid, dim3(256), TC_SMEM_BYTES, st, tmA, tmB, tmA2, tmB2, p));
  else MSFNO_CUDA_OK(launch_pdl(gemm_tc_kernel<false>, grid, dim3(256), TC_SMEM_BYTES, st, tmA, tmB, tmA2, tmB2, p));
  if (p.trace) {
    static long long h[1024 * 8];
    MSFNO_CUDA_OK(cudaStreamSynchronize(st));
    MSFNO_CUDA_OK(cudaMemcpy(h, d_trace1, sizeof(h), cudaMemcpyDeviceToHost));
    double a1 = 0, a2 = 0, a3 = 0; int n = 0;
    for (int i = 0; i < 1024; ++i) if (h[i * 8 + 3]) { a1 += h[i * 8 + 1]; a2 += h[i * 8 + 2]; a3 += h[i * 8 + 3]; ++n; }
    if (n) fprintf(stderr, "gemm_tc trace grid=(%u,%u) maxM=%d maxN=%d: mean clk since CTA start: prologue %.0f  mainloop_done %.0f  cta_done %.0f  (n=%d)\n",
                   grid.x, grid.y, g.maxM, g.maxN, a1 / n, a2 / n, a3 / n, n);
  }
  count_launch();
  MSFNO_CUDA_OK(cudaGetLastError());
  return MSFNO_OK;
}

}  // namespace msfno
