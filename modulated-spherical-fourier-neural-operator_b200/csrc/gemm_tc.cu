// Grouped GEMM on the 5th-generation tensor cores: tcgen05.mma (kind::tf32) issued by one thread, operands staged
// in shared memory by TMA (cp.async.bulk.tensor, 128-byte swizzle), fp32 accumulators in TMEM, read back with
// tcgen05.ld for the epilogue.  Engine of the spectral complex MLP, of the Legendre contractions (forward AND adjoint)
// and of the 1x1 convolutions either side of the path, in BOTH precision tiers:
//
//   tf32 tier (<= 2e-3 rel-L2): one MMA per k-step on operands the producers rounded to TF32.
//   fp32 tier (<= 1e-5 rel-L2): "3xTF32".  The tensor core reads only the upper 19 bits of an fp32 operand, i.e. it
//     sees hi(x) = x with the low 13 mantissa bits cleared.  lo(x) = x - hi(x) is exact in fp32, so
//         a b  =  hi(a) hi(b) + hi(a) lo(b) + lo(a) hi(b)  +  lo(a) lo(b)          (last term <= 2^-20 |a b|, dropped)
//     Four converter warps compute lo() of every staged operand tile in shared memory (same swizzled layout, the
//     operation is elementwise) while the previous stage is being multiplied; the issuing thread then launches three
//     MMAs per k-step into the same TMEM accumulator, small terms first.  The truncation of lo() itself (13 -> 11
//     bits) leaves <= 3 * 2^-24 |x| per operand.  The reference computes these contractions in fp32 (autocast is forced
//     off, /root/reference MSFNO/Models/sfno/layers.py:403-407,627-639); this is that tier on the tensor pipe.
//
// replaces: the cuBLAS / cuDNN GEMMs behind torch.einsum and nn.Conv2d(1x1) in /root/reference
//   MSFNO/Models/sfno/contractions.py:132-137 ("bixy,io->boxy"), torch_harmonics' Legendre einsums, their autograd
//   adjoints, and MSFNO/Models/sfno/layers.py:161-168 (MLP.fwd).
//
// D[M][N] = op(A) * op(B) (+ A2 * op(B2)),  D = mask( act(D + bias[row]) + add )  [+= D]
//   A : K-major [M][lda]  or  MN-major [K][lda] (M contiguous: adjoint of a K-major forward operand)
//   B : K-major [N][ldb]  or  MN-major [K][ldb] (N contiguous: NCHW activations, tables in the adjoints)
// Operands are described to TMA as plain 2-D tensors; a group selects its sub-problem by (row, column) coordinates, so
// one tensor map per operand serves all groups of a launch.  Out-of-range boxes are zero-filled.  An MN-major B that
// is a stack of equally sized tables can be described as a 3-D tensor (b_group_rows) so that the K rows past the end
// of one table read zeros instead of the next table.
//
// Warp roles (256 threads): warp 0 = TMA producer, warp 1 = MMA issuer, warp 2 = TMEM allocator, warps 4-7 = lo()
// converters (fp32 tier); afterwards all 8 warps drain the accumulator (TMEM lane quarter q = warp % 4, column half =
// warp / 4).  BLOCK_M = 128, BLOCK_N = 128, BLOCK_K = 32 fp32 (= one 128-byte swizzle span = 4 UMMA K-steps of 8).
#include <cstdio>
#include <cstdlib>

#include "plan.h"
#include "tc_common.cuh"

namespace msfno {

static constexpr int TC_BM = 128, TC_BN = 128, TC_STAGES = 3;
static constexpr int TC_A_BYTES = TC_BM * TC_BK * 4;  // 16 KB
static constexpr int TC_B_BYTES = TC_BN * TC_BK * 4;  // 16 KB
static constexpr int TC_HI_BYTES = TC_A_BYTES + TC_B_BYTES;   // what TMA writes per stage
static constexpr int TC_NCONV = 128;                          // converter threads (warps 4-7)
// tf32 tier: 3 x 32 KB stages -> two CTAs per SM (one's epilogue overlaps the other's main loop)
static constexpr int TC_SMEM_BYTES = TC_STAGES * TC_HI_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;

struct TcParams {
  float* D;
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
  const float* mask; long long ldmask;
  int accumulate;
  int act_gelu;
  long long lda2, ldb2, sa2, sb2;
  int K2;
  int b_group_rows;   // > 0: B (MN-major) is a 3-D tensor [tables][b_group_rows][ldb]
  long long* trace;   // builds with -DMSFNO_TRACE only: per-role clock64 totals of CTA (0, 0)
};

__device__ __forceinline__ void tma_load_3d(void* smem_dst, const CUtensorMap* tm, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];\n" ::"r"(
          smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(tm)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}

#define MSFNO_TC_LD32(r, taddr)                                                                                          \
  asm volatile(                                                                                                          \
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "                                                                          \
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "                                          \
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"                        \
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),      \
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),           \
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),          \
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])                        \
      : "r"(taddr))

// lo(x) = x - (x with the 13 low mantissa bits cleared): what the tensor core does not see of an fp32 operand
__device__ __forceinline__ float tf32_lo(float x) { return x - __uint_as_float(__float_as_uint(x) & 0xffffe000u); }

// one stage: lo() of `bytes` bytes at src -> dst (same offsets: the swizzled layout is preserved), by `nthr` threads
__device__ __forceinline__ void convert_stage_lo(const uint8_t* src, uint8_t* dst, int bytes, int t, int nthr) {
  const float4* s4 = reinterpret_cast<const float4*>(src);
  float4* d4 = reinterpret_cast<float4*>(dst);
#pragma unroll 4
  for (int i = t; i < bytes / 16; i += nthr) {
    const float4 v = s4[i];
    d4[i] = make_float4(tf32_lo(v.x), tf32_lo(v.y), tf32_lo(v.z), tf32_lo(v.w));
  }
}

// Epilogue of one 32 x 32 block held one row per lane: v[j] = raw accumulator of (row, n0 + c0 + j).
// EXACT: fp32 tier (erf GELU instead of the fit).
template <bool EXACT>
__device__ __forceinline__ void tc_epilogue_chunk(const TcParams& p, const GemmGroup& grp, float (&v)[32], float* wtile, int q,
                                                  int lane, int m0, int n0, int c0) {
  const int gn = n0 + c0;
  if (gn >= grp.N) return;                              // warp-uniform
  const int row = m0 + q * 32 + lane;
  const bool vec = (((grp.d_off | p.ldd) & 3) == 0) && ((reinterpret_cast<uintptr_t>(p.D) & 15) == 0);
  const bool row_ok = row < grp.M;
  const bool full_vec = vec && gn + 31 < grp.N;         // warp-uniform
  if (p.bias || p.act_gelu == 1) {
    const float bv = (p.bias && row_ok) ? p.bias[blockIdx.y * p.sbias + row] : 0.0f;
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      float t = v[j] + bv;
      if (p.act_gelu == 1) t = EXACT ? gelu_erf(t) : gelu_tanh3(t);
      v[j] = t;
    }
  }
  if (p.add && row_ok && p.act_gelu == 2) {
    // act_gelu == 2: the `add` operand is the pre-activation of a GELU and MULTIPLIES as gelu'(.) (activation adjoint)
    const float* arow = p.add + blockIdx.y * p.sadd + (long long)row * p.ldadd;
#pragma unroll
    for (int j = 0; j < 32; ++j)
      if (gn + j < grp.N) v[j] *= gelu_erf_grad(arow[gn + j]);
  } else if (p.add && row_ok) {
    const float* arow = p.add + blockIdx.y * p.sadd + (long long)row * p.ldadd;
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
  if (p.mask && row_ok) {   // ReLU backward on the real parts: zero where the forward activation was not positive
    const float* mrow = p.mask + grp.d_off + (long long)row * p.ldmask;
#pragma unroll
    for (int j = 0; j < 32; j += 2)
      if (gn + j < grp.N && !(mrow[gn + j] > 0.f)) v[j] = 0.f;
  }
#pragma unroll
  for (int j = 0; j < 32; ++j) {
    if (p.relu_even && !(j & 1)) v[j] = fmaxf(v[j], 0.f);
    if (p.round_tf32) v[j] = round_to_tf32(v[j]);
  }
  if (vec) {   // all lanes take part: rows beyond M and columns beyond N are masked inside
    store_block_transposed(v, wtile, p.D + grp.d_off + (long long)(m0 + q * 32) * p.ldd + gn, p.ldd, grp.M - (m0 + q * 32),
                           lane, grp.N - gn, p.accumulate != 0);
  } else if (row_ok) {
    float* drow = p.D + grp.d_off + (long long)row * p.ldd;
#pragma unroll
    for (int j = 0; j < 32; ++j)
      if (gn + j < grp.N) drow[gn + j] = p.accumulate ? drow[gn + j] + v[j] : v[j];
  }
}

// Epilogue from TMEM (tf32 tier: the accumulator lived in tensor memory): this warp drains TMEM lane quarter q, columns
// [c_begin, c_end) of a tile whose first row / column are m0 / n0.
template <bool EXACT>
__device__ __forceinline__ void tc_epilogue(const TcParams& p, const GemmGroup& grp, uint32_t tmem_base, float* wtile, int q,
                                            int lane, int m0, int n0, int c_begin, int c_end, bool have_acc) {
#pragma unroll 1
  for (int c0 = c_begin; c0 < c_end; c0 += 32) {
    uint32_t r[32];
    if (have_acc) {
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0;
      MSFNO_TC_LD32(r, taddr);
      asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j) r[j] = 0u;
    }
    float v[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
    tc_epilogue_chunk<EXACT>(p, grp, v, wtile, q, lane, m0, n0, c0);
  }
}

// shared-memory matrix descriptor of k-step k (8 tf32) of an operand tile at shared address `tile`.
// K-major: advance 32 bytes inside the swizzle span; MN-major (SW128_BASE32B): 8 k-rows = two 512-byte atoms (SBO),
// 32-element M / N blocks 4096 bytes apart (LBO)
template <bool MN>
__device__ __forceinline__ uint64_t tc_desc(uint32_t tile, int k) {
  return MN ? make_smem_desc(tile + 1024 * k, TC_BK * 128, 512, 1) : make_smem_desc(tile + 32 * k, 16, 1024);
}

// TMA producer loop of the single-CTA kernels (one thread): A tile at stage offset 0, B tile at TC_A_BYTES.
// first operand pair: offsets come from the group; second pair: plain strided batch (blockIdx.y * s?2)
template <bool A_MN, bool B_MN>
__device__ __forceinline__ void tc_producer(const TcParams& p, const GemmGroup& grp, const CUtensorMap& tmA, const CUtensorMap& tmB,
                                            const CUtensorMap& tmA2, const CUtensorMap& tmB2, uint8_t* tiles, int stage_bytes,
                                            int nstages, uint64_t* full, uint64_t* empty, int m0, int n0, int nkb1, int nkb) {
  // (row, column) of the operands' first elements, computed ONCE: the 64-bit divisions cost the issuing thread ~500 clk
  // each, and four of them per k-block made this single thread the slowest stage of the whole pipeline (measured: 1490 clk
  // per k-block with every consumer starved, 8 CTAs on an idle GPU as much as 472)
  const long long a_off2 = blockIdx.y * p.sa2, b_off2 = blockIdx.y * p.sb2;
  const int ac1 = (int)(grp.a_off % p.lda), ar1 = (int)(grp.a_off / p.lda);
  const int bc1 = (int)(grp.b_off % p.ldb);
  const long long br1 = grp.b_off / p.ldb;
  const int ac2 = nkb > nkb1 ? (int)(a_off2 % p.lda2) : 0, ar2 = nkb > nkb1 ? (int)(a_off2 / p.lda2) : 0;
  const int bc2 = nkb > nkb1 ? (int)(b_off2 % p.ldb2) : 0, br2 = nkb > nkb1 ? (int)(b_off2 / p.ldb2) : 0;
  const bool b3d = p.b_group_rows > 0;
  const int tab = b3d ? (int)(br1 / p.b_group_rows) : 0;
  const int br1_in = b3d ? (int)(br1 % p.b_group_rows) : (int)br1;
  for (int kb = 0; kb < nkb; ++kb) {
    const int s = kb % nstages;
    const uint32_t ph = (uint32_t)((kb / nstages) & 1);
    const bool second = kb >= nkb1;
    const int kk = (second ? kb - nkb1 : kb) * TC_BK;
    const int ac = second ? ac2 : ac1, ar = second ? ar2 : ar1, bc = second ? bc2 : bc1, br = second ? br2 : br1_in;
    const CUtensorMap* ma = second ? &tmA2 : &tmA;
    const CUtensorMap* mb = second ? &tmB2 : &tmB;
    mbar_wait_bounded(&empty[s], ph ^ 1u);
#ifdef MSFNO_TRACE
    if (p.trace && blockIdx.x == 0 && blockIdx.y == 0 && kb >= 8 && kb < 12) p.trace[16 + kb - 8] = clock64();   // TMA issue
#endif
    mbar_arrive_expect_tx(&full[s], TC_HI_BYTES);
    uint8_t* sa = tiles + s * stage_bytes;
    if (!A_MN) {
      tma_load_2d(sa, ma, &full[s], ac + kk, ar + m0);
    } else {
      // [K][M] operand: four boxes of 32 (m) x 32 (k); box j holds m in [m0 + 32 j, +32)
#pragma unroll
      for (int j = 0; j < TC_BM / 32; ++j) tma_load_2d(sa + j * (TC_BK * 128), ma, &full[s], ac + m0 + 32 * j, ar + kk);
    }
    if (!B_MN) {
      tma_load_2d(sa + TC_A_BYTES, mb, &full[s], bc + kk, br + n0);
    } else {
      // [K][N] operand: four boxes of 32 (n) x 32 (k); box j holds n in [n0 + 32 j, +32)
      if (b3d && !second) {
#pragma unroll
        for (int j = 0; j < TC_BN / 32; ++j) tma_load_3d(sa + TC_A_BYTES + j * (TC_BK * 128), mb, &full[s], bc + n0 + 32 * j, br + kk, tab);
      } else {
#pragma unroll
        for (int j = 0; j < TC_BN / 32; ++j) tma_load_2d(sa + TC_A_BYTES + j * (TC_BK * 128), mb, &full[s], bc + n0 + 32 * j, br + kk);
      }
    }
  }
}

// =============================================================================================================
// tf32 tier: one MMA per k-step, accumulator in TMEM for the whole K loop.
template <bool A_MN, bool B_MN>
__global__ void __launch_bounds__(256, 2)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               const __grid_constant__ CUtensorMap tmA2, const __grid_constant__ CUtensorMap tmB2, TcParams p) {
  extern __shared__ uint8_t smem_raw[];
  constexpr int STAGE = TC_HI_BYTES;
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

  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;  // swizzle-128B tiles need 1024-byte alignment
  uint8_t* tiles = smem_raw + (base - smem_u32(smem_raw));
  uint64_t* bars = reinterpret_cast<uint64_t*>(tiles + TC_STAGES * STAGE);
  uint64_t* full = bars;                          // TMA bytes of the stage have landed
  uint64_t* empty = bars + TC_STAGES;             // the MMAs that read the stage have completed
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

  if (nkb > 0) {
    if (warp == 0 && lane == 0) {
      tc_producer<A_MN, B_MN>(p, grp, tmA, tmB, tmA2, tmB2, tiles, STAGE, TC_STAGES, full, empty, m0, n0, nkb1, nkb);
    } else if (warp == 1 && lane == 0) {
      // ---------------- MMA issuer ----------------
      // instruction descriptor: D=f32 (bit 4), A=B=tf32 (bits 7,10), A / B major (bits 15, 16), N>>3 at bit 17, M>>4 at bit 24
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((A_MN ? 1u : 0u) << 15) | ((B_MN ? 1u : 0u) << 16) |
                             ((uint32_t)(TC_BN >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);
      for (int kb = 0; kb < nkb; ++kb) {
        const int s = kb % TC_STAGES;
        const uint32_t ph = (uint32_t)((kb / TC_STAGES) & 1);
        mbar_wait_bounded(&full[s], ph);
        tc_fence_after();
        const uint32_t sa = base + s * STAGE;
        const uint32_t sb = sa + TC_A_BYTES;
#pragma unroll
        for (int k = 0; k < TC_BK / 8; ++k)
          tc_mma_tf32(tmem_base, tc_desc<A_MN>(sa, k), tc_desc<B_MN>(sb, k), idesc, (kb | k) ? 1u : 0u);
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
    if (nkb > 0) {
      mbar_wait_bounded(tmem_full, 0);
      tc_fence_after();
    }
    tc_epilogue<false>(p, grp, tmem_base, reinterpret_cast<float*>(tiles) + warp * (32 * 36), q, lane, m0, n0, chalf * (TC_BN / 2),
                       (chalf + 1) * (TC_BN / 2), nkb > 0);
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem_base), "r"(TC_BN));
  }
}

// =============================================================================================================
// fp32 tier: 3xTF32 with the K loop accumulated OUTSIDE the tensor core.
//
// The tensor core adds into its TMEM accumulator with truncation, not round-to-nearest: measured on B200, the relative
// error of a K = 1024 product accumulated in TMEM is 7.6e-6 (biased towards zero, growing with K) against 5.8e-7 for
// FFMA.  So every 32-wide k-block is multiplied into a FRESH accumulator (accumulate = 0 on its first MMA; two TMEM
// buffers alternate) and eight accumulator warps drain it with tcgen05.ld into fp32 registers, adding with
// round-to-nearest while the tensor core works on the next one.  Draining is the expensive part (tensor memory reads at
// 64 B/clk: 64 KB per 128 x 128 buffer = 1 024 clk, more than the 768 clk of a k-block's twelve MMAs), so
//   * the small terms lo(a) hi(b) + hi(a) lo(b) -- 2^-11 of the result, their truncation error is irrelevant -- go to a
//     third accumulator S that stays in tensor memory for the whole K loop and is drained once, and
//   * only the hi(a) hi(b) products go to the alternating buffers, TC3_GROUP k-blocks (16 MMAs = 16 truncating adds, about
//     3e-7) per drain.
// 16 warps: 0 = TMA producer, 1 and 3 = MMA issuers, 2 = TMEM allocator, 4-7 = converters, 8-15 = accumulators / epilogue.
//
// Converters (thread = tile row): a K-major A tile is read once from shared memory (one swizzled 16-byte chunk per lane
// and access: conflict-free), split and written to TENSOR MEMORY as the A operand (hi and lo, 32 columns each per
// stage), so the MMAs read only B from shared memory: 3 x 4 KB per k-step instead of 6 x 4 KB -- with 4-byte operands
// the 128 B/clk shared-memory port, not the tensor pipe, is what bounds these kernels.  An MN-major A (adjoints) keeps
// its lo() tile in shared memory.  B's lo() tile is always made in shared memory (elementwise, layout preserved).
template <bool A_MN> struct Tc3Cfg {
  static constexpr bool A_TMEM = !A_MN;
  static constexpr int STAGE_BYTES = TC_HI_BYTES + TC_B_BYTES + (A_TMEM ? 0 : TC_A_BYTES);   // A, B, lo(B) [, lo(A)]
  static constexpr int STAGES = A_TMEM ? 4 : 3;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 + 256;
  static constexpr int NH = A_TMEM ? 1 : 2;   // hi * hi accumulators (K-major A: tensor memory holds the A operand instead)
  static constexpr int TMEM_COLS = 512;       // H, S + 4 stages x (32 hi + 32 lo) columns of A   or   H0, H1, S
};
static constexpr int TC3_GROUP = 4;        // k-blocks of hi * hi products per drain of an H buffer
static constexpr int TC3_THREADS = 512;

#define MSFNO_TC_ST32(taddr, r)                                                                                          \
  asm volatile(                                                                                                          \
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "                                                                    \
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "                                         \
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};\n" ::"r"(taddr),                 \
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),      \
      "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]),        \
      "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]),        \
      "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])                                                                     \
      : "memory")

// D[tmem] (+)= A[tmem] * B[smem desc]
__device__ __forceinline__ void tc_mma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}

template <bool A_MN, bool B_MN>
__global__ void __launch_bounds__(TC3_THREADS, 1)
gemm_tc3_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                const __grid_constant__ CUtensorMap tmA2, const __grid_constant__ CUtensorMap tmB2, TcParams p) {
  extern __shared__ uint8_t smem_raw[];
  using Cfg = Tc3Cfg<A_MN>;
  constexpr int STAGE = Cfg::STAGE_BYTES, NST = Cfg::STAGES;
  constexpr bool A_TMEM = Cfg::A_TMEM;
  constexpr uint32_t LO_B = TC_HI_BYTES;                   // byte offset of lo(B) in a stage
  constexpr uint32_t LO_A = TC_HI_BYTES + TC_B_BYTES;      // byte offset of lo(A) in a stage (MN-major A only)
  GemmGroup grp;
  if (p.use_single) {
    grp = p.single;
    grp.a_off += blockIdx.y * p.sa;
    grp.b_off += blockIdx.y * p.sb;
    grp.d_off += blockIdx.y * p.sd;
  } else {
    grp = p.groups[blockIdx.y];
  }
  const int tn = blockIdx.x / p.tilesM, tm = blockIdx.x - tn * p.tilesM;
  const int m0 = tm * TC_BM, n0 = tn * TC_BN;
  if (m0 >= grp.M || n0 >= grp.N) return;  // uniform for the CTA, before any barrier / allocation
  pdl_trigger();
#ifdef MSFNO_TRACE
  const long long t_cta0 = clock64();
#endif

  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* tiles = smem_raw + (base - smem_u32(smem_raw));
  uint64_t* bars = reinterpret_cast<uint64_t*>(tiles + NST * STAGE);
  uint64_t* full = bars;                 // TMA bytes of the stage have landed
  uint64_t* empty = bars + NST;          // the MMAs that read the stage have completed
  uint64_t* conv = bars + 2 * NST;       // the lo() tiles (and the TMEM copy of A) of the stage are written
  uint64_t* acc_full = bars + 3 * NST;   // [2] the k-block product in TMEM buffer b is complete
  uint64_t* acc_empty = acc_full + 2;    // [2] the accumulator warps have read TMEM buffer b
  uint64_t* s_full = acc_empty + 2;      // the small-term accumulator S is complete
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(s_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nkb1 = (grp.K + TC_BK - 1) / TC_BK;
  const int nkb = nkb1 + (p.K2 + TC_BK - 1) / TC_BK;

  if (warp == 0 && lane == 0) {
    mbar_init(s_full, 1);
    for (int s = 0; s < NST; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 2);   // one commit per issuing thread
      mbar_init(&conv[s], 12);   // one arrival per A-converter warp (4) and per accumulator warp (8, they make lo(B))
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&acc_full[b], 1);
      mbar_init(&acc_empty[b], 8);
    }
    fence_mbar_init();
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(tmem_slot)), "r"(Cfg::TMEM_COLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  constexpr int NH = Cfg::NH;
  const uint32_t tmem_s = tmem_base + (uint32_t)(NH * TC_BN);   // S: the small terms of the whole K loop
  const uint32_t tmem_a = tmem_base + 256u;   // K-major A: stage s at columns 256 + 64 s (hi) and + 32 (lo)
  pdl_wait();

  float acc[64];   // accumulator warps: rows q*32 + lane, columns chalf*64 .. +64 of the tile
#pragma unroll
  for (int j = 0; j < 64; ++j) acc[j] = 0.f;

  if (nkb > 0) {
    if (warp == 0 && lane == 0) {
      tc_producer<A_MN, B_MN>(p, grp, tmA, tmB, tmA2, tmB2, tiles, STAGE, NST, full, empty, m0, n0, nkb1, nkb);
    } else if ((warp == 1 || warp == 3) && lane == 0) {
      // ---------------- MMA issuers ----------------
      // TWO issuing threads with disjoint accumulators: warp 1 issues the hi * hi products (into H[b]), warp 3 the small
      // terms (into S).  A single issuer is nearly synchronous with the tensor pipe (its twelve MMAs of a k-block return
      // after ~880 clk, when they have all but completed) and every mbarrier wait costs it 100-200 clk even when the phase
      // completed long ago -- three waits per k-block left the pipe idle 40 % of the time (per-k-block timeline,
      // profiles/r02_x3_timeline_*.txt).  Now each thread waits once per k-block (conv[s] implies full[s]), and the
      // other thread's MMAs run under that wait.  Both commit on empty[s] (count 2).
      const bool issue_h = warp == 1;
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((A_MN ? 1u : 0u) << 15) | ((B_MN ? 1u : 0u) << 16) |
                             ((uint32_t)(TC_BN >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);
#ifdef MSFNO_TRACE
      long long tr_conv = 0, tr_acc = 0, tr_first = 0, tr_last = 0, tr_issue = 0;
#endif
      for (int kb = 0; kb < nkb; ++kb) {
        const int s = kb % NST, g = kb / TC3_GROUP, b = g % NH;
        const bool first = kb % TC3_GROUP == 0, last = kb % TC3_GROUP == TC3_GROUP - 1 || kb == nkb - 1;
#ifdef MSFNO_TRACE
        const long long w0 = clock64();
#endif
        mbar_wait_bounded(&conv[s], (uint32_t)((kb / NST) & 1));   // every converter saw full[s] before it arrived here
#ifdef MSFNO_TRACE
        const long long w2 = clock64();
#endif
        if (issue_h && first) mbar_wait_bounded(&acc_empty[b], (uint32_t)(((g / NH) & 1) ^ 1));
#ifdef MSFNO_TRACE
        const long long w3 = clock64();
        tr_conv += w2 - w0; tr_acc += w3 - w2;
        if (kb == 0) tr_first = w3;
        tr_last = w3;
#endif
        tc_fence_after();
        const uint32_t sa = base + s * STAGE, sb = sa + TC_A_BYTES;
        const uint32_t ta = tmem_a + (uint32_t)(s * 64);
        if (issue_h) {
          const uint32_t d = tmem_base + (uint32_t)(b * TC_BN);
#pragma unroll
          for (int k = 0; k < TC_BK / 8; ++k) {   // hi(a) hi(b): the tensor core truncates the raw fp32 operands itself
            if (A_TMEM) tc_mma_tf32_ts(d, ta + 8u * k, tc_desc<B_MN>(sb, k), idesc, (first && k == 0) ? 0u : 1u);
            else tc_mma_tf32(d, tc_desc<A_MN>(sa, k), tc_desc<B_MN>(sb, k), idesc, (first && k == 0) ? 0u : 1u);
          }
          tc_commit(&empty[s]);                 // (with the other issuer's commit) the stage may be refilled
          if (last) tc_commit(&acc_full[b]);    // the group's hi * hi product is complete
        } else {
          // small terms: they add up among themselves at 2^-11 of the magnitude, S stays in tensor memory for the whole K loop
#pragma unroll
          for (int k = 0; k < TC_BK / 8; ++k) {   // lo(a) hi(b): lo(A) from tensor memory (K-major A) or shared memory
            if (A_TMEM) tc_mma_tf32_ts(tmem_s, ta + 32u + 8u * k, tc_desc<B_MN>(sb, k), idesc, (kb | k) ? 1u : 0u);
            else tc_mma_tf32(tmem_s, tc_desc<A_MN>(sa + LO_A, k), tc_desc<B_MN>(sb, k), idesc, (kb | k) ? 1u : 0u);
          }
#pragma unroll
          for (int k = 0; k < TC_BK / 8; ++k) {   // hi(a) lo(b)
            if (A_TMEM) tc_mma_tf32_ts(tmem_s, ta + 8u * k, tc_desc<B_MN>(sa + LO_B, k), idesc, 1u);
            else tc_mma_tf32(tmem_s, tc_desc<A_MN>(sa, k), tc_desc<B_MN>(sa + LO_B, k), idesc, 1u);
          }
          tc_commit(&empty[s]);
          if (kb == nkb - 1) tc_commit(s_full);   // S is complete
        }
#ifdef MSFNO_TRACE
        const long long w4 = clock64();
        tr_issue += w4 - w3;
        if (issue_h && p.trace && blockIdx.x == 0 && blockIdx.y == 0 && kb >= 8 && kb < 12) {
          long long* e = p.trace + 32 + (kb - 8) * 12;
          e[0] = w0; e[1] = w0; e[2] = w2; e[3] = w3; e[4] = w4;
        }
        if (!issue_h && p.trace && blockIdx.x == 0 && blockIdx.y == 0 && kb >= 8 && kb < 12) {
          long long* e = p.trace + 32 + (kb - 8) * 12;
          e[10] = w0; e[11] = w4;
        }
#endif
      }
#ifdef MSFNO_TRACE
      if (issue_h && p.trace && blockIdx.x == 0 && blockIdx.y == 0) {
        p.trace[0] = 0; p.trace[1] = tr_conv; p.trace[2] = tr_acc; p.trace[3] = tr_first; p.trace[4] = tr_last; p.trace[8] = tr_issue;
      }
#endif
    } else if (warp >= 4 && warp < 8) {
      // ---------------- A converters (thread = tile row of A = TMEM lane: warp 4 + q owns lane quarter q) ----------------
      const int t = threadIdx.x - 128;
#ifdef MSFNO_TRACE
      long long cv_full = 0, cv_slot = 0, cv_work = 0;
#endif
      for (int kb = 0; kb < nkb; ++kb) {
        const int s = kb % NST;
#ifdef MSFNO_TRACE
        const long long c0 = clock64();
#endif
        mbar_wait_bounded(&full[s], (uint32_t)((kb / NST) & 1));
#ifdef MSFNO_TRACE
        const long long c1 = clock64();
        cv_full += c1 - c0;
        if (p.trace && blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 128 && kb >= 8 && kb < 12) p.trace[20 + kb - 8] = c1;   // landed
#endif
        uint8_t* st = tiles + s * STAGE;
        if (A_TMEM) {
          // row t of the K-major A tile: 128 bytes, its 16-byte chunk c stored at chunk position c ^ (t % 8).  Both halves go
          // to tensor memory (slot = stage: free whenever the stage is), so every MMA reads only B from shared memory: with
          // 4-byte operands an MMA with both operands in shared memory reads 8 KB in its 64 clk -- all of the 128 B/clk
          // port (measured: 1 300 clk per k-block with hi(A) left in shared memory, and writing lo(A) back to shared memory
          // costs the converter ~1 100 clk per k-block, half of it the generic -> async proxy fence).
          uint32_t hi[32], lo[32];
          const uint8_t* rowp = st + t * 128;
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            const float4 v = *reinterpret_cast<const float4*>(rowp + ((c ^ (t & 7)) << 4));
            hi[4 * c] = __float_as_uint(v.x); hi[4 * c + 1] = __float_as_uint(v.y);
            hi[4 * c + 2] = __float_as_uint(v.z); hi[4 * c + 3] = __float_as_uint(v.w);
            lo[4 * c] = __float_as_uint(tf32_lo(v.x)); lo[4 * c + 1] = __float_as_uint(tf32_lo(v.y));
            lo[4 * c + 2] = __float_as_uint(tf32_lo(v.z)); lo[4 * c + 3] = __float_as_uint(tf32_lo(v.w));
          }
          const uint32_t ta = tmem_a + (uint32_t)(s * 64) + ((uint32_t)((warp & 3) * 32) << 16);
          MSFNO_TC_ST32(ta, hi);
          MSFNO_TC_ST32(ta + 32u, lo);
          asm volatile("tcgen05.wait::st.sync.aligned;\n" ::: "memory");
          tc_fence_before();
        } else {
          convert_stage_lo(st, st + LO_A, TC_A_BYTES, t, TC_NCONV);
          fence_proxy_async();   // generic-proxy writes -> visible to the tensor core's (async proxy) reads
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&conv[s]);
#ifdef MSFNO_TRACE
        const long long c3 = clock64();
        cv_work += c3 - c1;
        if (p.trace && blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 128 && kb >= 8 && kb < 12) {
          long long* e = p.trace + 32 + (kb - 8) * 12;
          e[5] = c1; e[6] = c3;
        }
#endif
      }
#ifdef MSFNO_TRACE
      if (p.trace && blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 128) { p.trace[9] = cv_full; p.trace[10] = cv_slot; p.trace[11] = cv_work; }
#endif
    } else if (warp >= 8) {
      // ---------------- accumulators: acc += TMEM buffer (round-to-nearest adds on the CUDA cores) ----------------
      // They also make lo(B) in shared memory, AHEAD k-blocks ahead of the product they drain: the A converters alone
      // (one warp per scheduler, a serial chain of shared loads, ALU, tensor-memory stores and fences per k-block) were
      // the slowest stage of the pipeline.
      const int q = warp & 3, chalf = (warp - 8) >> 2;
      const int t = threadIdx.x - 256;
      const uint32_t lane_off = (uint32_t)(q * 32) << 16;
#ifdef MSFNO_TRACE
      long long ac_full = 0, ac_conv = 0, ac_wait = 0, ac_drain = 0;
#endif
      auto convert_b = [&](int kb) {
        const int s = kb % NST;
#ifdef MSFNO_TRACE
        const long long q0 = clock64();
#endif
        mbar_wait_bounded(&full[s], (uint32_t)((kb / NST) & 1));
#ifdef MSFNO_TRACE
        const long long q1 = clock64();
        ac_full += q1 - q0;
#endif
        uint8_t* st = tiles + s * STAGE;
        convert_stage_lo(st + TC_A_BYTES, st + LO_B, TC_B_BYTES, t, 256);
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(&conv[s]);
#ifdef MSFNO_TRACE
        const long long q4 = clock64();
        ac_conv += q4 - q1;
        if (p.trace && blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 256 && kb >= 8 && kb < 12) {
          long long* e = p.trace + 32 + (kb - 8) * 12;
          e[7] = q0; e[8] = q1; e[9] = q4;
        }
#endif
      };
      constexpr int AHEAD = NST - 1;   // lo(B) is made as soon as its stage lands: AHEAD k-blocks before its product is needed
#pragma unroll
      for (int j = 0; j < AHEAD; ++j)
        if (j < nkb) convert_b(j);
      auto drain = [&](uint32_t src, uint64_t* release) {
        uint32_t r[32];
        MSFNO_TC_LD32(r, src);
        asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
#pragma unroll
        for (int j = 0; j < 32; ++j) acc[j] += __uint_as_float(r[j]);
        MSFNO_TC_LD32(r, src + 32u);
        asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
        tc_fence_before();
        if (release && lane == 0) mbar_arrive(release);   // the tensor core may overwrite the buffer
#pragma unroll
        for (int j = 0; j < 32; ++j) acc[32 + j] += __uint_as_float(r[j]);
      };
      for (int kb = 0; kb < nkb; ++kb) {
        if (kb % TC3_GROUP != TC3_GROUP - 1 && kb != nkb - 1) {
          if (kb + AHEAD < nkb) convert_b(kb + AHEAD);
          continue;
        }
        // group end: the drain first -- the hi * hi issuer waits for it, and the stage of k-block kb + AHEAD can only
        // be refilled once the MMAs of k-block kb - 1 have completed, so that conversion would block here for a TMA
        // round trip (measured: acc_empty 1 850 clk late)
        const int g = kb / TC3_GROUP, b = g % NH;
#ifdef MSFNO_TRACE
        const long long q2 = clock64();
#endif
        mbar_wait_bounded(&acc_full[b], (uint32_t)((g / NH) & 1));
#ifdef MSFNO_TRACE
        const long long q3 = clock64();
        ac_wait += q3 - q2;
#endif
        tc_fence_after();
        drain(tmem_base + lane_off + (uint32_t)(b * TC_BN + chalf * 64), &acc_empty[b]);
#ifdef MSFNO_TRACE
        ac_drain += clock64() - q3;
#endif
        if (kb + AHEAD < nkb) convert_b(kb + AHEAD);
      }
#ifdef MSFNO_TRACE
      if (p.trace && blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 256) { p.trace[12] = ac_full; p.trace[13] = ac_conv; p.trace[14] = ac_wait; p.trace[15] = ac_drain; }
#endif
      mbar_wait_bounded(s_full, 0u);
      tc_fence_after();
      drain(tmem_s + lane_off + (uint32_t)(chalf * 64), nullptr);
    }
  }

#ifdef MSFNO_TRACE
  if (p.trace && blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 256) p.trace[5] = clock64();
#endif
  if (warp >= 8) {
    // ---------------- epilogue from the register accumulators ----------------
    // (all MMAs have completed and every converter has finished: the last acc_full was observed above)
    const int q = warp & 3, chalf = (warp - 8) >> 2;
    float* wtile = reinterpret_cast<float*>(tiles) + (warp - 8) * (32 * 36);
    float v[32];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = acc[32 * h + j];
      tc_epilogue_chunk<true>(p, grp, v, wtile, q, lane, m0, n0, chalf * 64 + 32 * h);
    }
  }
  tc_fence_before();
  __syncthreads();
#ifdef MSFNO_TRACE
  if (p.trace && blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 256) { p.trace[6] = clock64(); p.trace[7] = t_cta0; }
#endif
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem_base), "r"(Cfg::TMEM_COLS));
  }
}

// ---------------------------------------------------------------------------------------------
// CTA-pair variant (cta_group::2), tf32 tier: two CTAs of a cluster (the two SMs of a TPC) compute one 256 x 256 tile.
// Each CTA stages ITS 128 rows of A and ITS 128 rows of B (= 128 of the 256 output columns); the leader's single
// thread issues tcgen05.mma.cta_group::2 (M = 256, N = 256), which reads both CTAs' shared memory and writes each
// CTA's 128 x 256 accumulator slice into that CTA's TMEM.  With 4-byte operands a single-CTA 128 x 128 tile needs
// 128 B/clk of smem reads for the MMA plus 128 B/clk of TMA writes, twice what an SM's shared memory delivers; the
// pair halves both (each operand byte is staged once per pair), which is what lets the TF32 MLP GEMMs leave the
// 40 % plateau.  Barriers: full[s] lives in the leader (both producers' TMA bytes complete on it), empty[s] and
// tmem_full exist in both CTAs and are signalled by multicast commits.  K-major operands, single operand pair only.
static constexpr int TC2_BN = 256;
static constexpr int TC2_STAGE_BYTES = TC_A_BYTES + TC_A_BYTES;   // 128 rows of A + 128 rows of B per CTA
static constexpr int TC2_SMEM_BYTES = TC_STAGES * TC2_STAGE_BYTES + 1024 + 256;

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

  if (nkb > 0) {
    if (warp == 0 && lane == 0) {
      // ---------------- TMA producer (both CTAs) ----------------
      const int ac = (int)(grp.a_off % p.lda), ar = (int)(grp.a_off / p.lda) + m0;
      const int bc = (int)(grp.b_off % p.ldb), br = (int)(grp.b_off / p.ldb) + n0 + (int)rank * TC_BM;
      for (int kb = 0; kb < nkb; ++kb) {
        const int s = kb % TC_STAGES;
        const uint32_t ph = (uint32_t)((kb / TC_STAGES) & 1);
        const int kk = kb * TC_BK;
        mbar_wait_bounded(&empty[s], ph ^ 1u);
        if (rank == 0) mbar_arrive_expect_tx(&full[s], 2 * TC2_STAGE_BYTES);   // bytes of both CTAs
        uint8_t* sa = tiles + s * TC2_STAGE_BYTES;
        tma_load_2d_pair(sa, &tmA, &full[s], ac + kk, ar);
        tma_load_2d_pair(sa + TC_A_BYTES, &tmB, &full[s], bc + kk, br);
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
    if (nkb > 0) {
      mbar_wait_bounded(tmem_full, 0);
      tc_fence_after();
    }
    tc_epilogue<false>(p, grp, tmem_base, reinterpret_cast<float*>(tiles) + warp * (32 * 36), q, lane, m0, n0, chalf * (TC2_BN / 2),
                       (chalf + 1) * (TC2_BN / 2), nkb > 0);
    tc_fence_before();
  }
  cluster_sync_all();   // neither CTA may retire (or free TMEM) while its peer can still touch it
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;\n" ::"r"(tmem_base), "r"(TC2_BN));
  }
}

// ---------------------------------------------------------------------------------------------
static bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

bool gemm_tc_supported(const GemmLaunch& g) {
  if ((g.lda & 3) || (g.ldb & 3) || !aligned16(g.A) || !aligned16(g.B)) return false;
  if (g.A2 && (!g.a_kmajor || (g.lda2 & 3) || (g.ldb2 & 3) || !aligned16(g.A2) || !aligned16(g.B2))) return false;
  if (g.mask && (g.ldmask & 1)) return false;
  return get_encode() != nullptr;
}

// 3-D fp32 tensor [tables][rows][ld] (cols valid), box = 32 columns x 32 rows x 1 table, 128-byte swizzle with 32-byte atoms
static int make_map_3d(CUtensorMap* tm, const float* base, long long tables, long long rows, long long cols, long long ld) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return record_error(MSFNO_ERR_CUDA, "cuTensorMapEncodeTiled entry point unavailable");
  cuuint64_t dims[3] = {(cuuint64_t)cols, (cuuint64_t)rows, (cuuint64_t)tables};
  cuuint64_t strides[2] = {(cuuint64_t)ld * 4, (cuuint64_t)ld * 4 * (cuuint64_t)rows};
  cuuint32_t box[3] = {(cuuint32_t)TC_BK, (cuuint32_t)TC_BK, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return record_error(MSFNO_ERR_CUDA, "cuTensorMapEncodeTiled failed (3-D table map)");
  return MSFNO_OK;
}

template <typename K>
static cudaError_t opt_in_smem(K kern, int bytes) {
  return cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
}

// a_rows / a_cols (b_rows / b_cols): rows and valid columns of the underlying 2-D buffers as they lie in memory (TMA
// zero-fills beyond them).  For an MN-major operand the buffer is [K rows][M or N cols].  g.b_group_rows > 0 (MN-major
// B only): the buffer is b_rows / b_group_rows stacked tables of b_group_rows rows each.
int launch_gemm_tc(const GemmLaunch& g, long long a_rows, long long a_cols, long long b_rows, long long b_cols,
                   int round_tf32, cudaStream_t st, long long a2_rows, long long a2_cols, long long b2_rows,
                   long long b2_cols) {
  if (g.ngroups <= 0 || g.maxM <= 0 || g.maxN <= 0) return MSFNO_OK;
  const bool amn = !g.a_kmajor, bmn = !g.b_kmajor;
  const bool x3 = g.x3 != 0;
  TcParams p{};
  p.D = g.D; p.lda = g.lda; p.ldb = g.ldb; p.ldd = g.ldd;
  p.groups = g.groups; p.single = g.single; p.sa = g.sa; p.sb = g.sb; p.sd = g.sd; p.use_single = g.use_single;
  p.relu_even = g.relu_even; p.round_tf32 = x3 ? 0 : round_tf32;
  p.bias = g.bias; p.sbias = g.sbias; p.add = g.add; p.ldadd = g.ldadd; p.sadd = g.sadd; p.act_gelu = g.act_gelu;
  p.mask = g.mask; p.ldmask = g.ldmask; p.accumulate = g.accumulate;
  p.b_group_rows = (bmn && g.b_group_rows > 0) ? g.b_group_rows : 0;

  if (x3) {
    // ---- fp32 tier: 3xTF32, K loop accumulated in registers (gemm_tc3_kernel) ----
    CUtensorMap tmA, tmB, tmA2, tmB2;
    int rc = make_map(&tmA, g.A, a_rows, a_cols, g.lda, amn ? TC_BK : TC_BM, amn);
    if (rc) return rc;
    if (p.b_group_rows > 0) rc = make_map_3d(&tmB, g.B, b_rows / p.b_group_rows, p.b_group_rows, b_cols, g.ldb);
    else rc = make_map(&tmB, g.B, b_rows, b_cols, g.ldb, bmn ? TC_BK : TC_BN, bmn);
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
    static PerDeviceOnce once3;
    MSFNO_CUDA_OK(once3.run([] {
      cudaError_t e = cudaSuccess;
      auto set = [&](cudaError_t r) { if (e == cudaSuccess) e = r; };
      set(opt_in_smem(gemm_tc3_kernel<false, false>, Tc3Cfg<false>::SMEM_BYTES));
      set(opt_in_smem(gemm_tc3_kernel<false, true>, Tc3Cfg<false>::SMEM_BYTES));
      set(opt_in_smem(gemm_tc3_kernel<true, false>, Tc3Cfg<true>::SMEM_BYTES));
      set(opt_in_smem(gemm_tc3_kernel<true, true>, Tc3Cfg<true>::SMEM_BYTES));
      return e;
    }));
    p.lda2 = g.A2 ? g.lda2 : 4; p.ldb2 = g.A2 ? g.ldb2 : 4; p.sa2 = g.sa2; p.sb2 = g.sb2; p.K2 = g.A2 ? g.K2 : 0;
    p.tilesM = (g.maxM + TC_BM - 1) / TC_BM;
    p.tilesN = (g.maxN + TC_BN - 1) / TC_BN;
    dim3 grid(p.tilesM * p.tilesN, g.ngroups);
#ifdef MSFNO_TRACE
    static long long* d_trace = nullptr;
    if (!d_trace) MSFNO_CUDA_OK(cudaMalloc(&d_trace, 96 * sizeof(long long)));
    MSFNO_CUDA_OK(cudaMemsetAsync(d_trace, 0, 96 * sizeof(long long), st));
    p.trace = d_trace;
#endif
    if (amn && bmn) MSFNO_CUDA_OK(launch_pdl(gemm_tc3_kernel<true, true>, grid, dim3(TC3_THREADS), Tc3Cfg<true>::SMEM_BYTES, st, tmA, tmB, tmA2, tmB2, p));
    else if (amn) MSFNO_CUDA_OK(launch_pdl(gemm_tc3_kernel<true, false>, grid, dim3(TC3_THREADS), Tc3Cfg<true>::SMEM_BYTES, st, tmA, tmB, tmA2, tmB2, p));
    else if (bmn) MSFNO_CUDA_OK(launch_pdl(gemm_tc3_kernel<false, true>, grid, dim3(TC3_THREADS), Tc3Cfg<false>::SMEM_BYTES, st, tmA, tmB, tmA2, tmB2, p));
    else MSFNO_CUDA_OK(launch_pdl(gemm_tc3_kernel<false, false>, grid, dim3(TC3_THREADS), Tc3Cfg<false>::SMEM_BYTES, st, tmA, tmB, tmA2, tmB2, p));
#ifdef MSFNO_TRACE
    {
      long long h[96];
      MSFNO_CUDA_OK(cudaStreamSynchronize(st));
      MSFNO_CUDA_OK(cudaMemcpy(h, d_trace, sizeof(h), cudaMemcpyDeviceToHost));
      fprintf(stderr, "gemm_tc3 trace CTA(0,0) maxM=%d maxN=%d: issuer waits full %lld conv %lld acc_empty %lld | first MMA at %lld, last issue at %lld, "
                      "mainloop done %lld, cta done %lld (clk since CTA start); issue blocks %lld | A converter: waits full %lld slot %lld, busy %lld | accumulators: "
                      "waits full %lld, lo(B) %lld, waits product %lld, drains %lld | TMA issue -> landed, k-blocks 8..11: %lld %lld %lld %lld clk, issued at %lld %lld %lld %lld\n", g.maxM, g.maxN, h[0], h[1], h[2], h[3] - h[7], h[4] - h[7],
              h[5] - h[7], h[6] - h[7], h[8], h[9], h[10], h[11], h[12], h[13], h[14], h[15], h[20] - h[16], h[21] - h[17], h[22] - h[18], h[23] - h[19], h[16] - h[7], h[17] - h[7], h[18] - h[7], h[19] - h[7]);
      for (int j = 0; j < 4; ++j) {
        const long long* e = h + 32 + 12 * j;
        fprintf(stderr, "  k-block %d: TMA issued %lld | issuer arrives %lld, full %lld, conv %lld, acc %lld, issued %lld | S issuer arrives %lld issued %lld | A conv: landed %lld done %lld | lo(B) (warp 8): "
                        "starts waiting %lld landed %lld done %lld\n", 8 + j, h[16 + j] - h[7], e[0] - h[7], e[1] - h[7], e[2] - h[7], e[3] - h[7], e[4] - h[7], e[10] - h[7], e[11] - h[7], e[5] - h[7],
                e[6] - h[7], e[7] - h[7], e[8] - h[7], e[9] - h[7]);
      }
    }
#endif
    count_launch();
    MSFNO_CUDA_OK(cudaGetLastError());
    return MSFNO_OK;
  }

  // CTA pair (tf32 tier): 256-column tiles; with N = 512 only 60 pairs exist for the 7440-row MLP (< 74 TPCs) and the
  // single-CTA kernel's finer tiles win (tools/gemm_bench.py)
  if (!amn && !bmn && !g.A2 && !g.add && !g.mask && !g.accumulate && g.maxM >= 1024 && g.maxN >= 768) {
    CUtensorMap tmA, tmB;
    int rc = make_map(&tmA, g.A, a_rows, a_cols, g.lda, TC_BM);
    if (rc) return rc;
    rc = make_map(&tmB, g.B, b_rows, b_cols, g.ldb, TC_BM);
    if (rc) return rc;
    static PerDeviceOnce once2;
    MSFNO_CUDA_OK(once2.run([] { return opt_in_smem(gemm_tc2_kernel, TC2_SMEM_BYTES); }));
    p.tilesM = (g.maxM + 255) / 256;
    p.tilesN = (g.maxN + TC2_BN - 1) / TC2_BN;
    dim3 grid(2 * p.tilesM * p.tilesN, g.ngroups);
    MSFNO_CUDA_OK(launch_pdl(gemm_tc2_kernel, grid, dim3(256), TC2_SMEM_BYTES, st, tmA, tmB, p));
    count_launch();
    MSFNO_CUDA_OK(cudaGetLastError());
    return MSFNO_OK;
  }

  CUtensorMap tmA, tmB, tmA2, tmB2;
  int rc = make_map(&tmA, g.A, a_rows, a_cols, g.lda, amn ? TC_BK : TC_BM, amn);
  if (rc) return rc;
  if (p.b_group_rows > 0) rc = make_map_3d(&tmB, g.B, b_rows / p.b_group_rows, p.b_group_rows, b_cols, g.ldb);
  else rc = make_map(&tmB, g.B, b_rows, b_cols, g.ldb, bmn ? TC_BK : TC_BN, bmn);
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
  static PerDeviceOnce once;
  MSFNO_CUDA_OK(once.run([] {
    cudaError_t e = cudaSuccess;
    auto set = [&](cudaError_t r) { if (e == cudaSuccess) e = r; };
    set(opt_in_smem(gemm_tc_kernel<false, false>, TC_SMEM_BYTES));
    set(opt_in_smem(gemm_tc_kernel<false, true>, TC_SMEM_BYTES));
    set(opt_in_smem(gemm_tc_kernel<true, false>, TC_SMEM_BYTES));
    set(opt_in_smem(gemm_tc_kernel<true, true>, TC_SMEM_BYTES));
    return e;
  }));
  p.lda2 = g.A2 ? g.lda2 : 4; p.ldb2 = g.A2 ? g.ldb2 : 4; p.sa2 = g.sa2; p.sb2 = g.sb2; p.K2 = g.A2 ? g.K2 : 0;
  p.tilesM = (g.maxM + TC_BM - 1) / TC_BM;
  p.tilesN = (g.maxN + TC_BN - 1) / TC_BN;
  dim3 grid(p.tilesM * p.tilesN, g.ngroups);
  if (amn && bmn) MSFNO_CUDA_OK(launch_pdl(gemm_tc_kernel<true, true>, grid, dim3(256), TC_SMEM_BYTES, st, tmA, tmB, tmA2, tmB2, p));
  else if (amn) MSFNO_CUDA_OK(launch_pdl(gemm_tc_kernel<true, false>, grid, dim3(256), TC_SMEM_BYTES, st, tmA, tmB, tmA2, tmB2, p));
  else if (bmn) MSFNO_CUDA_OK(launch_pdl(gemm_tc_kernel<false, true>, grid, dim3(256), TC_SMEM_BYTES, st, tmA, tmB, tmA2, tmB2, p));
  else MSFNO_CUDA_OK(launch_pdl(gemm_tc_kernel<false, false>, grid, dim3(256), TC_SMEM_BYTES, st, tmA, tmB, tmA2, tmB2, p));
  count_launch();
  MSFNO_CUDA_OK(cudaGetLastError());
  return MSFNO_OK;
}

}  // namespace msfno
