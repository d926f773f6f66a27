// Persistent, weight-stationary TF32 tensor-core kernel for the 1x1 convolutions in NCHW (tensor-core tier of
// msfno_conv1x1_fwd).  A CTA owns 128 output channels and a strided set of 128-pixel tiles:
//   * the weight slab [128 x K] (both operand pairs of the big-skip form) is loaded ONCE by TMA and stays in smem;
//   * activations stream through a small TMA ring as MN-major operands (SWIZZLE_128B_BASE32B, the only layout
//     tcgen05 accepts for MN-major 32-bit data);
//   * two 128-column TMEM accumulators alternate, so the epilogue (bias, GELU (gelu_tanh3), residual / pos-embed add,
//     128-byte row stores) of tile i overlaps the MMAs of tile i+1.
// SM fill traffic per launch drops from (weights + activations) per tile to activations only.
//
// replaces: nn.Conv2d(.., 1) + bias + nn.GELU + residual adds + torch.cat in the reference's channel MLPs
//   (/root/reference MSFNO/Models/sfno/layers.py:161-168; sfnonet.py:232,249,671,682-684).
//
// Operand roles: the ACTIVATION tile is the MN-major A operand (M = 128 pixels -> TMEM lanes), the resident weight slab
// is the K-major B operand (N = 128 output channels -> TMEM columns).  With pixels on the lanes every epilogue warp
// instruction reads / writes 32 CONSECUTIVE pixels of ONE channel: a coalesced 128-byte access that touches one page.
// (Channels-on-lanes made each lane own a row; NCHW rows are HW*4 bytes = 4 MB apart, so one STG touched 32 pages and
// ran at 0.6 TB/s, a staged TMA store reached 1.2 TB/s and the strided `add` loads doubled the kernel time --
// profiles/r01_conv_experiments_*.json.)
//
// Warp roles (576 threads): warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer, warps 2-17 = epilogue
// (TMEM lane quarter w % 4 = 32-pixel block, 32-channel chunk (w - 2) / 4).
#include "plan.h"
#include "tc_common.cuh"

namespace msfno {

static constexpr int CT_BM = 128, CT_BN = 128;
static constexpr int CT_KB_BYTES = CT_BM * TC_BK * 4;  // one 128 x 32 fp32 operand block = 16 KB
static constexpr int CT_MAX_KB = 8;                     // resident weight k-blocks of the first pair (K1 <= 256)
static constexpr int CT_MAX_KB2 = 32;                   // streamed weight k-blocks (second pair, or a single pair too big to stay resident)

struct ConvTcParams {
  float* D;
  long long lda, lda2, ldb, ldb2, ldd;
  long long sa, sb, sb2, sd;   // per-sample strides (floats): weights, x, x2, y
  const float* bias; long long sbias;
  const float* add; long long ldadd, sadd;
  int M, N, K1, K2;
  int act_gelu;
  int nstages;   // ring depth for the streamed activation blocks
  int tilesN;
  int slot_bytes;  // ring slot: 16 KB (activations) or 32 KB (second-pair weights + activations)
  int round_tf32;  // round outputs to TF32 (they feed another tensor-core GEMM)
};

__global__ void __launch_bounds__(576, 1)
conv_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               const __grid_constant__ CUtensorMap tmA2, const __grid_constant__ CUtensorMap tmB2, ConvTcParams p) {
  extern __shared__ uint8_t smem_raw[];
  pdl_trigger();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.y * CT_BM;
  const int b = blockIdx.z;
  const int nkb1 = (p.K1 + TC_BK - 1) / TC_BK, nkb2 = (p.K2 + TC_BK - 1) / TC_BK, nkb = nkb1 + nkb2;
  const int NS = p.nstages;

  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* tiles = smem_raw + (base - smem_u32(smem_raw));
  uint8_t* a_res = tiles;                                  // nkb1 resident weight blocks (first operand pair)
  uint8_t* ring = tiles + (size_t)nkb1 * CT_KB_BYTES;      // NS slots: [optional second-pair weight block][activation block]
  const int SLOT = p.slot_bytes, BOFF = SLOT - CT_KB_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(ring + (size_t)NS * SLOT);
  uint64_t* full = bars;               // [NS]
  uint64_t* empty = bars + 8;          // [NS]
  uint64_t* a_full = bars + 16;
  uint64_t* tmem_full = bars + 17;     // [2]
  uint64_t* tmem_empty = bars + 19;    // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 21);
  float* bias_s = reinterpret_cast<float*>(ring + (size_t)NS * SLOT + 512);   // [128] bias of this CTA's channels

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < NS; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    mbar_init(a_full, 1);
    for (int a = 0; a < 2; ++a) { mbar_init(&tmem_full[a], 1); mbar_init(&tmem_empty[a], 16); }
    fence_mbar_init();
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(tmem_slot)), "r"(2 * CT_BN));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();   // everything above overlapped the previous kernel's tail
  if (threadIdx.x >= 64 && threadIdx.x < 64 + CT_BM) {
    const int ch = m0 + (int)threadIdx.x - 64;
    bias_s[threadIdx.x - 64] = (p.bias && ch < p.M) ? p.bias[(long long)b * p.sbias + ch] : 0.0f;
  }
  __syncthreads();

  if (warp == 0) {
    if (lane == 0) {
      // ---------------- TMA producer ----------------
      const long long aoff = (long long)b * p.sa;
      mbar_arrive_expect_tx(a_full, (uint32_t)(nkb1 * CT_KB_BYTES));
      for (int kb = 0; kb < nkb1; ++kb)
        tma_load_2d(a_res + (size_t)kb * CT_KB_BYTES, &tmA, a_full, (int)(aoff % p.lda) + kb * TC_BK, (int)(aoff / p.lda) + m0);
      const int krow1 = (int)(((long long)b * p.sb) / p.ldb);
      const int krow2 = p.K2 ? (int)(((long long)b * p.sb2) / p.ldb2) : 0;
      uint32_t kc = 0;
      for (int t = blockIdx.x; t < p.tilesN; t += gridDim.x) {
        const int n0 = t * CT_BN;
        for (int kb = 0; kb < nkb; ++kb, ++kc) {
          const int s = kc % NS;
          const uint32_t ph = (kc / NS) & 1u;
          const bool second = kb >= nkb1;
          const int krow = second ? krow2 + (kb - nkb1) * TC_BK : krow1 + kb * TC_BK;
          const CUtensorMap* mb = second ? &tmB2 : &tmB;
          mbar_wait_bounded(&empty[s], ph ^ 1u);
          mbar_arrive_expect_tx(&full[s], second ? 2 * CT_KB_BYTES : CT_KB_BYTES);
          uint8_t* slot = ring + (size_t)s * SLOT;
          if (second) tma_load_2d(slot, &tmA2, &full[s], (kb - nkb1) * TC_BK, m0);   // streamed weight block of pair 2
          uint8_t* dst = slot + BOFF;
#pragma unroll
          for (int j = 0; j < CT_BN / 32; ++j) tma_load_2d(dst + j * (TC_BK * 128), mb, &full[s], n0 + 32 * j, krow);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // ---------------- MMA issuer ----------------
      // D=f32, A=B=tf32, A MN-major (bit 15: activations, pixels contiguous), B K-major (weights), N = 128 channels
      // at bit 17, M = 128 pixels at bit 24
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | ((uint32_t)(CT_BM >> 3) << 17) |
                             ((uint32_t)(CT_BN >> 4) << 24);
      mbar_wait_bounded(a_full, 0);
      tc_fence_after();
      uint32_t kc = 0, it = 0;
      const uint32_t a_addr = base, r_addr = base + (uint32_t)nkb1 * CT_KB_BYTES;
      for (int t = blockIdx.x; t < p.tilesN; t += gridDim.x, ++it) {
        const uint32_t acc = it & 1u, use = it >> 1;
        mbar_wait_bounded(&tmem_empty[acc], (use & 1u) ^ 1u);
        tc_fence_after();
        for (int kb = 0; kb < nkb; ++kb, ++kc) {
          const int s = kc % NS;
          const uint32_t ph = (kc / NS) & 1u;
          mbar_wait_bounded(&full[s], ph);
          tc_fence_after();
          const uint32_t slot = r_addr + (uint32_t)s * (uint32_t)SLOT;
          const uint32_t sa = (kb < nkb1) ? a_addr + (uint32_t)kb * CT_KB_BYTES : slot, sb = slot + (uint32_t)BOFF;
#pragma unroll
          for (int k = 0; k < TC_BK / 8; ++k)
            tc_mma_tf32(tmem_base + acc * CT_BN, make_smem_desc(sb + 1024 * k, TC_BK * 128, 512, 1),
                        make_smem_desc(sa + 32 * k, 16, 1024), idesc, (kb | k) ? 1u : 0u);
          tc_commit(&empty[s]);
        }
        tc_commit(&tmem_full[acc]);
      }
    }
  } else {
    // ---------------- epilogue warps 2..17: lane = pixel (TMEM lane quarter w % 4), 32-channel chunk (w - 2) / 4 ------
    const int q = warp & 3, cq = (warp - 2) >> 2;
    const int c0 = cq * 32;                                  // first channel (accumulator column) of this warp
    const int nch = min(32, p.M - m0 - c0);                  // valid channels of the chunk (<= 0: nothing to store)
    uint32_t it = 0;
    for (int t = blockIdx.x; t < p.tilesN; t += gridDim.x, ++it) {
      const uint32_t acc = it & 1u, use = it >> 1;
      const int pix = t * CT_BN + q * 32 + lane;
      const bool pix_ok = pix < p.N;
      float* dptr = p.D + (long long)b * p.sd + (long long)(m0 + c0) * p.ldd + pix;
      const float* aptr = p.add ? p.add + (long long)b * p.sadd + (long long)(m0 + c0) * p.ldadd + pix : nullptr;
      mbar_wait_bounded(&tmem_full[acc], use & 1u);
      tc_fence_after();
      uint32_t r[32];
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + acc * CT_BN + (uint32_t)c0;
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
      // this warp's TMEM reads of accumulator `acc` are done: hand it back to the MMA issuer before the math
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[acc]);
      if (pix_ok && nch > 0) {
        // two halves of 16 channels keep the live register set small; the mode flags are tested once per half, not per
        // element (the kernel is issue-bound in this loop: profiles/r01_ncu_full_conv_tc_pixelmajor_raw.csv)
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int jb = h * 16;
          const int nv = nch - jb;   // valid channels in this half
          if (nv <= 0) break;
          float v[16];
          const float4* b4 = reinterpret_cast<const float4*>(bias_s + c0 + jb);
#pragma unroll
          for (int j = 0; j < 16; j += 4) {
            const float4 bb = b4[j >> 2];
            v[j] = __uint_as_float(r[jb + j]) + bb.x;
            v[j + 1] = __uint_as_float(r[jb + j + 1]) + bb.y;
            v[j + 2] = __uint_as_float(r[jb + j + 2]) + bb.z;
            v[j + 3] = __uint_as_float(r[jb + j + 3]) + bb.w;
          }
          if (p.act_gelu == 1) {
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] = gelu_tanh3(v[j]);
          }
          if (aptr) {
            const float* ap = aptr + (long long)jb * p.ldadd;
            float av[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) av[j] = (j < nv) ? __ldg(ap + (long long)j * p.ldadd) : 0.0f;   // coalesced over lanes
            if (p.act_gelu == 2) {   // activation adjoint: the operand is the GELU's pre-activation
#pragma unroll
              for (int j = 0; j < 16; ++j) v[j] *= gelu_erf_grad(av[j]);
            } else {
#pragma unroll
              for (int j = 0; j < 16; ++j) v[j] += av[j];
            }
          }
          if (p.round_tf32) {
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] = round_to_tf32(v[j]);
          }
          float* dp = dptr + (long long)jb * p.ldd;
          if (nv >= 16) {
#pragma unroll
            for (int j = 0; j < 16; ++j) __stcs(dp + (long long)j * p.ldd, v[j]);                          // coalesced over lanes
          } else {
#pragma unroll
            for (int j = 0; j < 16; ++j)
              if (j < nv) __stcs(dp + (long long)j * p.ldd, v[j]);
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem_base), "r"(2 * CT_BN));
  }
}

// Returns MSFNO_OK and sets *handled = 1 when the persistent kernel took the problem; *handled = 0 means "use the
// generic engine" (shape outside what the weight-stationary layout supports).
int launch_conv_tc(const GemmLaunch& g, long long a_rows, long long a_cols, long long b_rows, long long b_cols,
                   long long a2_rows, long long a2_cols, long long b2_rows, long long b2_cols, int* handled,
                   cudaStream_t st, int round_tf32) {
  *handled = 0;
  int K1 = g.single.K, K2 = g.A2 ? g.K2 : 0;
  // a single operand pair whose weight slab does not fit in smem is streamed entirely (it takes the role of pair 2)
  const bool stream_all = !g.A2 && (K1 + TC_BK - 1) / TC_BK > CT_MAX_KB && g.sa == 0;
  if (stream_all) { K2 = K1; K1 = 0; }
  const int nkb1 = (K1 + TC_BK - 1) / TC_BK, nkb2 = (K2 + TC_BK - 1) / TC_BK;
  if (g.b_kmajor || !g.use_single || g.relu_even || nkb1 + nkb2 < 1 || nkb1 > CT_MAX_KB || nkb2 > CT_MAX_KB2) return MSFNO_OK;
  if (g.sa % g.lda != 0 || g.sb % g.ldb != 0 || (g.A2 && (g.sa2 != 0 || g.sb2 % g.ldb2 != 0))) return MSFNO_OK;
  // smem: [align 1024][nkb1 resident weight blocks][ns ring slots][barriers + bias, 1024]
  const size_t total = 227 * 1024, fixed = 1024 + 1024 + (size_t)nkb1 * CT_KB_BYTES;
  const size_t slot = nkb2 ? 2 * CT_KB_BYTES : CT_KB_BYTES;
  if (fixed + 2 * slot > total) return MSFNO_OK;
  int ns = (int)((total - fixed) / slot);
  if (ns > 6) ns = 6;
  const size_t smem = fixed + (size_t)ns * slot;

  CUtensorMap tmA, tmB, tmA2, tmB2;
  int rc = make_map(&tmA, g.A, a_rows, a_cols, g.lda, CT_BM);
  if (rc) return rc;
  rc = make_map(&tmB, g.B, b_rows, b_cols, g.ldb, TC_BK, true);
  if (rc) return rc;
  if (g.A2) {
    rc = make_map(&tmA2, g.A2, a2_rows, a2_cols, g.lda2, CT_BM);
    if (rc) return rc;
    rc = make_map(&tmB2, g.B2, b2_rows, b2_cols, g.ldb2, TC_BK, true);
    if (rc) return rc;
  } else {
    tmA2 = tmA;   // stream_all: pair 2 is the only pair and uses the same buffers
    tmB2 = tmB;
  }
  static PerDeviceOnce once;
  MSFNO_CUDA_OK(once.run([] { return cudaFuncSetAttribute(conv_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024); }));
  ConvTcParams p{};
  p.D = g.D; p.lda = g.lda; p.lda2 = g.A2 ? g.lda2 : g.lda; p.ldb = g.ldb; p.ldb2 = g.A2 ? g.ldb2 : g.ldb; p.ldd = g.ldd;
  p.sa = g.sa; p.sb = g.sb; p.sb2 = g.A2 ? g.sb2 : g.sb; p.sd = g.sd;
  p.bias = g.bias; p.sbias = g.sbias; p.add = g.add; p.ldadd = g.ldadd; p.sadd = g.sadd;
  p.M = g.single.M; p.N = g.single.N; p.K1 = K1; p.K2 = K2; p.act_gelu = g.act_gelu; p.nstages = ns; p.slot_bytes = (int)slot; p.round_tf32 = round_tf32;
  p.tilesN = (p.N + CT_BN - 1) / CT_BN;
  const int tilesM = (p.M + CT_BM - 1) / CT_BM;
  int sms = 148;
  int gx = sms / (tilesM * g.ngroups);
  if (gx < 1) gx = 1;
  if (gx > p.tilesN) gx = p.tilesN;
  dim3 grid(gx, tilesM, g.ngroups);
  MSFNO_CUDA_OK(launch_pdl(conv_tc_kernel, grid, dim3(576), smem, st, tmA, tmB, tmA2, tmB2, p));
  count_launch();
  MSFNO_CUDA_OK(cudaGetLastError());
  *handled = 1;
  return MSFNO_OK;
}

}  // namespace msfno
