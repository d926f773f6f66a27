// Fused two-layer 1x1-conv MLP on the tensor cores (tensor-core tier of msfno_mlp1x1_fwd):
//
//   y[b][o][p] = sum_h w2[o][h] * gelu( sum_c w1[h][c] x[b][c][p] + sum_c w1b[h][c] x2[b][c][p] + b1[h] ) + b2[o] + add[b][o][p]
//
// replaces: MLP.fwd = Conv2d(1x1) -> GELU -> Conv2d(1x1) of the encoder and decoder at full resolution, plus the
//   pos_embed add and the big-skip torch.cat (/root/reference MSFNO/Models/sfno/layers.py:161-168,
//   sfnonet.py:671,682-684).  Run as two separate GEMM kernels the 256-channel hidden activation costs a 1.06 GB
//   write and a 1.06 GB read per MLP at 721x1440; here it never leaves the SM.
//
// A persistent CTA walks 128-pixel tiles:
//   GEMM1  acc1[128 px][Chid] = X[128 px][Cin (+Cin2)] * W1^T   -- activations are the MN-major A operand streamed by TMA
//                                                                  (pixels on the TMEM lanes), weights the K-major B operand
//   EPI1   acc1 <- tf32(gelu(acc1 + b1))  IN PLACE in TMEM (tcgen05.ld -> registers -> tcgen05.st)
//   GEMM2  acc2[128 px][Cout] = acc1 * W2^T                     -- A operand read straight from TMEM
//   ADD    acc2 += add tile * I32  -- the residual / pos-embed operand streams through the same TMA ring as the
//          activations and is accumulated by 32 tiny N = 32 MMAs against a resident 32 x 32 identity block
//          (plain loads of it from the epilogue cost 13 of 18 kclk per tile and nothing made them faster)
//   EPI2   y = acc2 + b2, coalesced 128-byte row stores (lane = pixel)
// A hidden layer wider than the 256 TMEM columns (the 512-channel block MLPs) is processed in chunks of 256: GEMM1 /
// EPI1 per chunk, GEMM2 accumulating the chunk's K range into acc2 (the activation blocks are re-streamed per chunk).
// Both weight matrices stream through a shared-memory ring of 32 KB k-blocks (L2 resident: 0.3-0.4 MB in total);
// the producer runs ahead across phases, so GEMM2 finds its blocks waiting.  EPI2 of tile t overlaps GEMM1 of t+1.
//
// Warp roles (576 threads): warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer, warps 2-17 = epilogues
// (TMEM lane quarter w % 4 = 32-pixel block; column quarter (w - 2) / 4).
#include <cstdlib>

#include "plan.h"
#include "tc_common.cuh"

namespace msfno {

static constexpr int ML_BM = 128;                       // pixels per tile
static constexpr int ML_XBLK = ML_BM * TC_BK * 4;       // activation k-block: 32 channels x 128 pixels = 16 KB
static constexpr int ML_WBLK = 256 * TC_BK * 4;         // weight k-block slot: up to 256 rows x 32 k = 32 KB
#ifndef MSFNO_ML_NSX
#define MSFNO_ML_NSX 5
#define MSFNO_ML_NSW 4
#endif
// activation blocks come from HBM (2 us under load), weight blocks from L2: the deep ring belongs to the activations
static constexpr int ML_NSX = MSFNO_ML_NSX, ML_NSW = MSFNO_ML_NSW;
static constexpr int ML_MAX_HID = 1024;
static constexpr int ML_IDBLK = 32 * TC_BK * 4;         // resident 32 x 32 identity (K-major, 128-byte swizzle) = 4 KB
static constexpr int ML_SMEM = 1024 + ML_NSX * ML_XBLK + ML_NSW * ML_WBLK + ML_IDBLK + (ML_MAX_HID + 256 + 4 * 512) * 4 + 512;

struct MlpTcParams {
  float* D;
  long long ldd, sd;               // output plane stride (= HW) and sample stride
  const float* b1; long long sb1;  // hidden bias [Chid], per-sample stride
  const float* b2;                 // output bias [Cout] (+ b * sb2: per-sample bias) or null
  long long sb2;
  const float* add; long long ldadd, sadd;
  long long x_rows_per_sample, x2_rows_per_sample, w1_rows_per_sample;   // row offsets (tensor-map rows) per sample
  int HW, K1a, K1b, Chid, Cout, N2pad;
  int tiles;
  int round_tf32;
  int add_tma;      // 1: `add` is accumulated on the tensor cores (tmAdd / tmId), the epilogue does not read it
  long long add_rows_per_sample;
  double* stats;   // [B][Cout][2] plane (sum, sum of squares) of y, accumulated with atomics, or null
};

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

// Sums over the 32 lanes (pixels) of 16 per-lane values (channels) and of their squares, added to acc[ch] / acc[256 + ch]
// in shared memory.  Halving butterfly: at each step a lane keeps half of its values and receives the partner's copy of
// the same half, so 16 values need 16 shuffles per quantity instead of 80.  Afterwards lane pair (2k, 2k+1) holds channel
// bitrev-free index ch = lane >> 1.
__device__ __forceinline__ void warp_channel_sums(const float (&v)[16], float* acc, int nv, int lane) {
  float s[16], q[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) { s[j] = v[j]; q[j] = v[j] * v[j]; }
#pragma unroll
  for (int half = 8, bit = 16; half >= 1; half >>= 1, bit >>= 1) {
    const bool up = (lane & bit) != 0;
#pragma unroll
    for (int j = 0; j < half; ++j) {
      const float send_s = up ? s[j] : s[j + half], send_q = up ? q[j] : q[j + half];
      const float keep_s = up ? s[j + half] : s[j], keep_q = up ? q[j + half] : q[j];
      s[j] = keep_s + __shfl_xor_sync(0xffffffffu, send_s, bit);
      q[j] = keep_q + __shfl_xor_sync(0xffffffffu, send_q, bit);
    }
  }
  s[0] += __shfl_xor_sync(0xffffffffu, s[0], 1);
  q[0] += __shfl_xor_sync(0xffffffffu, q[0], 1);
  // lane bits (4,3,2,1) selected the upper half at steps (8,4,2,1): channel = 8 b4 + 4 b3 + 2 b2 + b1 = lane >> 1
  const int ch = lane >> 1;
  if ((lane & 1) == 0 && ch < nv) {
    acc[ch] += s[0];         // the slot belongs to this warp alone (pixel quadrant x channel quarter): no atomics, and the
    acc[256 + ch] += q[0];   // order of the additions is the CTA's fixed tile order -> bit-reproducible plane sums
  }
}

#define MSFNO_TMEM_LD32(r, taddr)                                                                                        \
  asm volatile(                                                                                                          \
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "                                                                          \
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "                                          \
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"                        \
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),      \
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),           \
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),          \
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])                        \
      : "r"(taddr))

#define MSFNO_TMEM_ST32(taddr, r)                                                                                        \
  asm volatile(                                                                                                          \
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "                                                                    \
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "                                         \
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};\n" ::"r"(taddr),                 \
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),      \
      "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]),        \
      "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]),        \
      "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])                                                                     \
      : "memory")

__global__ void __launch_bounds__(576, 1)
mlp_tc_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmX2,
              const __grid_constant__ CUtensorMap tmW1, const __grid_constant__ CUtensorMap tmW1b,
              const __grid_constant__ CUtensorMap tmW2, const __grid_constant__ CUtensorMap tmAdd,
              const __grid_constant__ CUtensorMap tmId, MlpTcParams p) {
  extern __shared__ uint8_t smem_raw[];
  pdl_trigger();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.z;
  const int nkb1a = (p.K1a + TC_BK - 1) / TC_BK, nkb1b = (p.K1b + TC_BK - 1) / TC_BK, nkb1 = nkb1a + nkb1b;
  const int HC = min(p.Chid, 256), nh = p.Chid / HC;      // hidden chunk (TMEM columns of acc1) and chunk count
  const int nkb2 = HC / TC_BK;                            // GEMM2 k-blocks per chunk
  // EPI1 activates the hidden chunk in two column halves; GEMM2's first k-blocks start after the first half, so the
  // tensor pipe runs while the second half is still being activated
  const int nhalf = (HC % 64 == 0) ? 2 : 1, hwid = HC / nhalf;
  const uint32_t w2_bytes = (uint32_t)p.N2pad * TC_BK * 4, w1_bytes = (uint32_t)HC * TC_BK * 4;

  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* tiles = smem_raw + (base - smem_u32(smem_raw));
  uint8_t* xring = tiles;
  uint8_t* wring = tiles + ML_NSX * ML_XBLK;
  uint8_t* idblk = wring + ML_NSW * ML_WBLK;                           // identity block (add path)
  float* b1_s = reinterpret_cast<float*>(idblk + ML_IDBLK);           // [ML_MAX_HID]
  float* b2_s = b1_s + ML_MAX_HID;                                    // [256]
  float* stat_s = b2_s + 256;                                         // [4 pixel quadrants][2][256] per-CTA plane sums of the output
  uint64_t* bars = reinterpret_cast<uint64_t*>(stat_s + 4 * 512);
  uint64_t* xfull = bars;            // [NSX]
  uint64_t* xempty = bars + 8;       // [NSX]
  uint64_t* wfull = bars + 16;       // [NSW]
  uint64_t* wempty = bars + 24;      // [NSW]
  uint64_t* acc1_full = bars + 32;   // GEMM1 of a tile complete
  uint64_t* h_ready = bars + 33;     // [2] EPI1 wrote the activated hidden columns of half 0 / half 1 back to TMEM (16 warps each)
  uint64_t* acc2_full = bars + 35;   // GEMM2 complete
  uint64_t* acc2_empty = bars + 36;  // EPI2 finished reading acc2 (16 warps)
  uint64_t* id_full = bars + 37;     // identity block landed
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 38);

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < ML_NSX; ++s) { mbar_init(&xfull[s], 1); mbar_init(&xempty[s], 1); }
    for (int s = 0; s < ML_NSW; ++s) { mbar_init(&wfull[s], 1); mbar_init(&wempty[s], 1); }
    mbar_init(acc1_full, 1);
    mbar_init(&h_ready[0], 16);
    mbar_init(&h_ready[1], 16);
    mbar_init(acc2_full, 1);
    mbar_init(acc2_empty, 16);
    mbar_init(id_full, 1);
    fence_mbar_init();
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(tmem_slot)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();   // everything above overlapped the previous kernel's tail
  if (threadIdx.x >= 64 && threadIdx.x < 64 + 256) {
    const int c = (int)threadIdx.x - 64;
    for (int h = c; h < ML_MAX_HID; h += 256) b1_s[h] = (p.b1 && h < p.Chid) ? p.b1[(long long)b * p.sb1 + h] : 0.0f;
    b2_s[c] = (p.b2 && c < p.Cout) ? p.b2[(long long)b * p.sb2 + c] : 0.0f;
#pragma unroll
    for (int j = 0; j < 8; ++j) stat_s[j * 256 + c] = 0.0f;
  }
  __syncthreads();
  const uint32_t tmem_acc1 = tmem_base, tmem_acc2 = tmem_base + 256;

  if (warp == 0) {
    if (lane == 0) {
      // ---------------- TMA producer: activation ring + weight ring ----------------
      const int xrow = (int)(b * p.x_rows_per_sample), x2row = (int)(b * p.x2_rows_per_sample);
      const int w1row = (int)(b * p.w1_rows_per_sample);
      uint32_t xc = 0, wc = 0;
      if (p.add_tma) {
        mbar_arrive_expect_tx(id_full, ML_IDBLK);
        tma_load_2d(idblk, &tmId, id_full, 0, 0);
      }
      const int nadd = p.add_tma ? (p.Cout + TC_BK - 1) / TC_BK : 0;
      const int addrow = (int)(b * p.add_rows_per_sample);
      for (int t = blockIdx.x; t < p.tiles; t += gridDim.x) {
        const int n0 = t * ML_BM;
        for (int hh = 0; hh < nh; ++hh) {
          for (int kb = 0; kb < nkb1; ++kb, ++xc, ++wc) {
            const bool second = kb >= nkb1a;
            const int kk = (second ? kb - nkb1a : kb) * TC_BK;
            {  // weight block of GEMM1: HC hidden rows x 32 k
              const int s = wc % ML_NSW;
              mbar_wait_bounded(&wempty[s], ((wc / ML_NSW) & 1u) ^ 1u);
              mbar_arrive_expect_tx(&wfull[s], w1_bytes);
              tma_load_2d(wring + (size_t)s * ML_WBLK, second ? &tmW1b : &tmW1, &wfull[s], kk, (second ? 0 : w1row) + hh * HC);
            }
            {  // activation block: 32 channels x 128 pixels as four 32 x 32 boxes
              const int s = xc % ML_NSX;
              mbar_wait_bounded(&xempty[s], ((xc / ML_NSX) & 1u) ^ 1u);
              mbar_arrive_expect_tx(&xfull[s], ML_XBLK);
              uint8_t* dst = xring + (size_t)s * ML_XBLK;
#pragma unroll
              for (int j = 0; j < ML_BM / 32; ++j)
                tma_load_2d(dst + j * (TC_BK * 128), second ? &tmX2 : &tmX, &xfull[s], n0 + 32 * j, (second ? x2row : xrow) + kk);
            }
          }
          for (int kb = 0; kb < nkb2; ++kb, ++wc) {  // weight blocks of GEMM2: N2pad rows x 32 k of this hidden chunk
            const int s = wc % ML_NSW;
            mbar_wait_bounded(&wempty[s], ((wc / ML_NSW) & 1u) ^ 1u);
            mbar_arrive_expect_tx(&wfull[s], w2_bytes);
            tma_load_2d(wring + (size_t)s * ML_WBLK, &tmW2, &wfull[s], hh * HC + kb * TC_BK, 0);
          }
        }
        for (int kb = 0; kb < nadd; ++kb, ++xc) {  // `add` tile: 32 channels x 128 pixels per block, through the activation ring
          const int s = xc % ML_NSX;
          mbar_wait_bounded(&xempty[s], ((xc / ML_NSX) & 1u) ^ 1u);
          mbar_arrive_expect_tx(&xfull[s], ML_XBLK);
          uint8_t* dst = xring + (size_t)s * ML_XBLK;
#pragma unroll
          for (int j = 0; j < ML_BM / 32; ++j) tma_load_2d(dst + j * (TC_BK * 128), &tmAdd, &xfull[s], n0 + 32 * j, addrow + kb * TC_BK);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // ---------------- MMA issuer ----------------
      // GEMM1: A = activations (MN-major, bit 15), B = W1 (K-major), N = Chid, M = 128 pixels
      const uint32_t idesc1 = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | ((uint32_t)(HC >> 3) << 17) |
                              ((uint32_t)(ML_BM >> 4) << 24);
      // GEMM2: A = hidden tile in TMEM, B = W2 (K-major), N = N2pad
      const uint32_t idesc2 = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(p.N2pad >> 3) << 17) | ((uint32_t)(ML_BM >> 4) << 24);
      const uint32_t x_addr = base, w_addr = base + ML_NSX * ML_XBLK;
      uint32_t xc = 0, wc = 0, it = 0, seq = 0;
      for (int t = blockIdx.x; t < p.tiles; t += gridDim.x, ++it) {
        for (int hh = 0; hh < nh; ++hh, ++seq) {
          // acc1 is free: the tensor pipe executes in issue order, the previous GEMM2 (its last reader) is ahead
          for (int kb = 0; kb < nkb1; ++kb, ++xc, ++wc) {
            const int sx = xc % ML_NSX, sw = wc % ML_NSW;
            mbar_wait_bounded(&wfull[sw], (wc / ML_NSW) & 1u);
            mbar_wait_bounded(&xfull[sx], (xc / ML_NSX) & 1u);
            tc_fence_after();
            const uint32_t sa = x_addr + (uint32_t)sx * ML_XBLK, sb = w_addr + (uint32_t)sw * ML_WBLK;
#pragma unroll
            for (int k = 0; k < TC_BK / 8; ++k)
              tc_mma_tf32(tmem_acc1, make_smem_desc(sa + 1024 * k, TC_BK * 128, 512, 1), make_smem_desc(sb + 32 * k, 16, 1024),
                          idesc1, (kb | k) ? 1u : 0u);
            tc_commit(&xempty[sx]);
            tc_commit(&wempty[sw]);
          }
          tc_commit(acc1_full);
          mbar_wait_bounded(&h_ready[0], seq & 1u);                 // first half of the hidden chunk activated in place
          if (hh == 0) mbar_wait_bounded(acc2_empty, (it & 1u) ^ 1u);   // previous tile's EPI2 has drained acc2
          tc_fence_after();
          for (int kb = 0; kb < nkb2; ++kb, ++wc) {
            if (kb * TC_BK == hwid || (nhalf == 1 && kb == 0)) mbar_wait_bounded(&h_ready[1], seq & 1u);   // second half
            const int sw = wc % ML_NSW;
            mbar_wait_bounded(&wfull[sw], (wc / ML_NSW) & 1u);
            tc_fence_after();
            const uint32_t sb = w_addr + (uint32_t)sw * ML_WBLK;
#pragma unroll
            for (int k = 0; k < TC_BK / 8; ++k)
              tc_mma_tf32_ts(tmem_acc2, tmem_acc1 + (uint32_t)(kb * TC_BK + k * 8), make_smem_desc(sb + 32 * k, 16, 1024), idesc2,
                             (hh | kb | k) ? 1u : 0u);
            tc_commit(&wempty[sw]);
          }
        }
        if (p.add_tma) {
          // acc2[:, 32 kb .. 32 kb + 32) += add block * I32  (A = add tile, MN-major; B = identity, K-major; N = 32)
          const uint32_t idesc3 = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | ((uint32_t)(32 >> 3) << 17) | ((uint32_t)(ML_BM >> 4) << 24);
          const uint32_t id_addr = w_addr + ML_NSW * ML_WBLK;
          if (it == 0) { mbar_wait_bounded(id_full, 0); tc_fence_after(); }
          const int nadd = (p.Cout + TC_BK - 1) / TC_BK;
          for (int kb = 0; kb < nadd; ++kb, ++xc) {
            const int sx = xc % ML_NSX;
            mbar_wait_bounded(&xfull[sx], (xc / ML_NSX) & 1u);
            tc_fence_after();
            const uint32_t sa = x_addr + (uint32_t)sx * ML_XBLK;
#pragma unroll
            for (int k = 0; k < TC_BK / 8; ++k)
              tc_mma_tf32(tmem_acc2 + (uint32_t)(kb * TC_BK), make_smem_desc(sa + 1024 * k, TC_BK * 128, 512, 1),
                          make_smem_desc(id_addr + 32 * k, 16, 1024), idesc3, 1u);
            tc_commit(&xempty[sx]);
          }
        }
        tc_commit(acc2_full);
      }
    }
  } else {
    // ---------------- epilogue warps 2..17: lane = pixel, column quarter cq ----------------
    const int q = warp & 3, cq = (warp - 2) >> 2;
    const uint32_t lane_off = (uint32_t)(q * 32) << 16;
    uint32_t it = 0, seq = 0;
    for (int t = blockIdx.x; t < p.tiles; t += gridDim.x, ++it) {
      // ---- EPI1 per hidden chunk: columns [cq*64, cq*64+64) of acc1, in place
      for (int hh = 0; hh < nh; ++hh, ++seq) {
        mbar_wait_bounded(acc1_full, seq & 1u);
        tc_fence_after();
#pragma unroll 1
        for (int h = 0; h < 2; ++h) {
          if (h < nhalf) {
#pragma unroll 1
            for (int c0 = h * hwid + cq * 32; c0 < (h + 1) * hwid; c0 += 128) {
              uint32_t r[32];
              const uint32_t taddr = tmem_acc1 + lane_off + (uint32_t)c0;
              MSFNO_TMEM_LD32(r, taddr);
              asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
              const float* bb = b1_s + hh * HC + c0;
#pragma unroll
              for (int j = 0; j < 32; ++j) r[j] = __float_as_uint(round_to_tf32_pretrunc(gelu_tanh3(__uint_as_float(r[j]) + bb[j])));
              MSFNO_TMEM_ST32(taddr, r);
            }
            asm volatile("tcgen05.wait::st.sync.aligned;\n" ::: "memory");
          }
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&h_ready[h]);
        }
      }

      // ---- EPI2: output channels [cq*64, cq*64+64) of acc2
      const int pix = t * ML_BM + q * 32 + lane;
      const bool pix_ok = pix < p.HW;
      mbar_wait_bounded(acc2_full, it & 1u);
      tc_fence_after();
      auto store_chunk = [&](const uint32_t (&r)[32], const int c0) {
        const int nch = min(32, p.Cout - c0);
        if (nch <= 0 || (!pix_ok && !p.stats)) return;   // with statistics every lane takes part in the warp reductions
        float* dptr = p.D + (long long)b * p.sd + (long long)c0 * p.ldd + pix;
        const float* aptr = (p.add && !p.add_tma) ? p.add + (long long)b * p.sadd + (long long)c0 * p.ldadd + pix : nullptr;
#pragma unroll
        for (int g = 0; g < 2; ++g) {
          const int jb = g * 16, nv = nch - jb;
          if (nv > 0) {
            float v[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(r[jb + j]) + b2_s[c0 + jb + j];
            if (aptr) {
              float av[16];
#pragma unroll
              for (int j = 0; j < 16; ++j) av[j] = (j < nv && pix_ok) ? __ldg(aptr + (long long)(jb + j) * p.ldadd) : 0.0f;
#pragma unroll
              for (int j = 0; j < 16; ++j) v[j] += av[j];
            }
            if (p.round_tf32) {
#pragma unroll
              for (int j = 0; j < 16; ++j) v[j] = round_to_tf32(v[j]);
            }
#pragma unroll
            for (int j = 0; j < 16; ++j)
              if (j < nv && pix_ok) __stcs(dptr + (long long)(jb + j) * p.ldd, v[j]);
            if (!pix_ok) {
#pragma unroll
              for (int j = 0; j < 16; ++j) v[j] = 0.0f;
            }
            if (p.stats) warp_channel_sums(v, stat_s + q * 512 + c0 + jb, nv, lane);
          }
        }
      };
#pragma unroll 1
      for (int h = 0; h < 2; ++h) {
        const int c0 = cq * 64 + h * 32;
        uint32_t r[32];
        if (c0 < p.N2pad) {
          MSFNO_TMEM_LD32(r, tmem_acc2 + lane_off + (uint32_t)c0);
          asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
        }
        if (h == 1) {   // both chunks of this warp are out of acc2 (EPI1 of the next tile follows on these same warps)
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(acc2_empty);
        }
        if (c0 < p.N2pad) store_chunk(r, c0);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (p.stats && threadIdx.x < 512) {
    const int ch = threadIdx.x & 255, which = threadIdx.x >> 8;
    if (ch < p.Cout) {   // quadrant slots summed in a fixed order; only the cross-CTA fp64 atomics are unordered (1e-16)
      const float* sp = stat_s + which * 256 + ch;
      atomicAdd(&p.stats[2 * ((size_t)b * p.Cout + ch) + which], ((double)sp[0] + (double)sp[512]) + ((double)sp[1024] + (double)sp[1536]));
    }
  }
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem_base), "r"(512));
  }
}

// 2-D fp32 tensor [rows][ld] (cols valid) with an arbitrary box height (K-major weight blocks: 32 columns x box_rows rows)
static int make_wmap(CUtensorMap* tm, const float* base, long long rows, long long cols, long long ld, int box_rows) {
  return make_map(tm, base, rows, cols, ld, box_rows, false);
}

}  // namespace msfno

using namespace msfno;

extern "C" int msfno_mlp1x1_fwd(const float* x, long x_bstride, int Cin, const float* w1, long ldw1, long w1_bstride,
                                const float* x2, long x2_bstride, int Cin2, const float* w1b, long ldw1b, const float* b1,
                                long b1_bstride, int Chid, const float* w2, long ldw2, const float* b2, long b2_bstride, const float* add,
                                long add_bstride, float* y, double* stats, int B, int Cout, long HW, int flags, void* stream) {
  if (!x || !w1 || !w2 || !y || B < 1 || Cin < 1 || Chid < 1 || Cout < 1 || HW < 1 || ldw1 < Cin || ldw2 < Chid ||
      (x2 && (!w1b || Cin2 < 1 || ldw1b < Cin2)))
    return record_error(MSFNO_ERR_BAD_SHAPE, "mlp1x1_fwd: bad argument");
  auto al16 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
  if (Chid % 32 != 0 || (Chid > 256 && Chid % 256 != 0) || Chid > ML_MAX_HID || Cout > 256 || HW % 4 != 0 || HW > 0x7fffffffL || (ldw1 & 3) || (ldw2 & 3) ||
      (x2 && (ldw1b & 3)) || x_bstride % HW != 0 || (x2 && x2_bstride % HW != 0) || w1_bstride % ldw1 != 0 || !al16(x) ||
      !al16(w1) || !al16(w2) || (x2 && (!al16(x2) || !al16(w1b))) || get_encode() == nullptr)
    return record_error(MSFNO_ERR_UNSUPPORTED, "mlp1x1_fwd: shape / alignment outside the fused tensor-core kernel");
  cudaStream_t st = (cudaStream_t)stream;
  const int N2pad = (Cout + 15) / 16 * 16;
  CUtensorMap tmX, tmX2, tmW1, tmW1b, tmW2;
  int rc = make_map(&tmX, x, (long long)(B - 1) * (x_bstride / HW) + Cin, HW, HW, TC_BK, true);
  if (rc) return rc;
  rc = make_wmap(&tmW1, w1, (long long)(B - 1) * (w1_bstride / ldw1) + Chid, Cin, ldw1, Chid < 256 ? Chid : 256);
  if (rc) return rc;
  if (x2) {
    rc = make_map(&tmX2, x2, (long long)(B - 1) * (x2_bstride / HW) + Cin2, HW, HW, TC_BK, true);
    if (rc) return rc;
    rc = make_wmap(&tmW1b, w1b, Chid, Cin2, ldw1b, Chid < 256 ? Chid : 256);
    if (rc) return rc;
  } else {
    tmX2 = tmX;
    tmW1b = tmW1;
  }
  rc = make_wmap(&tmW2, w2, Cout, Chid, ldw2, N2pad);
  if (rc) return rc;
  // `add` on the tensor cores: needs a TMA-able operand (16-byte aligned, plane stride HW) and the identity block
  CUtensorMap tmAdd = tmX, tmId = tmW2;
  const float* d_ident = identity32_device();
  static const bool add_tma_off = dbg_env("MSFNO_MLP_ADD_LSU");
  const bool add_tma = add && !add_tma_off && al16(add) && add_bstride % HW == 0 && d_ident != nullptr;
  if (add_tma) {
    rc = make_map(&tmAdd, add, (add_bstride ? (long long)(B - 1) * (add_bstride / HW) : 0) + Cout, HW, HW, TC_BK, true);
    if (rc) return rc;
    rc = make_wmap(&tmId, d_ident, 32, 32, 32, 32);
    if (rc) return rc;
  }
  static PerDeviceOnce once;
  MSFNO_CUDA_OK(once.run([] { return cudaFuncSetAttribute(mlp_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, ML_SMEM); }));
  MlpTcParams p{};
  p.D = y; p.ldd = HW; p.sd = (long long)Cout * HW;
  p.b1 = b1; p.sb1 = b1_bstride; p.b2 = b2; p.sb2 = b2_bstride;
  p.add = add; p.ldadd = HW; p.sadd = add_bstride;
  p.x_rows_per_sample = x_bstride / HW; p.x2_rows_per_sample = x2 ? x2_bstride / HW : 0; p.w1_rows_per_sample = w1_bstride / ldw1;
  p.HW = (int)HW; p.K1a = Cin; p.K1b = x2 ? Cin2 : 0; p.Chid = Chid; p.Cout = Cout; p.N2pad = N2pad;
  p.tiles = (int)((HW + ML_BM - 1) / ML_BM);
  p.round_tf32 = (flags >> 1) & 1;
  p.add_tma = add_tma ? 1 : 0;
  p.add_rows_per_sample = add_bstride / HW;
  p.stats = stats;
  if (stats) MSFNO_CUDA_OK(cudaMemsetAsync(stats, 0, sizeof(double) * 2 * (size_t)B * Cout, st));
  int dev = 0, sms = 0;
  MSFNO_CUDA_OK(cudaGetDevice(&dev));
  MSFNO_CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  int gx = sms / B;
  if (gx < 1) gx = 1;
  if (gx > p.tiles) gx = p.tiles;
  MSFNO_CUDA_OK(launch_pdl(mlp_tc_kernel, dim3(gx, 1, B), dim3(576), ML_SMEM, st, tmX, tmX2, tmW1, tmW1b, tmW2, tmAdd, tmId, p));
  count_launch();
  MSFNO_CUDA_OK(cudaGetLastError());
  return MSFNO_OK;
}
