// Device / host helpers shared by the tcgen05 kernels (gemm_tc.cu, conv_tc.cu): TMA loads, bounded mbarrier waits,
// tcgen05 fences / commit / MMA, shared-memory matrix descriptors, tensor-map construction.
#pragma once
#include <cuda.h>

#include <mutex>

#include "common.cuh"

namespace msfno {

static constexpr int TC_BK = 32;  // fp32 elements per 128-byte swizzle span

__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* tm, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];\n" ::"r"(
          smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(tm)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
// try_wait in a bounded loop: a faulty descriptor traps (error surfaced to the host) instead of hanging the SM
__device__ __forceinline__ void mbar_wait_bounded(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
#pragma unroll 1   // (nvcc otherwise unrolls this poll 64 times at every call site: 4 200 of dft_tc.o's 23 600 SASS lines)
  for (uint32_t it = 0; it < (1u << 24); ++it) {
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n"
        "selp.b32 %0, 1, 0, P1;\n"
        "}\n"
        : "=r"(ok)
        : "r"(addr), "r"(parity)
        : "memory");
    if (ok) return;
  }
  __trap();
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}
// D[tmem] (+)= A[smem desc] * B[smem desc]
__device__ __forceinline__ void tc_mma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Shared-memory matrix descriptor, 128-byte swizzle, descriptor version 1.
//   K-major : rows of 128 B (32 tf32 along K), 8-row atoms of 1024 B -> SBO = 1024, LBO unused (= 1)
//   MN-major: k-rows of 128 B (32 tf32 along N), 8-k-row atoms of 1024 B -> SBO = 1024 (next 8 k),
//             LBO = byte distance between consecutive 32-element N blocks
//   layout_type: 2 = SWIZZLE_128B (16-byte swizzle atomicity), 1 = SWIZZLE_128B_BASE32B (32-byte atomicity: the only
//   layout the tensor core accepts for MN-major 32-bit operands; its atoms are 4 k-rows of 128 B)
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes,
                                                   uint32_t layout_type = 2) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)layout_type << 61;
  return d;
}
// round to nearest (ties away) TF32, as cvt.rna.tf32.f32 for every finite input; two integer instructions instead of
// the three (with an inf/nan test) the cvt expands to -- the epilogues that call it are issue-bound
__device__ __forceinline__ float round_to_tf32(float x) { return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xffffe000u); }
// operand that the tensor core itself will truncate (A operand written to TMEM): the add alone completes the rounding
__device__ __forceinline__ float round_to_tf32_pretrunc(float x) { return __uint_as_float(__float_as_uint(x) + 0x1000u); }

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}

// Epilogue store of a 32 x 32 block held one ROW per lane (v[32] = 32 consecutive columns) to a row-major matrix.
// Stored straight from the registers a warp instruction touches 32 different rows (32 L1 wavefronts per STG.128: the
// per-CTA clock trace of the pair GEMM showed 10-14 kclk of epilogue for 16 kclk of MMAs).  Transposed through a private
// 32 x 36 shared-memory tile (conflict-free 16-byte writes and reads), every instruction writes four full 128-byte lines.
//   wt: this warp's 32 x 36 floats, 16-byte aligned;  dst: address of (first row, first column);  rows: valid rows (<= 32)
__device__ __forceinline__ void store_block_transposed(const float (&v)[32], float* wt, float* dst, long long ldd, int rows,
                                                       int lane, int cols = 32, bool accumulate = false) {
  float4* wrow = reinterpret_cast<float4*>(wt + lane * 36);
#pragma unroll
  for (int jj = 0; jj < 8; ++jj) wrow[jj] = make_float4(v[4 * jj], v[4 * jj + 1], v[4 * jj + 2], v[4 * jj + 3]);
  __syncwarp();
  const int rr = lane >> 3, c4 = lane & 7;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int row = rr + 4 * i;
    float4 t = *reinterpret_cast<const float4*>(wt + row * 36 + 4 * c4);
    if (row < rows) {
      float* o = dst + (long long)row * ldd + 4 * c4;
      if (4 * c4 + 3 < cols) {
        if (accumulate) {   // D += tile (weight gradients summed over the batch)
          const float4 old = *reinterpret_cast<const float4*>(o);
          t.x += old.x; t.y += old.y; t.z += old.z; t.w += old.w;
        }
        *reinterpret_cast<float4*>(o) = t;
      } else {   // ragged right edge: the last (partial) group of four columns
        if (4 * c4 < cols) o[0] = accumulate ? o[0] + t.x : t.x;
        if (4 * c4 + 1 < cols) o[1] = accumulate ? o[1] + t.y : t.y;
        if (4 * c4 + 2 < cols) o[2] = accumulate ? o[2] + t.z : t.z;
      }
    }
  }
  __syncwarp();
}

// [32][32] fp32 identity in device memory, one per device of this process (operand of the "add by identity MMA" paths
// of mlp_tc.cu and dft_tc.cu).  Returns nullptr if it cannot be allocated.
inline const float* identity32_device() {
  static std::mutex mu;
  static float* per_dev[64] = {nullptr};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
  std::lock_guard<std::mutex> lk(mu);
  if (!per_dev[dev]) {
    float h[32 * 32] = {0};
    for (int i = 0; i < 32; ++i) h[i * 32 + i] = 1.0f;
    float* d = nullptr;
    if (cudaMalloc(&d, sizeof(h)) != cudaSuccess) return nullptr;
    if (cudaMemcpy(d, h, sizeof(h), cudaMemcpyHostToDevice) != cudaSuccess) { cudaFree(d); return nullptr; }
    per_dev[dev] = d;
  }
  return per_dev[dev];
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline EncodeTiledFn get_encode() {
  // cuTensorMapEncodeTiled is a DRIVER call: it fails on a thread that has no current context yet.  PyTorch's autograd
  // worker threads only bind the primary context at their first runtime call, and a backward pass may reach a
  // tensor-map encode before any (measured: the first full-size backward of a fresh process).  One cudaFree(0) per
  // thread binds it.
  static thread_local bool ctx_bound = false;
  if (!ctx_bound) {
    cudaFree(0);
    ctx_bound = true;
  }
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

// 2-D fp32 tensor [rows][ld] (cols valid), box = 32 columns x box_rows rows, 128-byte swizzle
inline int make_map(CUtensorMap* tm, const float* base, long long rows, long long cols, long long ld, int box_rows,
                    bool atom32 = false) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return record_error(MSFNO_ERR_CUDA, "cuTensorMapEncodeTiled entry point unavailable");
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 4};
  cuuint32_t box[2] = {(cuuint32_t)TC_BK, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, atom32 ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return record_error(MSFNO_ERR_CUDA, "cuTensorMapEncodeTiled failed");
  return MSFNO_OK;
}


}  // namespace msfno
