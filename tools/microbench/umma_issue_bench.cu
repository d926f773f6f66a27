// Micro-benchmark: cost of one tcgen05.mma kind::tf32 (K = 8) as a function of N, of the A operand's source (shared
// memory K-major / MN-major, or TMEM) and of the CTA group -- no loads, no epilogue: one thread issues NMMA instructions
// back to back over operand blocks that already sit in shared memory, commits, and the clocks are read at issue end and at
// completion.  Question behind it: idft_eo_kernel issues M 128 x N 128 x K 8 MMAs with an
// MN-major A operand and shows ~150 clk per MMA for 64 clk of tensor-pipe work.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o umma_issue_bench umma_issue_bench.cu && ./umma_issue_bench
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, int c) { asm volatile("mbarrier.init.shared::cta.b64 [%1], %0;" ::"r"(c), "r"(s32(b))); }
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t ph) {
  uint32_t ok = 0;
  for (int it = 0; it < (1 << 24) && !ok; ++it)
    asm volatile("{.reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.b32 %0,1,0,p;}" : "=r"(ok) : "r"(s32(b)), "r"(ph) : "memory");
  if (!ok) __trap();
}
__device__ __forceinline__ uint64_t desc(uint32_t addr, uint32_t lbo, uint32_t sbo, uint32_t layout) {
  uint64_t d = 0;
  d |= (uint64_t)((addr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)layout << 61;
  return d;
}
__device__ __forceinline__ void mma_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{.reg .pred p; setp.ne.b32 p, %4, 0; tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;}" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{.reg .pred p; setp.ne.b32 p, %4, 0; tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;}" ::"r"(d), "r"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void commit(uint64_t* b) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(s32(b)) : "memory");
}

__device__ __forceinline__ bool elect_one() {
  uint32_t ok;
  asm volatile("{.reg .pred p; elect.sync _|p, 0xffffffff; selp.b32 %0, 1, 0, p;}" : "=r"(ok));
  return ok != 0;
}

// MODE 0: A K-major in smem; 1: A MN-major in smem (SWIZZLE_128B_BASE32B); 2: A in TMEM
// STYLE 0: one thread (threadIdx.x == 0) runs the issue loop -- what the kernels of this repo do;
// STYLE 1: the whole warp runs the loop convergently and elect.sync picks the issuing lane per MMA (operands warp-uniform)
// The loop body is 16 MMAs over 4 operand blocks x 4 k steps with descriptors built before the clock starts.
template <int MODE, int STYLE>
__global__ void __launch_bounds__(128, 1) k(int N, int nmma, int nacc, long long* out) {
  extern __shared__ uint8_t raw[];
  const uint32_t base = (s32(raw) + 1023u) & ~1023u;
  uint8_t* tiles = raw + (base - s32(raw));
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  for (int i = threadIdx.x; i < (4 * 48 * 1024) / 4; i += blockDim.x)
    reinterpret_cast<float*>(tiles)[i] = (float)((i * 2654435761u) >> 20) * 1e-3f;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s32(&slot)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tm = slot;
  if (STYLE == 0 ? threadIdx.x == 0 : threadIdx.x < 32) {
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((MODE == 1 ? 1u : 0u) << 15) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    uint64_t ad[16], bd[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const uint32_t sa = base + (j >> 2) * 49152, sb = sa + 16384;
      const int kk = j & 3;
      ad[j] = MODE == 0 ? desc(sa + 32 * kk, 16, 1024, 2) : desc(sa + 1024 * kk, 4096, 512, 1);
      bd[j] = desc(sb + 32 * kk, 16, 1024, 2);
    }
    const uint32_t d0 = tm + 64, d1 = tm + 64 + (nacc > 1 ? (uint32_t)N : 0u);
    const long long t0 = clock64();
    for (int i = 0; i < nmma; i += 16) {
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const uint32_t d = (j & 4) ? d1 : d0;
        if (STYLE == 0 || elect_one()) {
          if (MODE == 2) mma_ts(d, tm + 8 * (j & 3), bd[j], idesc, 1u);
          else mma_ss(d, ad[j], bd[j], idesc, 1u);
        }
      }
    }
    const long long t1 = clock64();
    if (STYLE == 0 || elect_one()) commit(&bar);
    mbar_wait(&bar, 0);
    const long long t2 = clock64();
    if (threadIdx.x == 0) {
      out[blockIdx.x * 2] = t1 - t0;
      out[blockIdx.x * 2 + 1] = t2 - t0;
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(512));
}

typedef void (*KernelFn)(int, int, int, long long*);

int main() {
  int dev = 0, nsm = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev);
  const int smem = 4 * 49152 + 2048, nmma = 4096;
  KernelFn fns[3][2] = {{k<0, 0>, k<0, 1>}, {k<1, 0>, k<1, 1>}, {k<2, 0>, k<2, 1>}};
  for (int m = 0; m < 3; ++m)
    for (int st = 0; st < 2; ++st) cudaFuncSetAttribute(fns[m][st], cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  long long* d_out;
  cudaMalloc(&d_out, nsm * 2 * sizeof(long long));
  std::vector<long long> h(nsm * 2);
  const char* names[3] = {"A smem K-major", "A smem MN-major", "A tmem"};
  const char* styles[2] = {"one thread (divergent branch)", "whole warp + elect.sync"};
  printf("{\"bench\": \"umma_issue\", \"kind\": \"tf32 M128 K8 cta_group::1\", \"nmma\": %d, \"ctas\": %d, \"rows\": [\n", nmma, nsm);
  bool first = true;
  for (int st = 0; st < 2; ++st)
    for (int mode = 0; mode < 3; ++mode)
      for (int N : {64, 128, 256})
        for (int nacc : {1, 2}) {
          if (nacc * N > 448) continue;
          for (int rep = 0; rep < 2; ++rep) {   // second launch is the measurement
            fns[mode][st]<<<nsm, 128, smem>>>(N, nmma, nacc, d_out);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("CUDA error %s (mode %d N %d)\n", cudaGetErrorString(e), mode, N); return 1; }
          }
          cudaMemcpy(h.data(), d_out, nsm * 2 * sizeof(long long), cudaMemcpyDeviceToHost);
          double a = 0, b = 0;
          for (int i = 0; i < nsm; ++i) { a += h[2 * i]; b += h[2 * i + 1]; }
          a /= nsm * (double)nmma; b /= nsm * (double)nmma;
          const double ideal = 128.0 * N * 8 / 2048.0;   // 2048 tf32 MACs per clk and SM (dense nominal)
          printf("%s  {\"issue\": \"%s\", \"a\": \"%s\", \"N\": %d, \"accumulators\": %d, \"issue_clk_per_mma\": %.1f, \"clk_per_mma\": %.1f, \"nominal_clk\": %.0f, \"smem_bytes_per_mma\": %d}",
                 first ? "" : ",\n", styles[st], names[mode], N, nacc, a, b, ideal, (mode == 2 ? 0 : 128 * 32) + N * 32);
          first = false;
        }
  printf("\n]}\n");
  return 0;
}
