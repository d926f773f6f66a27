// Micro-benchmark: cost of a cp.async.bulk (1-D TMA) request as a function of its size and of the number of
// issuing warps per SM.  One CTA per SM; every warp owns a ring of DEPTH buffers + mbarriers, lane 0 issues.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tma_request_bench tma_request_bench.cu && ./tma_request_bench
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <vector>
#include <cstdlib>

__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, int c) { asm volatile("mbarrier.init.shared::cta.b64 [%1], %0;" ::"r"(c), "r"(s32(b))); }
__device__ __forceinline__ void expect_tx(uint64_t* b, uint32_t n) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%1], %0;" ::"r"(n), "r"(s32(b)) : "memory"); }
__device__ __forceinline__ void wait(uint64_t* b, uint32_t ph) {
  uint32_t ok = 0;
  for (int it = 0; it < (1 << 22) && !ok; ++it)
    asm volatile("{.reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.b32 %0,1,0,p;}" : "=r"(ok) : "r"(s32(b)), "r"(ph) : "memory");
  if (!ok) __trap();
}
__device__ __forceinline__ void bulk(void* dst, const void* src, uint32_t n, uint64_t* b) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(s32(dst)), "l"(src), "r"(n), "r"(s32(b)) : "memory");
}

template <int DEPTH>
__global__ void k(const char* src, size_t src_bytes, uint32_t req, int reqs_per_warp, int split, unsigned long long* sink) {
  extern __shared__ __align__(128) char sm[];
  const int nw = blockDim.x / 32, w = threadIdx.x / 32, lane = threadIdx.x % 32;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sm);            // [nw][DEPTH]
  char* bufs = sm + 1024;                                       // [nw][DEPTH][req]
  if (threadIdx.x == 0) for (int i = 0; i < nw * DEPTH; ++i) mbar_init(&bars[i], 1);
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  __syncthreads();
  if (lane != 0) return;
  const size_t nreq_total = (size_t)gridDim.x * nw * reqs_per_warp;
  const size_t slots = src_bytes / req;
  uint64_t* mybar = bars + w * DEPTH;
  char* mybuf = bufs + (size_t)w * DEPTH * req;
  const uint32_t piece = req / split;
  for (int i = 0; i < reqs_per_warp + DEPTH; ++i) {
    const int st = i % DEPTH;
    if (i >= DEPTH) wait(&mybar[st], ((i / DEPTH) - 1) & 1);
    if (i < reqs_per_warp) {
      // consecutive requests of the whole grid tile the source linearly (like rows of a tensor)
      size_t idx = ((size_t)i * gridDim.x * nw + (size_t)blockIdx.x * nw + w) % slots;
      expect_tx(&mybar[st], req);
      for (int s = 0; s < split; ++s) bulk(mybuf + (size_t)st * req + s * piece, src + idx * req + s * piece, piece, &mybar[st]);
    }
  }
  if (nreq_total == 0) *sink = 1;
}

template <int DEPTH>
static void run(const char* src, size_t big, unsigned long long* sink, int sms, int khz, bool& first) {
  cudaFuncSetAttribute(k<DEPTH>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
  for (size_t footprint : {(size_t)24 << 20, big})
    for (int nw : {1, 2, 4, 8})
      for (uint32_t req : {960u, 2048u, 5760u, 16384u, 32768u})
        for (int split : {1, 8}) {
          if ((size_t)nw * DEPTH * req + 1024 > 220 * 1024) continue;
          if ((req / split) % 16) continue;
          const size_t total = (footprint == big) ? ((size_t)3 << 30) : ((size_t)1 << 30);
          int rpw = (int)(total / ((size_t)sms * nw * req));
          if (rpw < 8) rpw = 8;
          const size_t smem = 1024 + (size_t)nw * DEPTH * req;
          cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
          k<DEPTH><<<sms, nw * 32, smem>>>(src, footprint, req, rpw, split, sink);   // warm (L2 fill for the small footprint)
          cudaEventRecord(e0);
          k<DEPTH><<<sms, nw * 32, smem>>>(src, footprint, req, rpw, split, sink);
          cudaEventRecord(e1);
          cudaError_t err = cudaDeviceSynchronize();
          if (err != cudaSuccess) { printf("CUDA error %s\n", cudaGetErrorString(err)); exit(1); }
          float ms; cudaEventElapsedTime(&ms, e0, e1);
          const double bytes = (double)sms * nw * rpw * req;
          const double clk_per_req_sm = ms * 1e-3 * khz * 1e3 / ((double)nw * rpw * split);
          printf("%s{\"src\": \"%s\", \"warps\": %d, \"req_bytes\": %u, \"split\": %d, \"depth\": %d, \"GBps\": %.0f, \"clk_per_request_per_sm\": %.1f}",
                 first ? "" : ",\n", footprint == big ? "dram" : "l2", nw, req, split, DEPTH, bytes / ms / 1e6, clk_per_req_sm);
          first = false;
        }
}

int main() {
  int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  int khz; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
  const size_t big = (size_t)2 << 30;
  char* src; cudaMalloc(&src, big); cudaMemset(src, 1, big);
  unsigned long long* sink; cudaMalloc(&sink, 8);
  printf("{\"sms\": %d, \"clock_khz\": %d, \"rows\": [\n", sms, khz);
  bool first = true;
  run<4>(src, big, sink, sms, khz, first);
  run<16>(src, big, sink, sms, khz, first);
  printf("\n]}\n");
  return 0;
}
