// Shared device helpers (PTX wrappers for mbarrier / bulk async copy) and host-side error plumbing.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

#include <mutex>

#include "../../include/msfno_b200.h"

namespace msfno {

#define MSFNO_CUDA_OK(expr)                                  \
  do {                                                       \
    cudaError_t _e = (expr);                                 \
    if (_e != cudaSuccess) return msfno::record_cuda_error(_e, __FILE__, __LINE__); \
  } while (0)

// Experiment switches read from the environment exist only in builds with -DMSFNO_DEBUG_SWITCHES (csrc/build.sh:
// MSFNO_EXTRA_FLAGS); the product build compiles every one of them to its default.
#ifdef MSFNO_DEBUG_SWITCHES
inline bool dbg_env(const char* name) { return getenv(name) != nullptr; }
inline int dbg_env_int(const char* name, int dflt) { const char* e = getenv(name); return e ? atoi(e) : dflt; }
#else
constexpr bool dbg_env(const char*) { return false; }
constexpr int dbg_env_int(const char*, int dflt) { return dflt; }
#endif

// One-time set-up that is PER DEVICE (cudaFuncSetAttribute applies to the current device's copy of a kernel): a process that
// drives several GPUs -- a module on cuda:1 while cuda:0 is current, tests/test_gpu_serving.py -- must opt every device in,
// which a plain std::call_once does not.
struct PerDeviceOnce {
  std::mutex mu;
  bool done[64] = {};
  cudaError_t err[64] = {};
  template <typename F>
  cudaError_t run(F&& f) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (dev < 0 || dev >= 64) return f();
    std::lock_guard<std::mutex> lk(mu);
    if (!done[dev]) {
      err[dev] = f();
      done[dev] = true;
    }
    return err[dev];
  }
};

int record_cuda_error(cudaError_t e, const char* file, int line);
int record_error(int code, const char* msg);
void count_launch(int n = 1);  // kernels of this library launched so far (msfno_launch_count)
bool pdl_enabled();            // false when MSFNO_NO_PDL is set

#if defined(__CUDACC__)
// Launch `kern` allowing it to overlap the tail of the previous kernel in the stream (the kernel must call pdl_wait()
// before it touches anything its predecessor wrote).
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl_cluster(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, int cluster_x,
                                      Args&&... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
  attr[1].id = cudaLaunchAttributeClusterDimension;
  attr[1].val.clusterDim.x = cluster_x > 1 ? cluster_x : 1;
  attr[1].val.clusterDim.y = 1;
  attr[1].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = cluster_x > 1 ? 2 : 1;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  return launch_pdl_cluster(kern, grid, block, smem, st, 1, static_cast<Args&&>(args)...);
}
#endif

#if defined(__CUDACC__)
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%1], %0;\n" ::"r"(count), "r"(smem_u32(bar)));
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%1], %0;\n" ::"r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra DONE;\n"
      "bra WAIT_LOOP;\n"
      "DONE:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
// 1-D bulk async copy global -> shared (TMA engine, SASS UBLKCP); bytes % 16 == 0, both 16B aligned.
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(
                   smem_u32(smem_dst)),
               "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
// Programmatic dependent launch (PDL): a kernel launched with the programmatic-serialization attribute may start while
// its predecessor in the stream is still running.  pdl_trigger() lets the NEXT kernel's CTAs be scheduled as soon as
// every CTA of this kernel has started; pdl_wait() blocks until the PREVIOUS kernel has completed and its writes are
// visible -- everything before it (barrier init, TMEM allocation, descriptor prefetch) overlaps the predecessor's tail.
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;\n" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;\n" ::: "memory"); }

__device__ __forceinline__ float rna_tf32_dev(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;\n" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}
__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f)); }
// GELU with erf from Abramowitz-Stegun 7.1.26 (|erf error| <= 1.5e-7 absolute): 2 MUFU + ~12 FMA instead of the ~40
// instructions of erff.  Used by the tensor-core kernels, whose tier tolerance (2e-3) it undercuts by four orders.
__device__ __forceinline__ float gelu_fast(float x) {
  // 0.5 x (1 + erf(x / sqrt 2)) = hx + |hx| * erf(|x| / sqrt 2),  hx = x / 2
  const float hx = 0.5f * x;
  const float z = fabsf(x) * 0.70710678118654752440f;
  const float t = __fdividef(1.0f, fmaf(0.3275911f, z, 1.0f));          // MUFU.RCP
  float poly = fmaf(1.061405429f, t, -1.453152027f);
  poly = fmaf(poly, t, 1.421413741f);
  poly = fmaf(poly, t, -0.284496736f);
  poly = fmaf(poly, t, 0.254829592f);
  const float ex = exp2f(z * z * -1.44269504088896340736f);              // MUFU.EX2 (e^{-z^2})
  const float e = fmaf(-poly * t, ex, 1.0f);                             // erf(z)
  return fmaf(fabsf(hx), e, hx);
}
// GELU for the tensor-core tier epilogues: erf(x / sqrt 2) = tanh(x (c0 + c1 x^2 + c2 x^4)) fitted on [-8, 8]
// (max |gelu error| 3e-5 absolute, an eighth of a TF32 rounding step at |y| ~ 1; beyond |x| = 8 the clamp keeps the
// polynomial positive and tanh saturates to the exact limit).  8 instructions with one MUFU.TANH: the fused conv / MLP
// epilogues are issue-bound, gelu_fast alone was half of their instruction count.
__device__ __forceinline__ float gelu_tanh3(float x) {
  const float x2 = fminf(x * x, 64.0f);
  float q = fmaf(x2, -0.00035873236f, 0.0370503451f);
  q = fmaf(x2, q, 0.79745847f);
  float th;
  asm("tanh.approx.f32 %0, %1;\n" : "=f"(th) : "f"(x * q));
  const float hx = 0.5f * x;
  return fmaf(hx, th, hx);
}
// d/dx of exact GELU
__device__ __forceinline__ float gelu_erf_grad(float x) {
  const float cdf = 0.5f * (1.0f + erff(x * 0.70710678118654752440f));
  const float pdf = 0.39894228040143267794f * __expf(-0.5f * x * x);
  return cdf + x * pdf;
}
#endif

}  // namespace msfno
