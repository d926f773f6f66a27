// FiLM modulation, InstanceNorm statistics and the per-plane affine that the two collapse into.
//
// replaces: FiLM.forward (/root/reference MSFNO/Models/sfno/sfnonet.py:689-697; two einops.repeat
// materialisations + 3 elementwise passes), nn.InstanceNorm2d(eps=1e-6, affine=True) as configured at
// sfnonet.py:491-499 (3 passes), and their autograd.  InstanceNorm followed by FiLM is one affine per
// (b, c) plane: y = A*x + S, so the pair costs one read + one write of the tensor.
#include "common.cuh"

namespace msfno {

__global__ void film_fwd_kernel(const float* __restrict__ x, const float* __restrict__ gamma,
                                const float* __restrict__ beta, float scale, float* __restrict__ y, long long HW) {
  const int plane = blockIdx.y;
  const float a = 1.0f + gamma[plane] * scale;
  const float s = beta[plane] * scale;
  const float* xp = x + (size_t)plane * HW;
  float* yp = y + (size_t)plane * HW;
  const bool vec = ((HW & 3) == 0) && (((reinterpret_cast<uintptr_t>(xp) | reinterpret_cast<uintptr_t>(yp)) & 15) == 0);
  if (vec) {
    const long long n4 = HW >> 2;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
      float4 v = __ldcs(reinterpret_cast<const float4*>(xp) + i);
      v.x = fmaf(a, v.x, s); v.y = fmaf(a, v.y, s); v.z = fmaf(a, v.z, s); v.w = fmaf(a, v.w, s);
      __stcs(reinterpret_cast<float4*>(yp) + i, v);
    }
  } else {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < HW; i += (long long)gridDim.x * blockDim.x)
      yp[i] = fmaf(a, xp[i], s);
  }
}

__global__ void plane_affine_kernel(const float* __restrict__ x, const float* __restrict__ A, const float* __restrict__ S,
                                    float* __restrict__ y, long long HW) {
  const int plane = blockIdx.y;
  const float a = A[plane], s = S[plane];
  const float* xp = x + (size_t)plane * HW;
  float* yp = y + (size_t)plane * HW;
  const bool vec = ((HW & 3) == 0) && (((reinterpret_cast<uintptr_t>(xp) | reinterpret_cast<uintptr_t>(yp)) & 15) == 0);
  if (vec) {
    const long long n4 = HW >> 2;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
      float4 v = __ldcs(reinterpret_cast<const float4*>(xp) + i);
      v.x = fmaf(a, v.x, s); v.y = fmaf(a, v.y, s); v.z = fmaf(a, v.z, s); v.w = fmaf(a, v.w, s);
      __stcs(reinterpret_cast<float4*>(yp) + i, v);
    }
  } else {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < HW; i += (long long)gridDim.x * blockDim.x)
      yp[i] = fmaf(a, xp[i], s);
  }
}

__device__ __forceinline__ double block_sum(double v, double* sh) {
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __syncthreads();
  if (lane == 0) sh[warp] = v;
  __syncthreads();
  double t = 0.0;
  if (warp == 0) {
    t = (lane < (blockDim.x >> 5)) ? sh[lane] : 0.0;
    for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
  }
  return t;  // valid in thread 0
}

// grid: (chunks, planes); partial sums accumulated with fp64 atomics into zeroed stats
__global__ void plane_stats_kernel(const float* __restrict__ x, double* __restrict__ stats, long long HW) {
  __shared__ double sh[32];
  const int plane = blockIdx.y;
  const float* xp = x + (size_t)plane * HW;
  float s = 0.f, q = 0.f;
  double ds = 0.0, dq = 0.0;
  const bool vec = ((HW & 3) == 0) && ((reinterpret_cast<uintptr_t>(xp) & 15) == 0);
  int cnt = 0;
  if (vec) {
    const long long n4 = HW >> 2;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(xp) + i);
      s += (v.x + v.y) + (v.z + v.w);
      q += (v.x * v.x + v.y * v.y) + (v.z * v.z + v.w * v.w);
      if (++cnt == 64) { ds += s; dq += q; s = q = 0.f; cnt = 0; }
    }
  } else {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < HW; i += (long long)gridDim.x * blockDim.x) {
      const float v = xp[i];
      s += v; q += v * v;
      if (++cnt == 256) { ds += s; dq += q; s = q = 0.f; cnt = 0; }
    }
  }
  ds += s; dq += q;
  const double ts = block_sum(ds, sh);
  const double tq = block_sum(dq, sh);
  if (threadIdx.x == 0) {
    atomicAdd(&stats[2 * plane], ts);
    atomicAdd(&stats[2 * plane + 1], tq);
  }
}

// Fold a per-plane affine y = A x + S (InstanceNorm o FiLM) into the 1x1 conv that consumes y:
//   Wb[b][o][c] = W[o][c] * A[b][c]   (zero-padded columns stay zero; optionally rounded to TF32)
//   bb[b][o]    = sum_c W[o][c] * S[b][c] + bias[o]
// One launch instead of the five small library kernels the same algebra costs in eager PyTorch.
__global__ void fold_affine_kernel(const float* __restrict__ W, const float* __restrict__ A, const float* __restrict__ S,
                                   const float* __restrict__ bias, float* __restrict__ Wb, float* __restrict__ bb, int O, int C,
                                   int ld, int round_tf32) {
  pdl_trigger();
  pdl_wait();
  const int o = blockIdx.x, b = blockIdx.y;
  const float* w = W + (size_t)o * ld;
  float* wb = Wb + ((size_t)b * O + o) * ld;
  float acc = 0.0f;
  for (int c = threadIdx.x; c < ld; c += blockDim.x) {
    float v = 0.0f;
    if (c < C) {
      const float wv = w[c];
      v = wv * A[(size_t)b * C + c];
      acc = fmaf(wv, S[(size_t)b * C + c], acc);
      if (round_tf32) v = __uint_as_float((__float_as_uint(v) + 0x1000u) & 0xffffe000u);
    }
    wb[c] = v;
  }
  __shared__ float red[32];
  for (int off = 16; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.0f;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t += red[i];
    bb[(size_t)b * O + o] = t + (bias ? bias[o] : 0.0f);
  }
}

// InstanceNorm(eps, affine nw / nb) followed by optional FiLM(gamma, beta, scale) of one plane as y = a x + s, from the
// plane's fp64 (sum, sum of squares).  One definition for norm_film_coeffs_kernel and fold_norm_affine_kernel, so the
// fused launch reproduces the two-launch result bit for bit.
__device__ __forceinline__ void norm_film_coeff(const double* __restrict__ stats, const float* __restrict__ nw,
                                                const float* __restrict__ nb, const float* __restrict__ gamma,
                                                const float* __restrict__ beta, float scale, float eps, double inv_hw, int i,
                                                int c, float& A, float& S) {
  const double mean = stats[2 * i] * inv_hw;
  double var = stats[2 * i + 1] * inv_hw - mean * mean;  // biased variance, as InstanceNorm uses
  if (var < 0.0) var = 0.0;
  const double rstd = 1.0 / sqrt(var + (double)eps);
  double a = rstd * (nw ? (double)nw[c] : 1.0);
  double s = (nb ? (double)nb[c] : 0.0) - mean * a;
  if (gamma) {
    const double f = 1.0 + (double)gamma[i] * (double)scale;
    a *= f;
    s = s * f + (double)beta[i] * (double)scale;
  }
  A = (float)a;
  S = (float)s;
}

// fold_affine_kernel with the coefficients computed in place from the plane statistics (norm_film_coeffs + fold_affine in
// one launch on the InstanceNorm -> FiLM -> fc1 path of every block)
__global__ void fold_norm_affine_kernel(const float* __restrict__ W, const double* __restrict__ stats,
                                        const float* __restrict__ nw, const float* __restrict__ nb,
                                        const float* __restrict__ gamma, const float* __restrict__ beta, float scale, float eps,
                                        double inv_hw, const float* __restrict__ bias, float* __restrict__ Wb,
                                        float* __restrict__ bb, int O, int C, int ld, int round_tf32) {
  pdl_trigger();
  pdl_wait();
  const int o = blockIdx.x, b = blockIdx.y;
  const float* w = W + (size_t)o * ld;
  float* wb = Wb + ((size_t)b * O + o) * ld;
  float acc = 0.0f;
  for (int c = threadIdx.x; c < ld; c += blockDim.x) {
    float v = 0.0f;
    if (c < C) {
      float A, S;
      norm_film_coeff(stats, nw, nb, gamma, beta, scale, eps, inv_hw, b * C + c, c, A, S);
      const float wv = w[c];
      v = wv * A;
      acc = fmaf(wv, S, acc);
      if (round_tf32) v = __uint_as_float((__float_as_uint(v) + 0x1000u) & 0xffffe000u);
    }
    wb[c] = v;
  }
  __shared__ float red[32];
  for (int off = 16; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.0f;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t += red[i];
    bb[(size_t)b * O + o] = t + (bias ? bias[o] : 0.0f);
  }
}

// out[i] = g[i] * gelu'(h[i])  (exact erf GELU): the activation's adjoint in the frozen-weight MLP backward.  RND: the
// product is rounded to TF32 (round-to-nearest) where it is made -- it is the tensor-core operand of the next GEMM, and
// rounding it afterwards cost two more passes over a tensor of 1 GB per sample.
template <bool RND>
__device__ __forceinline__ float gelu_bwd_one(float g, float h) {
  const float v = g * gelu_erf_grad(h);
  return RND ? rna_tf32_dev(v) : v;
}
template <bool RND>
__global__ void gelu_bwd_mul_kernel(const float* __restrict__ g, const float* __restrict__ h, float* __restrict__ out, long long n) {
  const long long n4 = n >> 2;
  const bool vec = (((reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(h) | reinterpret_cast<uintptr_t>(out)) & 15) == 0);
  if (vec) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
      const float4 a = __ldcs(reinterpret_cast<const float4*>(g) + i), b = __ldcs(reinterpret_cast<const float4*>(h) + i);
      __stcs(reinterpret_cast<float4*>(out) + i,
             make_float4(gelu_bwd_one<RND>(a.x, b.x), gelu_bwd_one<RND>(a.y, b.y), gelu_bwd_one<RND>(a.z, b.z), gelu_bwd_one<RND>(a.w, b.w)));
    }
    for (long long i = (n4 << 2) + blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
      out[i] = gelu_bwd_one<RND>(g[i], h[i]);
  } else {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
      out[i] = gelu_bwd_one<RND>(g[i], h[i]);
  }
}

__global__ void norm_film_coeffs_kernel(const double* __restrict__ stats, const float* __restrict__ nw,
                                        const float* __restrict__ nb, const float* __restrict__ gamma,
                                        const float* __restrict__ beta, float scale, float eps, float* __restrict__ A,
                                        float* __restrict__ S, int B, int C, double inv_hw) {
  pdl_trigger();
  pdl_wait();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * C) return;
  float a, s;
  norm_film_coeff(stats, nw, nb, gamma, beta, scale, eps, inv_hw, i, i % C, a, s);
  A[i] = a;
  S[i] = s;
}

// Mean-carrying residual stream (tensor-core tier, sfnonet.py: _fused): everything a block needs from the plane sums of the
// stored stream X = x - mu and the carried offset mu, in one launch (it replaced six PyTorch element-wise / gemv launches
// per block, each of which also broke the programmatic-dependent-launch chain of the step):
//   mean[b][c]   = stats[b C + c][0] / HW
//   b2[b][o]     = bias2[o] - mean[b][o]                        (per-sample bias of the fused MLP's second conv)
//   mu_out[b][c] = mu[b][c] + mean[b][c]                        (offset of the stream the block stores)
//   sb[b][o]     = skip_bias[o] + sum_c Wskip[o][c] mu[b][c]    (per-sample bias of the inner skip conv; only with mu)
// grid = B, block = 256: warps walk the rows of Wskip with lanes along c (coalesced), fixed-order shuffle reduction.
__global__ void mean_carry_kernel(const double* __restrict__ stats, double inv_hw, const float* __restrict__ mu,
                                  const float* __restrict__ bias2, const float* __restrict__ Wskip, long long ldw,
                                  const float* __restrict__ skip_bias, float* __restrict__ b2, float* __restrict__ mu_out,
                                  float* __restrict__ sb, int C) {
  extern __shared__ float mu_s[];
  pdl_trigger();
  pdl_wait();
  const int b = blockIdx.x;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const float m = mu ? mu[(size_t)b * C + c] : 0.0f;
    mu_s[c] = m;
    if (stats) {
      const float mean = (float)(stats[2 * ((size_t)b * C + c)] * inv_hw);
      if (b2) b2[(size_t)b * C + c] = (bias2 ? bias2[c] : 0.0f) - mean;
      if (mu_out) mu_out[(size_t)b * C + c] = m + mean;
    }
  }
  if (!sb) return;
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  for (int o = warp; o < C; o += nw) {
    const float* wr = Wskip + (size_t)o * ldw;
    float acc = 0.0f;
    for (int c = lane; c < C; c += 32) acc = fmaf(wr[c], mu_s[c], acc);
    for (int off = 16; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
    if (lane == 0) sb[(size_t)b * C + o] = (skip_bias ? skip_bias[o] : 0.0f) + acc;
  }
}

// one CTA per plane: gx = (1 + gamma*scale) * gy, ggamma = scale * sum(gy*x), gbeta = scale * sum(gy)
__global__ void film_bwd_kernel(const float* __restrict__ gy, const float* __restrict__ x, const float* __restrict__ gamma,
                                float scale, float* __restrict__ gx, float* __restrict__ ggamma,
                                float* __restrict__ gbeta, long long HW) {
  __shared__ double sh[32];
  const int plane = blockIdx.x;
  const float a = 1.0f + gamma[plane] * scale;
  const float* gp = gy + (size_t)plane * HW;
  const float* xp = x + (size_t)plane * HW;
  float* op = gx ? gx + (size_t)plane * HW : nullptr;
  double dg = 0.0, db = 0.0;
  float sg = 0.f, sb = 0.f;
  int cnt = 0;
  const bool vec = ((HW & 3) == 0) && (((reinterpret_cast<uintptr_t>(gp) | reinterpret_cast<uintptr_t>(xp) | reinterpret_cast<uintptr_t>(op)) & 15) == 0);
  if (vec) {   // 16-byte streaming loads: this pass reads 2 x 8.5 GB at B = 8 (it ran at 1.9 TB/s with 4-byte loads)
    const long long n4 = HW >> 2;
    for (long long i = threadIdx.x; i < n4; i += blockDim.x) {
      const float4 g = __ldcs(reinterpret_cast<const float4*>(gp) + i), xv = __ldcs(reinterpret_cast<const float4*>(xp) + i);
      sg = fmaf(g.x, xv.x, sg); sg = fmaf(g.y, xv.y, sg); sg = fmaf(g.z, xv.z, sg); sg = fmaf(g.w, xv.w, sg);
      sb += (g.x + g.y) + (g.z + g.w);
      if (op) __stcs(reinterpret_cast<float4*>(op) + i, make_float4(a * g.x, a * g.y, a * g.z, a * g.w));
      if (++cnt == 32) { dg += sg; db += sb; sg = sb = 0.f; cnt = 0; }
    }
  } else {
    for (long long i = threadIdx.x; i < HW; i += blockDim.x) {
      const float g = gp[i];
      sg = fmaf(g, xp[i], sg);
      sb += g;
      if (op) op[i] = a * g;
      if (++cnt == 128) { dg += sg; db += sb; sg = sb = 0.f; cnt = 0; }
    }
  }
  dg += sg; db += sb;
  const double tg = block_sum(dg, sh);
  const double tb = block_sum(db, sh);
  if (threadIdx.x == 0) {
    ggamma[plane] = (float)(tg * scale);
    gbeta[plane] = (float)(tb * scale);
  }
}

}  // namespace msfno

using namespace msfno;

static inline int chunks_for(long long HW, int planes) {
  long long per = (HW / 4 + 255) / 256;               // CTAs to cover a plane once with float4 + 256 threads
  long long want = (148LL * 8 + planes - 1) / planes;  // fill the machine ~8 CTAs per SM
  long long c = per < want ? per : want;
  return (int)(c < 1 ? 1 : c);
}

extern "C" {

int msfno_film_affine_fwd(const float* x, const float* gamma, const float* beta, float scale, float* y, int B, int C,
                          long HW, void* stream) {
  if (!x || !gamma || !beta || !y || B < 1 || C < 1 || HW < 1) return record_error(MSFNO_ERR_BAD_SHAPE, "film_fwd: bad argument");
  dim3 grid(chunks_for(HW, B * C), B * C);
  film_fwd_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(x, gamma, beta, scale, y, HW);
  count_launch();
  MSFNO_CUDA_OK(cudaGetLastError());
  return MSFNO_OK;
}

int msfno_film_affine_bwd(const float* gy, const float* x, const float* gamma, float scale, float* gx, float* ggamma,
                          float* gbeta, int B, int C, long HW, void* stream) {
  if (!gy || !x || !gamma || !ggamma || !gbeta || B < 1 || C < 1 || HW < 1) return record_error(MSFNO_ERR_BAD_SHAPE, "film_bwd: bad argument");
  film_bwd_kernel<<<B * C, 1024, 0, (cudaStream_t)stream>>>(gy, x, gamma, scale, gx, ggamma, gbeta, HW);
  count_launch();
  MSFNO_CUDA_OK(cudaGetLastError());
  return MSFNO_OK;
}

int msfno_plane_stats(const float* x, double* stats, int planes, long HW, void* stream) {
  if (!x || !stats || planes < 1 || HW < 1) return record_error(MSFNO_ERR_BAD_SHAPE, "plane_stats: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  MSFNO_CUDA_OK(cudaMemsetAsync(stats, 0, sizeof(double) * 2 * (size_t)planes, st));
  dim3 grid(chunks_for(HW, planes), planes);
  plane_stats_kernel<<<grid, 256, 0, st>>>(x, stats, HW);
  count_launch();
  MSFNO_CUDA_OK(cudaGetLastError());
  return MSFNO_OK;
}

int msfno_norm_film_coeffs(const double* stats, const float* nw, const float* nb, const float* gamma, const float* beta,
                           float scale, float eps, float* A, float* S, int B, int C, long HW, void* stream) {
  if (!stats || !A || !S || B < 1 || C < 1 || HW < 1 || ((gamma == nullptr) != (beta == nullptr)))
    return record_error(MSFNO_ERR_BAD_SHAPE, "norm_film_coeffs: bad argument");
  MSFNO_CUDA_OK(launch_pdl(norm_film_coeffs_kernel, dim3((B * C + 127) / 128), dim3(128), 0, (cudaStream_t)stream, stats, nw, nb, gamma,
                           beta, scale, eps, A, S, B, C, 1.0 / (double)HW));
  count_launch();
  MSFNO_CUDA_OK(cudaGetLastError());
  return MSFNO_OK;
}

int msfno_fold_affine(const float* W, const float* A, const float* S, const float* bias, float* Wb, float* bb, int B, int O, int C,
                      int ld, int round_tf32, void* stream) {
  if (!W || !A || !S || !Wb || !bb || B < 1 || O < 1 || C < 1 || ld < C) return record_error(MSFNO_ERR_BAD_SHAPE, "fold_affine: bad argument");
  MSFNO_CUDA_OK(launch_pdl(fold_affine_kernel, dim3(O, B), dim3(128), 0, (cudaStream_t)stream, W, A, S, bias, Wb, bb, O, C, ld, round_tf32));
  count_launch();
  MSFNO_CUDA_OK(cudaGetLastError());
  return MSFNO_OK;
}

int msfno_fold_norm_affine(const float* W, const double* stats, const float* nw, const float* nb, const float* gamma,
                           const float* beta, float scale, float eps, long HW, const float* bias, float* Wb, float* bb, int B,
                           int O, int C, int ld, int round_tf32, void* stream) {
  if (!W || !stats || !Wb || !bb || B < 1 || O < 1 || C < 1 || ld < C || HW < 1 || ((gamma == nullptr) != (beta == nullptr)))
    return record_error(MSFNO_ERR_BAD_SHAPE, "fold_norm_affine: bad argument");
  MSFNO_CUDA_OK(launch_pdl(fold_norm_affine_kernel, dim3(O, B), dim3(128), 0, (cudaStream_t)stream, W, stats, nw, nb, gamma, beta,
                           scale, eps, 1.0 / (double)HW, bias, Wb, bb, O, C, ld, round_tf32));
  count_launch();
  MSFNO_CUDA_OK(cudaGetLastError());
  return MSFNO_OK;
}

int msfno_mean_carry(const double* stats, long HW, const float* mu, const float* bias2, const float* Wskip, long ldw,
                     const float* skip_bias, float* b2, float* mu_out, float* sb, int B, int C, void* stream) {
  if (B < 1 || C < 1 || (stats && HW < 1) || (!stats && (b2 || mu_out)) || (sb && (!mu || !Wskip || ldw < C)) || (!b2 && !mu_out && !sb))
    return record_error(MSFNO_ERR_BAD_SHAPE, "mean_carry: bad argument");
  MSFNO_CUDA_OK(launch_pdl(mean_carry_kernel, dim3(B), dim3(256), sizeof(float) * (size_t)C, (cudaStream_t)stream, stats,
                           stats ? 1.0 / (double)HW : 0.0, mu, bias2, Wskip, (long long)ldw, skip_bias, b2, mu_out, sb, C));
  count_launch();
  MSFNO_CUDA_OK(cudaGetLastError());
  return MSFNO_OK;
}

int msfno_gelu_bwd_mul(const float* g, const float* h, float* out, long long n, int round_tf32, void* stream) {
  if (!g || !h || !out || n < 1) return record_error(MSFNO_ERR_BAD_SHAPE, "gelu_bwd_mul: bad argument");
  long long blocks = (n / 4 + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  if (blocks < 1) blocks = 1;
  if (round_tf32) gelu_bwd_mul_kernel<true><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(g, h, out, n);
  else gelu_bwd_mul_kernel<false><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(g, h, out, n);
  count_launch();
  MSFNO_CUDA_OK(cudaGetLastError());
  return MSFNO_OK;
}

int msfno_plane_affine(const float* x, const float* A, const float* S, float* y, int planes, long HW, void* stream) {
  if (!x || !A || !S || !y || planes < 1 || HW < 1) return record_error(MSFNO_ERR_BAD_SHAPE, "plane_affine: bad argument");
  dim3 grid(chunks_for(HW, planes), planes);
  plane_affine_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(x, A, S, y, HW);
  count_launch();
  MSFNO_CUDA_OK(cudaGetLastError());
  return MSFNO_OK;
}

}  // extern "C"
