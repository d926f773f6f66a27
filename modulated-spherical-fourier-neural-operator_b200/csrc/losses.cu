// Latitude-weighted squared-error reductions of the spherical losses, forward and backward.
//
// replaces: L2Sphere / L2Sphere_noSine / CosineMSELoss.forward (/root/reference MSFNO/Models/losses.py:6-37,80-155):
//   loss[b,c] = sum_{h,w} wlat[h] (prd - tar)^2   and   norm[b,c] = sum_{h,w} wlat[h] tar^2
// which the reference computes with six elementwise passes and four [B,C,H,W] temporaries per call, after rebuilding
// the quadrature weights on the HOST every call (losses.py:90,129: numpy leggauss + a host-to-device copy).  Here the
// two tensors are read once, the weights are a cached device vector, the plane sums are fp64.
#include "common.cuh"

namespace msfno {

__device__ __forceinline__ double loss_block_sum(double v, double* sh) {
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

// grid (row chunks, planes): a CTA walks whole latitude rows (one weight per row); out[plane] += (sum w d^2, sum w t^2)
__global__ void weighted_sq_sums_kernel(const float* __restrict__ prd, const float* __restrict__ tar, const float* __restrict__ wlat,
                                        double* __restrict__ out, int H, int W) {
  __shared__ double sh[32];
  const int plane = blockIdx.y;
  const float* pp = prd + (size_t)plane * H * W;
  const float* tp = tar + (size_t)plane * H * W;
  const bool vec = ((W & 3) == 0) && (((reinterpret_cast<uintptr_t>(pp) | reinterpret_cast<uintptr_t>(tp)) & 15) == 0);
  double dd = 0.0, dt = 0.0;
  for (int h = blockIdx.x; h < H; h += gridDim.x) {
    const float w = wlat[h];
    const float* pr = pp + (size_t)h * W;
    const float* tr = tp + (size_t)h * W;
    float sd = 0.f, st = 0.f;
    if (vec) {
      for (int i = threadIdx.x; i < (W >> 2); i += blockDim.x) {
        const float4 a = __ldg(reinterpret_cast<const float4*>(pr) + i);
        const float4 b = __ldg(reinterpret_cast<const float4*>(tr) + i);
        const float d0 = a.x - b.x, d1 = a.y - b.y, d2 = a.z - b.z, d3 = a.w - b.w;
        sd += (d0 * d0 + d1 * d1) + (d2 * d2 + d3 * d3);
        st += (b.x * b.x + b.y * b.y) + (b.z * b.z + b.w * b.w);
      }
    } else {
      for (int i = threadIdx.x; i < W; i += blockDim.x) {
        const float d = pr[i] - tr[i];
        sd += d * d;
        st += tr[i] * tr[i];
      }
    }
    dd += (double)w * sd;   // at most W / blockDim.x terms were summed in fp32
    dt += (double)w * st;
  }
  const double td = loss_block_sum(dd, sh);
  const double tt = loss_block_sum(dt, sh);
  if (threadIdx.x == 0) {
    atomicAdd(&out[2 * plane], td);
    atomicAdd(&out[2 * plane + 1], tt);
  }
}

// gprd[plane][h][w] = coef[plane] * wlat[h] * (prd - tar)      (the caller folds 2, 1/norm, 1/(2 sqrt) and the upstream
// gradient into coef)
__global__ void weighted_diff_kernel(const float* __restrict__ prd, const float* __restrict__ tar, const float* __restrict__ wlat,
                                     const float* __restrict__ coef, float* __restrict__ gprd, int H, int W) {
  const int plane = blockIdx.y;
  const float c = coef[plane];
  const size_t off = (size_t)plane * H * W;
  const bool vec = ((W & 3) == 0) && (((reinterpret_cast<uintptr_t>(prd + off) | reinterpret_cast<uintptr_t>(tar + off) |
                                        reinterpret_cast<uintptr_t>(gprd + off)) & 15) == 0);
  for (int h = blockIdx.x; h < H; h += gridDim.x) {
    const float cw = c * wlat[h];
    const size_t r = off + (size_t)h * W;
    if (vec) {
      for (int i = threadIdx.x; i < (W >> 2); i += blockDim.x) {
        const float4 a = __ldg(reinterpret_cast<const float4*>(prd + r) + i);
        const float4 b = __ldg(reinterpret_cast<const float4*>(tar + r) + i);
        reinterpret_cast<float4*>(gprd + r)[i] = make_float4(cw * (a.x - b.x), cw * (a.y - b.y), cw * (a.z - b.z), cw * (a.w - b.w));
      }
    } else {
      for (int i = threadIdx.x; i < W; i += blockDim.x) gprd[r + i] = cw * (prd[r + i] - tar[r + i]);
    }
  }
}

}  // namespace msfno

using namespace msfno;

static inline int row_chunks(int H, int planes) {
  int want = (148 * 8 + planes - 1) / planes;   // ~8 CTAs per SM over the whole launch
  if (want > H) want = H;
  return want < 1 ? 1 : want;
}

extern "C" {

int msfno_weighted_sq_sums(const float* prd, const float* tar, const float* wlat, double* out, int planes, int H, int W,
                           void* stream) {
  if (!prd || !tar || !wlat || !out || planes < 1 || H < 1 || W < 1) return record_error(MSFNO_ERR_BAD_SHAPE, "weighted_sq_sums: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  MSFNO_CUDA_OK(cudaMemsetAsync(out, 0, sizeof(double) * 2 * (size_t)planes, st));
  weighted_sq_sums_kernel<<<dim3(row_chunks(H, planes), planes), 256, 0, st>>>(prd, tar, wlat, out, H, W);
  count_launch();
  MSFNO_CUDA_OK(cudaGetLastError());
  return MSFNO_OK;
}

int msfno_weighted_diff(const float* prd, const float* tar, const float* wlat, const float* coef, float* gprd, int planes, int H,
                        int W, void* stream) {
  if (!prd || !tar || !wlat || !coef || !gprd || planes < 1 || H < 1 || W < 1)
    return record_error(MSFNO_ERR_BAD_SHAPE, "weighted_diff: bad argument");
  weighted_diff_kernel<<<dim3(row_chunks(H, planes), planes), 256, 0, (cudaStream_t)stream>>>(prd, tar, wlat, coef, gprd, H, W);
  count_launch();
  MSFNO_CUDA_OK(cudaGetLastError());
  return MSFNO_OK;
}

}  // extern "C"
