// Longitude transforms of the SHT: truncated real-to-complex FFT (forward) and zero-padded
// complex-to-real FFT (inverse), one warp per latitude row, rows streamed into shared memory by
// the TMA engine (cp.async.bulk + mbarrier), mixed-radix Stockham stages in shared memory.
//
// replaces: torch.fft.rfft(x, dim=-1, norm="forward") * 2*pi  and
//           torch.fft.irfft(X, n=nlon, dim=-1, norm="forward")
//           inside torch_harmonics RealSHT / InverseRealSHT (call sites
//           /root/reference MSFNO/Models/sfno/layers.py:405,421,629,638).
// Unlike cuFFT it only ever writes / reads the mlim = min(mmax, lmax) orders the Legendre stage
// uses, directly in the [b][m][2c+ri][lat] layout that stage contracts over (lat contiguous).
#include <stdlib.h>

#include "fft_core.cuh"
#include "common.cuh"
#include "plan.h"

namespace msfno {

static constexpr int ROWS_PER_TILE = 32;  // latitude rows per CTA == one 128-byte store segment per (m, ri)
static constexpr int OST = 33;            // padded row length of the staging tile (bank-conflict free)

// ---------------------------------------------------------------------------------------------
// forward: x[bc][lat][nlon] -> Xt[b][m][2c+ri][kpad]
//   X[m] = mscale[m] * ( in_scale * DFT(x)[m] + (m == 0) * nlon * in_shift )
//   zero_imag: force Im X[0] = Im X[nlon/2] = 0 (adjoint of irfft)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256, 1)
rfft_trunc_kernel(const float* __restrict__ x, float* __restrict__ Xt, const cf* __restrict__ g_tw,
                  const cf* __restrict__ g_tw2, const float* __restrict__ mscale, const float* __restrict__ in_scale,
                  const float* __restrict__ in_shift, FftSchedule sched, int nlat, int nlon, int mlim, int kpad, int C,
                  int zero_imag, int use_bulk, int round_tf32) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int H = nlon >> 1;
  const int nw = blockDim.x >> 5;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int bc = blockIdx.y;
  const int b = bc / C, c = bc - b * C;
  const int k0 = blockIdx.x * ROWS_PER_TILE;

  // carve shared memory: [tw H][tw2 mlim+1][ostage 2*mlim*OST][bars 2*nw][per-warp 3 buffers of H cf]
  cf* tw = reinterpret_cast<cf*>(smem_raw);
  cf* tw2 = tw + H;
  float* ostage = reinterpret_cast<float*>(tw2 + (mlim + 1));
  size_t off = (size_t)(reinterpret_cast<unsigned char*>(ostage + 2 * mlim * OST) - smem_raw);
  off = (off + 15) & ~(size_t)15;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw + off);
  off += sizeof(uint64_t) * 2 * nw;
  off = (off + 127) & ~(size_t)127;
  cf* wbuf = reinterpret_cast<cf*>(smem_raw + off) + (size_t)warp * 3 * H;
  cf* inb[2] = {wbuf, wbuf + H};
  cf* work = wbuf + 2 * H;

  for (int i = threadIdx.x; i < H; i += blockDim.x) tw[i] = g_tw[i];
  for (int i = threadIdx.x; i <= mlim; i += blockDim.x) tw2[i] = g_tw2[i];
  if (lane == 0) {
    mbar_init(&bars[2 * warp + 0], 1);
    mbar_init(&bars[2 * warp + 1], 1);
    fence_mbar_init();
  }
  __syncthreads();

  const float sc_in = in_scale ? in_scale[bc] : 1.0f;
  const float sh_in = in_shift ? in_shift[bc] : 0.0f;
  const float* xbase = x + ((size_t)bc * nlat) * nlon;
  const uint32_t row_bytes = (uint32_t)nlon * 4u;
  const int iters = ROWS_PER_TILE / nw;

  auto issue_load = [&](int it) {
    const int r = warp + it * nw;
    const float* src = xbase + (size_t)(k0 + r) * nlon;
    if (use_bulk) {
      if (lane == 0) {
        mbar_arrive_expect_tx(&bars[2 * warp + (it & 1)], row_bytes);
        bulk_g2s(inb[it & 1], src, row_bytes, &bars[2 * warp + (it & 1)]);
      }
    } else {
      float2* dst = reinterpret_cast<float2*>(inb[it & 1]);
      const float2* s2 = reinterpret_cast<const float2*>(src);
      for (int i = lane; i < H; i += 32) dst[i] = s2[i];
    }
  };

  if (k0 + warp < nlat) issue_load(0);
  for (int it = 0; it < iters; ++it) {
    const int r = warp + it * nw;
    if (k0 + r >= nlat) break;  // warp-uniform
    if (it + 1 < iters && k0 + r + nw < nlat) issue_load(it + 1);
    if (use_bulk) mbar_wait(&bars[2 * warp + (it & 1)], (uint32_t)((it >> 1) & 1));
    else __syncwarp();

    cf* a = inb[it & 1];
    const int res = stockham_fft<-1>(a, work, tw, H, sched, lane, 32, WarpSync());
    const cf* Z = res ? work : a;

    for (int m = lane; m < mlim; m += 32) {
      cf X = r2c_split(Z, tw2, H, m);
      const float ms = mscale[m];
      X.x *= ms * sc_in;
      X.y *= ms * sc_in;
      if (m == 0) X.x += ms * sh_in * (float)nlon;
      if (zero_imag && (m == 0 || m == H)) X.y = 0.0f;
      ostage[(2 * m) * OST + r] = X.x;
      ostage[(2 * m + 1) * OST + r] = X.y;
    }
    // the next bulk copy into inb[it&1] (issued in iteration it+1 for row it+2) must be ordered after
    // this warp's generic-proxy accesses to it
    fence_proxy_async();
    __syncwarp();
  }
  __syncthreads();

  // coalesced store: one 128-byte segment (32 latitudes) per (m, ri)
  const int nvalid = min(ROWS_PER_TILE, nlat - k0);
  for (int seg = warp; seg < 2 * mlim; seg += nw) {
    const int m = seg >> 1, ri = seg & 1;
    float v = (lane < nvalid) ? ostage[seg * OST + lane] : 0.0f;
    if (round_tf32) { uint32_t rr; asm("cvt.rna.tf32.f32 %0, %1;\n" : "=r"(rr) : "f"(v)); v = __uint_as_float(rr); }
    Xt[(((size_t)b * mlim + m) * (2 * C) + 2 * c + ri) * kpad + k0 + lane] = v;
  }
}

// ---------------------------------------------------------------------------------------------
// inverse: Yt[b][m][2c+ri][kpad] -> y[bc][lat][nlon]
//   y[j] = sum_{m<mlim} cm * Re( mscale[m] * Y[m] * exp(+2 pi i m j / nlon) ),  cm = 1 for m in {0, nlon/2}, else 2
//   (Im of bins 0 and nlon/2 ignored, exactly like irfft).  Epilogue: + skip, GELU, plane statistics.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256, 1)
irfft_trunc_kernel(const float* __restrict__ Yt, float* __restrict__ y, const cf* __restrict__ g_tw,
                   const cf* __restrict__ g_tw2, const float* __restrict__ mscale, const float* __restrict__ skip,
                   const float* __restrict__ out_scale, double* __restrict__ stats, FftSchedule sched, int nlat,
                   int nlon, int mlim, int kpad, int C, int act_gelu, int vec_ok) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int H = nlon >> 1;
  const int nw = blockDim.x >> 5;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int bc = blockIdx.y;
  const int b = bc / C, c = bc - b * C;
  const int k0 = blockIdx.x * ROWS_PER_TILE;

  cf* tw = reinterpret_cast<cf*>(smem_raw);
  cf* tw2 = tw + H;
  float* istage = reinterpret_cast<float*>(tw2 + (mlim + 1));
  size_t off = (size_t)(reinterpret_cast<unsigned char*>(istage + 2 * mlim * OST) - smem_raw);
  off = (off + 15) & ~(size_t)15;
  double* red = reinterpret_cast<double*>(smem_raw + off);  // [nw][2]
  off += sizeof(double) * 2 * nw;
  off = (off + 127) & ~(size_t)127;
  cf* bufa = reinterpret_cast<cf*>(smem_raw + off) + (size_t)warp * 2 * H;
  cf* bufb = bufa + H;

  for (int i = threadIdx.x; i < H; i += blockDim.x) tw[i] = g_tw[i];
  for (int i = threadIdx.x; i <= mlim; i += blockDim.x) tw2[i] = g_tw2[i];
  const int nvalid = min(ROWS_PER_TILE, nlat - k0);
  for (int seg = warp; seg < 2 * mlim; seg += nw) {
    const int m = seg >> 1, ri = seg & 1;
    float v = 0.0f;
    if (lane < nvalid) v = Yt[(((size_t)b * mlim + m) * (2 * C) + 2 * c + ri) * kpad + k0 + lane] * mscale[m];
    istage[seg * OST + lane] = v;
  }
  __syncthreads();

  const float osc = out_scale ? out_scale[bc] : 1.0f;
  float lsum = 0.0f, lsq = 0.0f;
  const int iters = ROWS_PER_TILE / nw;
  for (int it = 0; it < iters; ++it) {
    const int r = warp + it * nw;
    if (k0 + r >= nlat) break;
    // Hermitian half-spectrum lookup (zero beyond mlim; Im dropped at bins 0 and H)
    auto Xh = [&](int q) -> cf {
      if (q >= mlim) return cf{0.0f, 0.0f};
      cf v{istage[(2 * q) * OST + r], istage[(2 * q + 1) * OST + r]};
      if (q == 0 || q == H) v.y = 0.0f;
      return v;
    };
    for (int k = lane; k < H; k += 32) {
      const int kk = H - k;
      cf out{0.0f, 0.0f};
      if (k < mlim || kk < mlim) {
        cf w;
        if (k <= mlim) w = tw2[k];
        else { w = tw2[kk]; w.x = -w.x; }  // exp(-2 pi i k/N) = -conj(exp(-2 pi i (H-k)/N))
        out = c2r_merge(Xh(k), Xh(kk), w);
      }
      bufa[k] = out;
    }
    __syncwarp();
    const int res = stockham_fft<+1>(bufa, bufb, tw, H, sched, lane, 32, WarpSync());
    const float* row = reinterpret_cast<const float*>(res ? bufb : bufa);

    const size_t goff = ((size_t)bc * nlat + (k0 + r)) * nlon;
    if (vec_ok) {
      const float4* row4 = reinterpret_cast<const float4*>(row);
      const float4* skip4 = skip ? reinterpret_cast<const float4*>(skip + goff) : nullptr;
      float4* y4 = reinterpret_cast<float4*>(y + goff);
      for (int i = lane; i < (nlon >> 2); i += 32) {
        float4 v = row4[i];
        v.x *= osc; v.y *= osc; v.z *= osc; v.w *= osc;
        if (skip4) { const float4 s = skip4[i]; v.x += s.x; v.y += s.y; v.z += s.z; v.w += s.w; }
        if (act_gelu & 1) { v.x = gelu_erf(v.x); v.y = gelu_erf(v.y); v.z = gelu_erf(v.z); v.w = gelu_erf(v.w); }
        if (act_gelu & 2) { v.x = rna_tf32_dev(v.x); v.y = rna_tf32_dev(v.y); v.z = rna_tf32_dev(v.z); v.w = rna_tf32_dev(v.w); }
        lsum += (v.x + v.y) + (v.z + v.w);
        lsq += (v.x * v.x + v.y * v.y) + (v.z * v.z + v.w * v.w);
        y4[i] = v;
      }
    } else {
      for (int i = lane; i < nlon; i += 32) {
        float v = row[i] * osc;
        if (skip) v += skip[goff + i];
        if (act_gelu & 1) v = gelu_erf(v);
        if (act_gelu & 2) v = rna_tf32_dev(v);
        lsum += v;
        lsq += v * v;
        y[goff + i] = v;
      }
    }
    __syncwarp();  // all lanes done reading the result before the next row overwrites bufa
  }

  if (stats) {
    double ds = (double)lsum, dq = (double)lsq;
    for (int o = 16; o > 0; o >>= 1) {
      ds += __shfl_xor_sync(0xffffffffu, ds, o);
      dq += __shfl_xor_sync(0xffffffffu, dq, o);
    }
    if (lane == 0) { red[2 * warp] = ds; red[2 * warp + 1] = dq; }
    __syncthreads();
    if (threadIdx.x == 0) {
      double s = 0.0, q = 0.0;
      for (int w = 0; w < nw; ++w) { s += red[2 * w]; q += red[2 * w + 1]; }
      atomicAdd(&stats[2 * bc], s);
      atomicAdd(&stats[2 * bc + 1], q);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// MSFNO_FFT_GENERIC=1 forces the radix-stage kernels (A/B comparison and debugging)
static bool force_generic_fft() {
  static const bool v = dbg_env("MSFNO_FFT_GENERIC");
  return v;
}

static int pick_warps(size_t fixed_bytes, size_t per_warp_bytes, int* nwarps, size_t* total) {
  for (int nw = 8; nw >= 1; nw >>= 1) {
    size_t t = fixed_bytes + 256 + nw * (per_warp_bytes + 16 * 2);
    if (t <= 227 * 1024) { *nwarps = nw; *total = t; return MSFNO_OK; }
  }
  return record_error(MSFNO_ERR_UNSUPPORTED, "longitude FFT does not fit in shared memory (nlon too large)");
}

int launch_rfft_trunc(const msfno_plan* p, const float* x, float* Xt, const float* mscale, int zero_imag,
                      const float* in_scale, const float* in_shift, int B, int C, cudaStream_t st) {
  const int H = p->nlon / 2;
  size_t fixed = sizeof(cf) * (H + p->mlim + 1) + sizeof(float) * 2 * p->mlim * OST;
  int nw; size_t smem;
  int rc = pick_warps(fixed, sizeof(cf) * 3 * H, &nw, &smem);
  if (rc) return rc;
  MSFNO_CUDA_OK(cudaFuncSetAttribute(rfft_trunc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int use_bulk = (p->nlon % 4 == 0) && ((reinterpret_cast<uintptr_t>(x) & 15) == 0);
  if (use_bulk && fft2d_supported(p->nlon) && !force_generic_fft())
    return launch_rfft2d(p, x, Xt, mscale, zero_imag, in_scale, in_shift, B, C, st);
  dim3 grid((p->nlat + ROWS_PER_TILE - 1) / ROWS_PER_TILE, B * C);
  rfft_trunc_kernel<<<grid, nw * 32, smem, st>>>(x, Xt, reinterpret_cast<const cf*>(p->d_tw),
                                                 reinterpret_cast<const cf*>(p->d_tw2), mscale, in_scale, in_shift,
                                                 p->sched, p->nlat, p->nlon, p->mlim, p->kpad, C, zero_imag, use_bulk,
                                                 (p->precision == MSFNO_PREC_TF32 && !zero_imag) ? 1 : 0);
  count_launch();
  MSFNO_CUDA_OK(cudaGetLastError());
  return MSFNO_OK;
}

int launch_irfft_trunc(const msfno_plan* p, const float* Yt, float* y, const float* mscale, const float* skip,
                       const float* out_scale, int act_gelu, double* stats, int B, int C, cudaStream_t st) {
  const int H = p->nlon / 2;
  size_t fixed = sizeof(cf) * (H + p->mlim + 1) + sizeof(float) * 2 * p->mlim * OST + 16 * 8;
  int nw; size_t smem;
  int rc = pick_warps(fixed, sizeof(cf) * 2 * H, &nw, &smem);
  if (rc) return rc;
  MSFNO_CUDA_OK(cudaFuncSetAttribute(irfft_trunc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int vec_ok = (p->nlon % 4 == 0) && ((reinterpret_cast<uintptr_t>(y) & 15) == 0) &&
                     (!skip || (reinterpret_cast<uintptr_t>(skip) & 15) == 0);
  if (vec_ok && fft2d_supported(p->nlon) && !force_generic_fft())
    return launch_irfft2d(p, Yt, y, mscale, skip, out_scale, act_gelu, stats, B, C, st);
  dim3 grid((p->nlat + ROWS_PER_TILE - 1) / ROWS_PER_TILE, B * C);
  irfft_trunc_kernel<<<grid, nw * 32, smem, st>>>(Yt, y, reinterpret_cast<const cf*>(p->d_tw),
                                                  reinterpret_cast<const cf*>(p->d_tw2), mscale, skip, out_scale, stats,
                                                  p->sched, p->nlat, p->nlon, p->mlim, p->kpad, C, act_gelu, vec_ok);
  count_launch();
  MSFNO_CUDA_OK(cudaGetLastError());
  return MSFNO_OK;
}

}  // namespace msfno
