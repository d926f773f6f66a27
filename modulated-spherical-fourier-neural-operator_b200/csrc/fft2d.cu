// Four-step longitude FFT kernels (forward truncated r2c, inverse zero-padded c2r) for the production sizes:
//   nlon = 1440 (H = 720 = 24 x 30), nlon = 240 (H = 120 = 15 x 8, 4 rows per warp), nlon = 2880 (H = 1440 = 48 x 30).
// Same interface, data layout and fused prologue/epilogue as the generic Stockham kernels in fft.cu (which remain
// the fallback for every other size); the difference is the arithmetic core: two rounds of fully unrolled
// in-register DFTs (fft_reg.cuh) with compile-time twiddles, one shared-memory exchange in between, no integer
// division, bank-conflict-free pitches.  ~4x fewer issued instructions per row than the radix-stage kernel.
//
// replaces: torch.fft.rfft(norm="forward") * 2 pi and torch.fft.irfft(n=nlon, norm="forward") inside
// torch_harmonics RealSHT / InverseRealSHT (/root/reference MSFNO/Models/sfno/layers.py:405,421,629,638).
#include "common.cuh"
#include "fft2d_core.cuh"
#include "plan.h"

namespace msfno {

// RT rows (latitudes) per CTA tile: 32 (one 128-byte segment of the lat-contiguous intermediate per order and re/im) for
// the production grids.  The 2880-point rows of the 0.125 degree grid need 23 KB of shared memory per warp and a
// [2 mlim][RT + 1] staging tile: with RT = 32, twiddles in shared memory and a double-buffered row only TWO warps fit in
// an SM (measured: 9.6 ms per 1441 x 2880 x 256 transform, 0.5 TB/s) -- RT = 16, TWG (the w_H twiddles read through L1
// instead of a shared copy) and a single row buffer per warp that also holds the compact spectrum (!DB) let eight warps in.

// NZ2 > 0: mlim <= NZ2 * P1, so only output columns k2 in [0, NZ2) and [P2 - NZ2, P2) of the second step are ever
// needed by the real split -- a compile-time set: the other outputs of the in-register DFT are dead code.
template <int P1, int P2, int RW, int NZ2, int RT, bool TWG, bool DB>
__global__ void __launch_bounds__(256, 1)
rfft2d_kernel(const float* __restrict__ x, float* __restrict__ Xt, const cf* __restrict__ g_tw,
              const cf* __restrict__ g_tw2, const float* __restrict__ mscale, const float* __restrict__ in_scale,
              const float* __restrict__ in_shift, int nlat, int mlim, int kpad, int C, int zero_imag, int round_tf32,
              const PeerMapDev pm) {
  constexpr int H = P1 * P2, NLON = 2 * H, WP = WorkPitch<P2>::value;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int nw = blockDim.x >> 5;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int bc = blockIdx.y;
  const int b = bc / C, c = bc - b * C;
  constexpr int OST2 = RT + 1;
  const int k0 = blockIdx.x * RT;
  constexpr int KEEP_STATIC = NZ2 * P1;                 // bins [0, KEEP) and [H - KEEP, H) are stored
  const int keep = (NZ2 > 0) ? KEEP_STATIC : mlim;
  const int XS = xs_size(H, keep);

  cf* tw_s = reinterpret_cast<cf*>(smem_raw);
  const cf* tw = TWG ? g_tw : tw_s;
  cf* tw2 = tw_s + (TWG ? 0 : H);
  float* ostage = reinterpret_cast<float*>(tw2 + (mlim + 1));
  size_t off = (size_t)(reinterpret_cast<unsigned char*>(ostage + 2 * mlim * OST2) - smem_raw);
  off = (off + 15) & ~(size_t)15;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw + off);
  off += sizeof(uint64_t) * 2 * nw;
  off = (off + 127) & ~(size_t)127;
  constexpr int NRAW = DB ? 2 : 1;
  // !DB: the compact spectrum store xs re-uses the row buffer (dead once step 1 has read it; XS <= H)
  const size_t per_warp = sizeof(cf) * ((size_t)NRAW * RW * H + (size_t)RW * P1 * WP + (DB ? (size_t)RW * XS : 0));
  unsigned char* wbase = smem_raw + off + (size_t)warp * ((per_warp + 127) & ~(size_t)127);
  cf* raw[2] = {reinterpret_cast<cf*>(wbase), reinterpret_cast<cf*>(wbase) + (DB ? RW * H : 0)};
  cf* work = reinterpret_cast<cf*>(wbase) + NRAW * RW * H;
  cf* xs = DB ? work + RW * P1 * WP : raw[0];

  if (!TWG)
    for (int i = threadIdx.x; i < H; i += blockDim.x) tw_s[i] = g_tw[i];
  for (int i = threadIdx.x; i <= mlim; i += blockDim.x) tw2[i] = g_tw2[i];
  if (lane == 0) {
    mbar_init(&bars[2 * warp + 0], 1);
    mbar_init(&bars[2 * warp + 1], 1);
    fence_mbar_init();
  }
  __syncthreads();

  const float sc_in = in_scale ? in_scale[bc] : 1.0f;
  const float sh_in = in_shift ? in_shift[bc] : 0.0f;
  const float* xbase = x + ((size_t)bc * nlat) * NLON;
  constexpr int NGROUPS = RT / RW;
  const int iters = (NGROUPS + nw - 1) / nw;

  auto rows_valid = [&](int it) -> int {  // valid rows of the row group this warp handles in iteration `it`
    const int g = warp + it * nw;
    if (g >= NGROUPS) return 0;
    const int r0 = k0 + g * RW;
    const int nv = nlat - r0;
    return nv < 0 ? 0 : (nv > RW ? RW : nv);
  };
  auto issue_load = [&](int it, int nv) {
    if (lane == 0) {
      const int g = warp + it * nw;
      const uint32_t bytes = (uint32_t)nv * NLON * 4u;
      mbar_arrive_expect_tx(&bars[2 * warp + (it & 1)], bytes);
      bulk_g2s(raw[it & 1], xbase + (size_t)(k0 + g * RW) * NLON, bytes, &bars[2 * warp + (it & 1)]);
    }
  };

  int nv = rows_valid(0);
  if (nv > 0) issue_load(0, nv);
  for (int it = 0; it < iters; ++it) {
    if (nv <= 0) break;  // warp-uniform; groups are ascending so later ones are invalid too
    const int nv_next = (it + 1 < iters) ? rows_valid(it + 1) : 0;
    if (DB && nv_next > 0) issue_load(it + 1, nv_next);
    mbar_wait(&bars[2 * warp + (it & 1)], (uint32_t)((it >> 1) & 1));
    const cf* in = raw[it & 1];
    const int g = warp + it * nw;

    // step 1: RW*P2 column transforms of length P1
    for (int t = lane; t < RW * P2; t += 32) {
      const int r = t / P2, n2 = t - r * P2;
      if (r < nv) fft2d_step1<P1, P2, -1>(in + r * H, P2, work + r * P1 * WP, tw, n2);
    }
    __syncwarp();
    // step 2: RW*P1 row transforms of length P2; keep only the bins the real split needs
    for (int t = lane; t < RW * P1; t += 32) {
      const int r = t / P1, k1 = t - r * P1;
      if (r < nv) {
        cf v[P2];
        fft2d_step2<P1, P2, -1>(work + r * P1 * WP, k1, v);
        cf* xr = xs + r * XS;
        static_for<0, P2>([&](auto cc) {
          constexpr int k2 = decltype(cc)::value;
          if constexpr (NZ2 > 0) {
            if constexpr (k2 < NZ2) xr[k1 + P1 * k2] = v[k2];
            else if constexpr (k2 >= P2 - NZ2) xr[k1 + P1 * k2 - (H - 2 * KEEP_STATIC)] = v[k2];
          } else {
            const int xi = xs_index(k1 + P1 * k2, H, mlim);
            if (xi >= 0) xr[xi] = v[k2];
          }
        });
      }
    }
    __syncwarp();
    // real split -> staging tile
    for (int r = 0; r < nv; ++r) {
      const int rr = g * RW + r;  // row inside the 32-row tile
      for (int m = lane; m < mlim; m += 32) {
        cf X = r2c_split_xs(xs + r * XS, tw2, H, keep, m);
        const float ms = mscale[m];
        X.x *= ms * sc_in;
        X.y *= ms * sc_in;
        if (m == 0) X.x += ms * sh_in * (float)NLON;
        if (zero_imag && (m == 0 || m == H)) X.y = 0.0f;
        ostage[(2 * m) * OST2 + rr] = X.x;
        ostage[(2 * m + 1) * OST2 + rr] = X.y;
      }
    }
    fence_proxy_async();
    __syncwarp();
    if (!DB && nv_next > 0) issue_load(it + 1, nv_next);   // single row buffer (also the spectrum store): free only now
    nv = nv_next;
  }
  __syncthreads();

  const int nvalid = min(RT, nlat - k0);
  for (int seg = warp; seg < 2 * mlim; seg += nw) {
    const int m = seg >> 1, ri = seg & 1;
    if (lane >= RT) continue;   // (RT < 32: half-segment stores)
    float v = (lane < nvalid) ? ostage[seg * OST2 + lane] : 0.0f;
    if (round_tf32) { uint32_t rr; asm("cvt.rna.tf32.f32 %0, %1;\n" : "=r"(rr) : "f"(v)); v = __uint_as_float(rr); }
    if (pm.world) {
      // sharded transform: the segment goes straight into the operand buffer of the rank that owns order m (an NVLink
      // store through its CUDA IPC mapping) -- the lat<->m exchange is this kernel's epilogue, not a collective
      int s = 0;
      while (s + 1 < pm.world && m >= pm.mb[s + 1]) ++s;
      if (lane < nvalid) pm.buf[s][((size_t)(m - pm.mb[s]) * (2 * C) + 2 * c + ri) * pm.pitch + pm.lat_lo + k0 + lane] = v;
    } else {
      Xt[(((size_t)b * mlim + m) * (2 * C) + 2 * c + ri) * kpad + k0 + lane] = v;
    }
  }
}

// NZ > 0: only the first and last NZ rows of the P1 x P2 spectrum matrix can be non-zero (mlim <= NZ * P2).
template <int P1, int P2, int RW, int NZ, int RT, bool TWG>
__global__ void __launch_bounds__(256, 1)
irfft2d_kernel(const float* __restrict__ Yt, float* __restrict__ y, const cf* __restrict__ g_tw,
               const cf* __restrict__ g_tw2, const float* __restrict__ mscale, const float* __restrict__ skip,
               const float* __restrict__ out_scale, double* __restrict__ stats, int nlat, int mlim, int kpad, int C,
               int act_gelu, const PeerMapDev pm) {
  constexpr int H = P1 * P2, NLON = 2 * H, WP = WorkPitch<P2>::value;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int nw = blockDim.x >> 5;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int bc = blockIdx.y;
  const int b = bc / C, c = bc - b * C;
  constexpr int OST2 = RT + 1;
  const int k0 = blockIdx.x * RT;

  cf* tw_s = reinterpret_cast<cf*>(smem_raw);
  const cf* tw = TWG ? g_tw : tw_s;
  cf* tw2 = tw_s + (TWG ? 0 : H);
  float* istage = reinterpret_cast<float*>(tw2 + (mlim + 1));
  size_t off = (size_t)(reinterpret_cast<unsigned char*>(istage + 2 * mlim * OST2) - smem_raw);
  off = (off + 15) & ~(size_t)15;
  double* red = reinterpret_cast<double*>(smem_raw + off);
  off += sizeof(double) * 2 * nw;
  float* msc = reinterpret_cast<float*>(smem_raw + off);   // per-order scale
  off += sizeof(float) * mlim;
  off = (off + 127) & ~(size_t)127;
  const size_t per_warp = sizeof(cf) * ((size_t)RW * P1 * WP + (size_t)RW * H);
  unsigned char* wbase = smem_raw + off + (size_t)warp * ((per_warp + 127) & ~(size_t)127);
  cf* work = reinterpret_cast<cf*>(wbase);
  cf* outb = work + RW * P1 * WP;

  if (!TWG)
    for (int i = threadIdx.x; i < H; i += blockDim.x) tw_s[i] = g_tw[i];
  for (int i = threadIdx.x; i <= mlim; i += blockDim.x) tw2[i] = g_tw2[i];
  const int nvalid = min(RT, nlat - k0);
  // staging fill: every 128-byte segment is fetched with fire-and-forget cp.async (LDGSTS), so all ~2*mlim/nw loads
  // of a lane are in flight at once instead of one DRAM round trip per batch (the profile showed half of the kernel's
  // stall samples on the first use of these loads); the per-order scale is applied when the spectrum is read
  for (int seg = warp; seg < 2 * mlim; seg += nw) {
    const int m = seg >> 1, ri = seg & 1;
    if (lane >= RT) continue;
    if (lane < nvalid) {
      const float* src;
      if (pm.world) {   // sharded transform: order m is read from its owner's buffer (NVLink load through the IPC mapping)
        int s = 0;
        while (s + 1 < pm.world && m >= pm.mb[s + 1]) ++s;
        src = pm.buf[s] + ((size_t)(m - pm.mb[s]) * (2 * C) + 2 * c + ri) * pm.pitch + pm.lat_lo + k0 + lane;
      } else {
        src = Yt + (((size_t)b * mlim + m) * (2 * C) + 2 * c + ri) * kpad + k0 + lane;
      }
      asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(smem_u32(istage + seg * OST2 + lane)), "l"(src) : "memory");
    } else {
      istage[seg * OST2 + lane] = 0.0f;
    }
  }
  for (int i = threadIdx.x; i < mlim; i += blockDim.x) msc[i] = mscale[i];
  asm volatile("cp.async.wait_all;\n" ::: "memory");
  __syncthreads();

  const float osc = out_scale ? out_scale[bc] : 1.0f;
  float lsum = 0.0f, lsq = 0.0f;
  constexpr int NGROUPS = RT / RW;
  const int iters = (NGROUPS + nw - 1) / nw;
  for (int it = 0; it < iters; ++it) {
    const int g = warp + it * nw;
    if (g >= NGROUPS) break;
    int nv = nlat - (k0 + g * RW);
    if (nv <= 0) break;
    if (nv > RW) nv = RW;

    // Hermitian merge into the [n1][n2] work matrix (zero beyond mlim; Im dropped at bins 0 and H)
    for (int r = 0; r < nv; ++r) {
      const int rr = g * RW + r;
      auto Xh = [&](int q) -> cf {
        if (q >= mlim) return cf{0.0f, 0.0f};
        const float ms = msc[q];
        cf v{istage[(2 * q) * OST2 + rr] * ms, istage[(2 * q + 1) * OST2 + rr] * ms};
        if (q == 0 || q == H) v.y = 0.0f;
        return v;
      };
      cf* wr = work + r * P1 * WP;
      constexpr int NBUILD = (NZ > 0) ? 2 * NZ * P2 : H;   // bins that can be non-zero: [0, NZ*P2) and [H - NZ*P2, H)
      for (int kb = lane; kb < NBUILD; kb += 32) {
        const int k = (NZ > 0 && kb >= NZ * P2) ? kb + (P1 - 2 * NZ) * P2 : kb;
        const int kk = H - k;
        cf o{0.0f, 0.0f};
        if (k < mlim || kk < mlim) {
          cf w;
          if (k <= mlim) w = tw2[k];
          else { w = tw2[kk]; w.x = -w.x; }
          o = c2r_merge(Xh(k), Xh(kk), w);
        }
        const int n1 = k / P2;
        wr[n1 * WP + (k - n1 * P2)] = o;
      }
    }
    __syncwarp();
    for (int t = lane; t < RW * P2; t += 32) {
      const int r = t / P2, n2 = t - r * P2;
      if (r < nv) fft2d_step1<P1, P2, +1, (NZ > 0 ? NZ : 0), (NZ > 0 ? P1 - NZ : 0)>(work + r * P1 * WP, WP, work + r * P1 * WP, tw, n2);
    }
    __syncwarp();
    for (int t = lane; t < RW * P1; t += 32) {
      const int r = t / P1, k1 = t - r * P1;
      if (r < nv) {
        cf v[P2];
        fft2d_step2<P1, P2, +1>(work + r * P1 * WP, k1, v);
        cf* orow = outb + r * H;
        static_for<0, P2>([&](auto cc) {
          constexpr int k2 = decltype(cc)::value;
          orow[k1 + P1 * k2] = v[k2];
        });
      }
    }
    __syncwarp();
    // epilogue: the row buffer is the output row in natural float order
    for (int r = 0; r < nv; ++r) {
      const size_t goff = ((size_t)bc * nlat + (k0 + g * RW + r)) * NLON;
      const float4* row4 = reinterpret_cast<const float4*>(outb + r * H);
      const float4* skip4 = skip ? reinterpret_cast<const float4*>(skip + goff) : nullptr;
      float4* y4 = reinterpret_cast<float4*>(y + goff);
      for (int i = lane; i < (NLON >> 2); i += 32) {
        float4 v = row4[i];
        v.x *= osc; v.y *= osc; v.z *= osc; v.w *= osc;
        if (skip4) { const float4 s = skip4[i]; v.x += s.x; v.y += s.y; v.z += s.z; v.w += s.w; }
        if (act_gelu & 1) { v.x = gelu_erf(v.x); v.y = gelu_erf(v.y); v.z = gelu_erf(v.z); v.w = gelu_erf(v.w); }
        if (act_gelu & 2) { v.x = rna_tf32_dev(v.x); v.y = rna_tf32_dev(v.y); v.z = rna_tf32_dev(v.z); v.w = rna_tf32_dev(v.w); }
        lsum += (v.x + v.y) + (v.z + v.w);
        lsq += (v.x * v.x + v.y * v.y) + (v.z * v.z + v.w * v.w);
        y4[i] = v;
      }
    }
    __syncwarp();
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
static bool pick_warps2(size_t fixed, size_t per_warp, int max_groups, int* nwarps, size_t* total) {
  for (int nw = 8; nw >= 1; nw >>= 1) {
    if (nw > max_groups && nw > 1) continue;
    const size_t t = fixed + 512 + (size_t)nw * (((per_warp + 127) & ~(size_t)127) + 32);
    if (t <= 227 * 1024) { *nwarps = nw; *total = t; return true; }
  }
  return false;
}

template <int P1, int P2, int RW, int NZ2 = 0, int RT = 32, bool TWG = false, bool DB = true>
static int launch_fwd(const msfno_plan* p, const float* x, float* Xt, const float* mscale, int zero_imag,
                      const float* in_scale, const float* in_shift, int B, int C, cudaStream_t st, const PeerMapDev& pm) {
  constexpr int H = P1 * P2, WP = WorkPitch<P2>::value;
  const int XS = xs_size(H, NZ2 > 0 ? NZ2 * P1 : p->mlim);
  const size_t fixed = sizeof(cf) * ((TWG ? 0 : H) + p->mlim + 1) + sizeof(float) * 2 * p->mlim * (RT + 1);
  const size_t per_warp = sizeof(cf) * ((size_t)(DB ? 2 : 1) * RW * H + (size_t)RW * P1 * WP + (DB ? (size_t)RW * XS : 0));
  int nw; size_t smem;
  if (!pick_warps2(fixed, per_warp, RT / RW, &nw, &smem))
    return record_error(MSFNO_ERR_UNSUPPORTED, "four-step FFT does not fit in shared memory");
  auto kern = rfft2d_kernel<P1, P2, RW, NZ2, RT, TWG, DB>;
  MSFNO_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid((p->nlat + RT - 1) / RT, B * C);
  kern<<<grid, nw * 32, smem, st>>>(x, Xt, reinterpret_cast<const cf*>(p->d_tw), reinterpret_cast<const cf*>(p->d_tw2),
                                    mscale, in_scale, in_shift, p->nlat, p->mlim, p->kpad, C, zero_imag,
                                    (p->precision == MSFNO_PREC_TF32 && !zero_imag) ? 1 : 0, pm);
  count_launch();
  MSFNO_CUDA_OK(cudaGetLastError());
  return MSFNO_OK;
}

template <int P1, int P2, int RW, int NZ = 0, int RT = 32, bool TWG = false>
static int launch_inv(const msfno_plan* p, const float* Yt, float* y, const float* mscale, const float* skip,
                      const float* out_scale, int act_gelu, double* stats, int B, int C, cudaStream_t st, const PeerMapDev& pm) {
  constexpr int H = P1 * P2, WP = WorkPitch<P2>::value;
  const size_t fixed = sizeof(cf) * ((TWG ? 0 : H) + p->mlim + 1) + sizeof(float) * 2 * p->mlim * (RT + 1) + 16 * 8 + sizeof(float) * p->mlim + 128;
  const size_t per_warp = sizeof(cf) * ((size_t)RW * P1 * WP + (size_t)RW * H);
  int nw; size_t smem;
  if (!pick_warps2(fixed, per_warp, RT / RW, &nw, &smem))
    return record_error(MSFNO_ERR_UNSUPPORTED, "four-step FFT does not fit in shared memory");
  // prefer two resident CTAs per SM (one CTA's staging fill overlaps another's FFT phase) when 4 warps allow it
  if (nw == 8) {
    const size_t t4 = fixed + 512 + 4 * (((per_warp + 127) & ~(size_t)127) + 32);
    if (t4 <= 112 * 1024 && RW == 1) { nw = 4; smem = t4; }   // (measured: helps the 1440-point rows, hurts RW = 4)
  }
  auto kern = irfft2d_kernel<P1, P2, RW, NZ, RT, TWG>;
  MSFNO_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid((p->nlat + RT - 1) / RT, B * C);
  kern<<<grid, nw * 32, smem, st>>>(Yt, y, reinterpret_cast<const cf*>(p->d_tw), reinterpret_cast<const cf*>(p->d_tw2),
                                    mscale, skip, out_scale, stats, p->nlat, p->mlim, p->kpad, C, act_gelu, pm);
  count_launch();
  MSFNO_CUDA_OK(cudaGetLastError());
  return MSFNO_OK;
}

bool fft2d_supported(int nlon) { return nlon == 1440 || nlon == 240 || nlon == 2880; }

int launch_rfft2d(const msfno_plan* p, const float* x, float* Xt, const float* mscale, int zero_imag,
                  const float* in_scale, const float* in_shift, int B, int C, cudaStream_t st, const PeerMapDev* pmp) {
  PeerMapDev pm{};
  if (pmp) pm = *pmp;
  switch (p->nlon) {
    case 1440:
      if (p->mlim <= 5 * 24) return launch_fwd<24, 30, 1, 5>(p, x, Xt, mscale, zero_imag, in_scale, in_shift, B, C, st, pm);
      return launch_fwd<24, 30, 1>(p, x, Xt, mscale, zero_imag, in_scale, in_shift, B, C, st, pm);
    case 240: return launch_fwd<15, 8, 4>(p, x, Xt, mscale, zero_imag, in_scale, in_shift, B, C, st, pm);
    case 2880:   // 0.125 degree grid: eight warps per SM (see RT above); second-step outputs pruned to the kept orders
      // H = 1440 = 48 x 30: step 1 is ONE pass of 30 lane tasks (36 x 40 needs two passes of 40 and of 36 tasks, the
      // second with 8 / 4 active lanes)
      if (p->mlim <= 5 * 48) return launch_fwd<48, 30, 1, 5, 16, true, false>(p, x, Xt, mscale, zero_imag, in_scale, in_shift, B, C, st, pm);
      return launch_fwd<48, 30, 1, 0, 16, true, false>(p, x, Xt, mscale, zero_imag, in_scale, in_shift, B, C, st, pm);
    default: return record_error(MSFNO_ERR_UNSUPPORTED, "four-step FFT: unsupported nlon");
  }
}

int launch_irfft2d(const msfno_plan* p, const float* Yt, float* y, const float* mscale, const float* skip,
                   const float* out_scale, int act_gelu, double* stats, int B, int C, cudaStream_t st, const PeerMapDev* pmp) {
  PeerMapDev pm{};
  if (pmp) pm = *pmp;
  switch (p->nlon) {
    case 1440:
      if (p->mlim <= 4 * 30) return launch_inv<24, 30, 1, 4>(p, Yt, y, mscale, skip, out_scale, act_gelu, stats, B, C, st, pm);
      return launch_inv<24, 30, 1>(p, Yt, y, mscale, skip, out_scale, act_gelu, stats, B, C, st, pm);
    case 240: return launch_inv<15, 8, 4>(p, Yt, y, mscale, skip, out_scale, act_gelu, stats, B, C, st, pm);
    case 2880:
      if (p->mlim <= 8 * 30) return launch_inv<48, 30, 1, 8, 16, true>(p, Yt, y, mscale, skip, out_scale, act_gelu, stats, B, C, st, pm);
      return launch_inv<48, 30, 1, 0, 16, true>(p, Yt, y, mscale, skip, out_scale, act_gelu, stats, B, C, st, pm);
    default: return record_error(MSFNO_ERR_UNSUPPORTED, "four-step FFT: unsupported nlon");
  }
}

}  // namespace msfno
