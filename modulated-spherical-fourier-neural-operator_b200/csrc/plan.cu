// Plan construction, Legendre-table re-layout, packed-position maps, grouped-GEMM descriptors and
// the coefficient layout-change kernel.
//
// replaces: the device-side state of torch_harmonics RealSHT/InverseRealSHT (buffers `weights`,
// `pct`; /root/reference MSFNO/Models/sfno/sfnonet.py:537-555) and the index plumbing of
// SpectralConvS2 (torch.tril_indices gather/scatter, layers.py:366-370,408-413).
#include <math.h>
#include <string.h>

#include <atomic>
#include <string>

#include "common.cuh"
#include <cstdlib>

#include "plan.h"

namespace msfno {

static thread_local std::string g_last_error;
static std::atomic<unsigned long long> g_launches{0};
bool pdl_enabled() {
  static const bool on = !dbg_env("MSFNO_NO_PDL");
  return on;
}

void count_launch(int n) { g_launches.fetch_add((unsigned long long)n, std::memory_order_relaxed); }

// engine of the fp32 (1e-5) tier's GEMM-shaped stages: 3xTF32 on the tensor cores (default) or CUDA-core FFMA
static std::atomic<int> g_fp32_engine{MSFNO_FP32_ENGINE_TC3X};
bool fp32_engine_x3() { return g_fp32_engine.load(std::memory_order_relaxed) == MSFNO_FP32_ENGINE_TC3X; }

int record_error(int code, const char* msg) {
  g_last_error = msg;
  return code;
}
int record_cuda_error(cudaError_t e, const char* file, int line) {
  g_last_error = std::string("CUDA error: ") + cudaGetErrorString(e) + " at " + file + ":" + std::to_string(line);
  return MSFNO_ERR_CUDA;
}

// ---- table re-layout --------------------------------------------------------------------
// T: [mmax][lmax][nlat] (reference layout).  tab_lk[m][j][k] = T[m][m+j][k], zero padded.
__device__ __forceinline__ float rna_tf32_f(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;\n" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}
// round_tf32: store round-to-nearest TF32 values (tensor-core tier: the MMA truncates its fp32 operands, which would bias
// every Legendre sum towards zero; rounding the tables and the other operand at their producers removes the bias)
__global__ void relayout_lk_kernel(const float* __restrict__ T, float* __restrict__ out, int lmax, int nlat, int mlim,
                                   int Lj, int kpad, int round_tf32) {
  const long long total = (long long)mlim * Lj * kpad;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int k = (int)(i % kpad);
    const long long r = i / kpad;
    const int j = (int)(r % Lj), m = (int)(r / Lj);
    const int l = m + j;
    const float v = (l < lmax && k < nlat) ? T[((long long)m * lmax + l) * nlat + k] : 0.0f;
    out[i] = round_tf32 ? rna_tf32_f(v) : v;
  }
}
// tab_kl[m][k][j] = T[m][m+j][k], zero padded in j.
__global__ void relayout_kl_kernel(const float* __restrict__ T, float* __restrict__ out, int lmax, int nlat, int mlim,
                                   int Lj, int round_tf32) {
  const long long total = (long long)mlim * nlat * Lj;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int j = (int)(i % Lj);
    const long long r = i / Lj;
    const int k = (int)(r % nlat), m = (int)(r / nlat);
    const int l = m + j;
    const float v = (l < lmax) ? T[((long long)m * lmax + l) * nlat + k] : 0.0f;
    out[i] = round_tf32 ? rna_tf32_f(v) : v;
  }
}
// flag != 0 if any entry with l < m (or m >= mlim) is non-zero
__global__ void check_triangular_kernel(const float* __restrict__ T, int* flag, int mmax, int lmax, int nlat) {
  const long long total = (long long)mmax * lmax * nlat;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / nlat;
    const int l = (int)(r % lmax), m = (int)(r / lmax);
    if (l < m && T[i] != 0.0f) atomicOr(flag, 1);
  }
}

// ---- coefficient layout change ------------------------------------------------------------
__global__ void coef_relayout_kernel(const float* __restrict__ src, int sl, float* __restrict__ dst, int dl,
                                     const int* __restrict__ poff, const int* __restrict__ p2lm, int B, int C, int lmax,
                                     int mmax, int mlim, int P) {
  const int C2 = 2 * C;
  long long total;
  if (dl == MSFNO_LAYOUT_STD) total = (long long)B * C * lmax * mmax * 2;
  else total = (long long)B * P * C2;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    int b, ch, l, m, p;
    if (dl == MSFNO_LAYOUT_STD) {
      long long r = i;
      const int ri = (int)(r & 1); r >>= 1;
      m = (int)(r % mmax); r /= mmax;
      l = (int)(r % lmax); r /= lmax;
      const int c = (int)(r % C);
      b = (int)(r / C);
      ch = 2 * c + ri;
      p = (l >= m && m < mlim) ? poff[m] + (l - m) : -1;
    } else {
      if (dl == MSFNO_LAYOUT_PM) {
        ch = (int)(i % C2);
        const long long r = i / C2;
        p = (int)(r % P);
        b = (int)(r / P);
      } else {
        p = (int)(i % P);
        const long long r = i / P;
        ch = (int)(r % C2);
        b = (int)(r / C2);
      }
      const int lm = p2lm[p];
      if (lm < 0) { l = 0; m = 0; p = -1; }
      else { l = lm / mmax; m = lm - l * mmax; }
    }
    float v = 0.0f;
    if (p >= 0) {
      if (sl == MSFNO_LAYOUT_STD) v = src[((((long long)b * C + (ch >> 1)) * lmax + l) * mmax + m) * 2 + (ch & 1)];
      else if (sl == MSFNO_LAYOUT_PM) v = src[((long long)b * P + p) * C2 + ch];
      else v = src[((long long)b * C2 + ch) * P + p];
    }
    dst[i] = v;
  }
}

template <typename T>
static int upload(T** dptr, const std::vector<T>& h) {
  MSFNO_CUDA_OK(cudaMalloc(dptr, sizeof(T) * (h.size() ? h.size() : 1)));
  if (!h.empty()) MSFNO_CUDA_OK(cudaMemcpy(*dptr, h.data(), sizeof(T) * h.size(), cudaMemcpyHostToDevice));
  return MSFNO_OK;
}

int plan_groups(msfno_plan* p, int kind, int B, int C, const GemmGroup** out, int* ngroups, int m_lo, int m_hi) {
  // orders [m_lo, m_hi) only (spatially sharded SHT: a rank owns a contiguous range of azimuthal orders); the
  // lat<->m intermediate and the coefficient buffers are then indexed relative to m_lo / poff[m_lo].
  if (m_hi < 0 || m_hi > p->mlim) m_hi = p->mlim;
  if (m_lo < 0) m_lo = 0;
  const int mloc = m_hi - m_lo;
  std::lock_guard<std::mutex> lk(p->mu);
  std::vector<int> key = {kind, B, C, m_lo, m_hi};
  auto it = p->groups.find(key);
  if (it == p->groups.end()) {
    const long long C2 = 2 * C, kpad = p->kpad, Lj = p->Lj, nlat = p->nlat;
    const long long p0 = (m_lo < p->mlim) ? p->h_poff[m_lo] : p->P;
    const long long p1 = (m_hi < p->mlim) ? p->h_poff[m_hi] : p->P;
    const long long Ploc = p1 - p0;
    std::vector<GemmGroup> h((size_t)B * (mloc > 0 ? mloc : 0));
    for (int b = 0; b < B; ++b)
      for (int m = m_lo; m < m_hi; ++m) {
        GemmGroup g{};
        const long long xt = ((long long)b * mloc + (m - m_lo)) * C2 * kpad;       // lat<->m intermediate
        const long long pm = ((long long)b * Ploc + p->h_poff[m] - p0) * C2;       // PM coefficient rows of order m
        const long long cm = (long long)b * C2 * Ploc + p->h_poff[m] - p0;         // CM coefficient columns of order m
        const int len = p->h_plen4[m];
        switch (kind) {
          case GK_ANALYSIS:      g = GemmGroup{(long long)m * Lj * kpad, xt, pm, len, (int)C2, (int)nlat, 0}; break;
          case GK_ANALYSIS_ADJ:  g = GemmGroup{pm, (long long)m * Lj * kpad, xt, (int)C2, (int)nlat, len, 0}; break;
          case GK_SYNTHESIS:     g = GemmGroup{cm, (long long)m * nlat * Lj, xt, (int)C2, (int)nlat, len, 0}; break;
          default:               g = GemmGroup{xt, (long long)m * nlat * Lj, cm, (int)C2, len, (int)nlat, 0}; break;
        }
        h[(size_t)b * mloc + (m - m_lo)] = g;
      }
    GemmGroup* d = nullptr;
    int rc = upload(&d, h);
    if (rc) return rc;
    it = p->groups.emplace(key, d).first;
  }
  *out = it->second;
  *ngroups = B * (mloc > 0 ? mloc : 0);
  return MSFNO_OK;
}

}  // namespace msfno

using namespace msfno;

extern "C" {

const char* msfno_last_error(void) { return g_last_error.c_str(); }

unsigned long long msfno_launch_count(void) { return g_launches.load(); }

int msfno_set_fp32_engine(int engine) {
  if (engine != MSFNO_FP32_ENGINE_TC3X && engine != MSFNO_FP32_ENGINE_FFMA)
    return record_error(MSFNO_ERR_BAD_SHAPE, "set_fp32_engine: bad argument");
  g_fp32_engine.store(engine, std::memory_order_relaxed);
  return MSFNO_OK;
}
int msfno_get_fp32_engine(void) { return g_fp32_engine.load(std::memory_order_relaxed); }

const char* msfno_build_info(void) {
  return "{\"arch\": \"sm_100a\", \"abi\": 4, \"tiers\": [\"fp32 (3xTF32 on tcgen05, or FFMA)\", \"tf32\"], \"fft\": \"four-step in-register (fp32 tier; nlon 240 / 1440 / 2880)\", "
         "\"tensor_core\": [\"tcgen05 tf32 gemm (cta_group::1 and ::2)\", \"3xTF32 gemm (two issuing threads, A operand in TMEM, small terms in a persistent accumulator)\", \"dft gemm\", "
         "\"parity-split persistent inverse dft (tma stores)\", \"fused mlp (A operand in TMEM)\", \"1x1 conv\"], "
         "\"streams\": [\"specconv tma ring\"], \"multi_gpu\": [\"cuda ipc peer buffers\", \"nvlink block copy\", \"exchange fused into the fft kernels\", \"flag barrier\"]}";
}

int msfno_plan_create(msfno_plan** out, int nlat, int nlon, int lmax, int mmax) {
  if (!out || nlat < 1 || nlon < 2 || lmax < 1 || mmax < 1) return record_error(MSFNO_ERR_BAD_SHAPE, "plan_create: bad sizes");
  if (nlon % 2 != 0) return record_error(MSFNO_ERR_UNSUPPORTED, "plan_create: nlon must be even");
  if (mmax > nlon / 2 + 1) return record_error(MSFNO_ERR_BAD_SHAPE, "plan_create: mmax > nlon/2+1");
  msfno_plan* p = new msfno_plan();
  p->nlat = nlat; p->nlon = nlon; p->lmax = lmax; p->mmax = mmax;
  if (!make_schedule(nlon / 2, &p->sched)) {
    delete p;
    return record_error(MSFNO_ERR_UNSUPPORTED, "plan_create: nlon/2 must factor into 2, 3 and 5");
  }
  cudaGetDevice(&p->device);
  p->mlim = mmax < lmax ? mmax : lmax;
  p->kpad = (nlat + 31) / 32 * 32;
  p->Lj = (lmax + 3) / 4 * 4;
  p->h_poff.assign(mmax, 0);
  p->h_plen4.assign(mmax, 0);
  int P = 0;
  for (int m = 0; m < mmax; ++m) {
    p->h_poff[m] = P;
    if (m < p->mlim) {
      p->h_plen4[m] = (lmax - m + 3) / 4 * 4;
      P += p->h_plen4[m];
    }
  }
  p->P = P;
  std::vector<int32_t> p2lm(P, -1);
  for (int m = 0; m < p->mlim; ++m)
    for (int l = m; l < lmax; ++l) p2lm[p->h_poff[m] + (l - m)] = l * mmax + m;
  for (int l = 0; l < lmax; ++l)
    for (int m = 0; m <= l && m < mmax; ++m) p->h_n2p.push_back(p->h_poff[m] + (l - m));
  p->ntril = (int)p->h_n2p.size();

  const int H = nlon / 2;
  std::vector<float> tw(2 * H), tw2(2 * (p->mlim + 1));
  for (int t = 0; t < H; ++t) {
    const double a = -2.0 * M_PI * t / H;
    tw[2 * t] = (float)cos(a); tw[2 * t + 1] = (float)sin(a);
  }
  for (int m = 0; m <= p->mlim; ++m) {
    const double a = -2.0 * M_PI * m / nlon;
    tw2[2 * m] = (float)cos(a); tw2[2 * m + 1] = (float)sin(a);
  }
  std::vector<float> s_rfft(p->mlim), s_irfft_adj(p->mlim), s_irfft(p->mlim), s_rfft_adj(p->mlim);
  for (int m = 0; m < p->mlim; ++m) {
    const double cm = (m == 0 || m == H) ? 1.0 : 2.0;
    s_rfft[m] = (float)(2.0 * M_PI / nlon);
    s_irfft_adj[m] = (float)cm;
    s_irfft[m] = 1.0f;
    s_rfft_adj[m] = (float)(2.0 * M_PI / nlon / cm);
  }
  int rc = 0;
  rc = rc ? rc : upload(&p->d_tw, tw);
  rc = rc ? rc : upload(&p->d_tw2, tw2);
  rc = rc ? rc : upload(&p->d_scale_rfft, s_rfft);
  rc = rc ? rc : upload(&p->d_scale_irfft_adj, s_irfft_adj);
  rc = rc ? rc : upload(&p->d_scale_irfft, s_irfft);
  rc = rc ? rc : upload(&p->d_scale_rfft_adj, s_rfft_adj);
  rc = rc ? rc : upload(&p->d_poff, p->h_poff);
  rc = rc ? rc : upload(&p->d_n2p, p->h_n2p);
  rc = rc ? rc : upload(&p->d_p2lm, p2lm);
  if (!rc && cudaMalloc(&p->d_flag, sizeof(int32_t)) != cudaSuccess) rc = record_error(MSFNO_ERR_CUDA, "cudaMalloc flag");
  if (rc) { msfno_plan_destroy(p); return rc; }
  *out = p;
  return MSFNO_OK;
}

int msfno_plan_destroy(msfno_plan* p) {
  if (!p) return MSFNO_OK;
  cudaFree(p->d_tw); cudaFree(p->d_tw2);
  cudaFree(p->d_scale_rfft); cudaFree(p->d_scale_irfft_adj); cudaFree(p->d_scale_irfft); cudaFree(p->d_scale_rfft_adj);
  cudaFree(p->d_dft_fwd); cudaFree(p->d_dft_inv); cudaFree(p->d_dft_inv_eo);
  cudaFree(p->d_poff); cudaFree(p->d_n2p); cudaFree(p->d_p2lm);
  cudaFree(p->d_tab_lk); cudaFree(p->d_tab_kl); cudaFree(p->d_flag);
  for (auto& kv : p->groups) cudaFree(kv.second);
  delete p;
  return MSFNO_OK;
}

long msfno_plan_query(const msfno_plan* p, int what) {
  if (!p) return -1;
  switch (what) {
    case MSFNO_Q_KPAD: return p->kpad;
    case MSFNO_Q_MLIM: return p->mlim;
    case MSFNO_Q_NPACK: return p->P;
    case MSFNO_Q_NTRIL: return p->ntril;
    case MSFNO_Q_LJ: return p->Lj;
    default: return -1;
  }
}

int msfno_plan_get_maps(const msfno_plan* p, int32_t* poff_host, int32_t* n2p_host) {
  if (!p) return record_error(MSFNO_ERR_BAD_STATE, "null plan");
  if (poff_host) memcpy(poff_host, p->h_poff.data(), sizeof(int32_t) * p->h_poff.size());
  if (n2p_host) memcpy(n2p_host, p->h_n2p.data(), sizeof(int32_t) * p->h_n2p.size());
  return MSFNO_OK;
}

int msfno_plan_set_table(msfno_plan* p, const float* table, int analysis, void* stream) {
  if (!p || !table) return record_error(MSFNO_ERR_BAD_STATE, "plan_set_table: null argument");
  cudaStream_t st = (cudaStream_t)stream;
  MSFNO_CUDA_OK(cudaMemsetAsync(p->d_flag, 0, sizeof(int32_t), st));
  check_triangular_kernel<<<592, 256, 0, st>>>(table, p->d_flag, p->mmax, p->lmax, p->nlat);
  if (analysis) {
    if (!p->d_tab_lk) MSFNO_CUDA_OK(cudaMalloc(&p->d_tab_lk, sizeof(float) * (size_t)p->mlim * p->Lj * p->kpad));
    relayout_lk_kernel<<<592, 256, 0, st>>>(table, p->d_tab_lk, p->lmax, p->nlat, p->mlim, p->Lj, p->kpad,
                                            p->precision == MSFNO_PREC_TF32);
  } else {
    if (!p->d_tab_kl) MSFNO_CUDA_OK(cudaMalloc(&p->d_tab_kl, sizeof(float) * (size_t)p->mlim * p->nlat * p->Lj));
    relayout_kl_kernel<<<592, 256, 0, st>>>(table, p->d_tab_kl, p->lmax, p->nlat, p->mlim, p->Lj,
                                            p->precision == MSFNO_PREC_TF32);
  }
  count_launch(2);
  MSFNO_CUDA_OK(cudaGetLastError());
  int32_t flag = 0;
  MSFNO_CUDA_OK(cudaMemcpyAsync(&flag, p->d_flag, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
  MSFNO_CUDA_OK(cudaStreamSynchronize(st));
  if (flag) return record_error(MSFNO_ERR_UNSUPPORTED, "plan_set_table: table has non-zero entries with l < m");
  return MSFNO_OK;
}

int msfno_coef_relayout(const msfno_plan* p, const float* src, int sl, float* dst, int dl, int B, int C, void* stream) {
  if (!p || !src || !dst) return record_error(MSFNO_ERR_BAD_STATE, "coef_relayout: null argument");
  if (sl == dl || sl < 0 || sl > 2 || dl < 0 || dl > 2) return record_error(MSFNO_ERR_BAD_SHAPE, "coef_relayout: bad layouts");
  coef_relayout_kernel<<<148 * 8, 256, 0, (cudaStream_t)stream>>>(src, sl, dst, dl, p->d_poff, p->d_p2lm, B, C, p->lmax,
                                                                 p->mmax, p->mlim, p->P);
  count_launch();
  MSFNO_CUDA_OK(cudaGetLastError());
  return MSFNO_OK;
}

}  // extern "C"
