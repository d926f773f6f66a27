// Forward / inverse spherical harmonic transforms and their adjoints: longitude FFT kernel + per-order
// Legendre contraction (grouped GEMM over the azimuthal order m).
//
// replaces: torch_harmonics.RealSHT.forward / InverseRealSHT.forward, called
// from /root/reference MSFNO/Models/sfno/layers.py:405,421 (SpectralConvS2) and :629,638
// (SpectralAttentionS2), and the autograd adjoints of both (Appendix A.4).
#include "common.cuh"
#include "plan.h"

using namespace msfno;

// One grouped Legendre contraction (forward or adjoint) over the azimuthal orders [m_lo, m_hi).
// Engine: tensor cores whenever TMA can address the operands -- plain TF32 for the forward transforms of the tf32 tier,
// 3xTF32 (fp32 grade) for the fp32 tier and for every adjoint (gradients are not TF32-rounded by their producers, and
// the truncation bias of a plain TF32 MMA would compound over the 24 transforms of a backward pass); the CUDA-core
// FFMA kernel otherwise (msfno_set_fp32_engine(FFMA), unaligned operands).
static int legendre_gemm(msfno_plan* p, int kind, const float* A, long long lda, int a_k, const float* B, long long ldb,
                         int b_k, float* D, long long ldd, int maxM, int maxN, int Bsz, int C, cudaStream_t st,
                         int m_lo = 0, int m_hi = -1, int round_out = 0) {
  const GemmGroup* groups = nullptr;
  int ng = 0;
  int rc = plan_groups(p, kind, Bsz, C, &groups, &ng, m_lo, m_hi);
  if (rc) return rc;
  GemmLaunch g{};
  g.A = A; g.B = B; g.D = D;
  g.lda = lda; g.ldb = ldb; g.ldd = ldd;
  g.a_kmajor = a_k; g.b_kmajor = b_k;
  g.groups = groups; g.ngroups = ng; g.maxM = maxM; g.maxN = maxN;
  const bool fwd = kind == GK_ANALYSIS || kind == GK_SYNTHESIS;
  const bool tf32 = p->precision == MSFNO_PREC_TF32 && fwd;
  if ((tf32 || fp32_engine_x3()) && gemm_tc_supported(g)) {
    g.x3 = tf32 ? 0 : 1;
    // the 2-D buffers behind A and B as TMA sees them (zero-fill outside)
    if (m_hi < 0 || m_hi > p->mlim) m_hi = p->mlim;
    const long long mloc = m_hi - m_lo;
    const long long Ploc = ((m_hi < p->mlim) ? p->h_poff[m_hi] : p->P) - p->h_poff[m_lo];
    const long long C2 = 2 * C;
    switch (kind) {
      case GK_ANALYSIS:       // tab_lk [mlim Lj][kpad] x Xt [B mloc 2C][kpad]
        return launch_gemm_tc(g, (long long)p->mlim * p->Lj, p->kpad, Bsz * mloc * C2, p->kpad, 0, st);
      case GK_SYNTHESIS:      // coef_cm [B 2C][Ploc] x tab_kl [mlim nlat][Lj]
        return launch_gemm_tc(g, Bsz * C2, Ploc, (long long)p->mlim * p->nlat, p->Lj, tf32 ? round_out : 0, st);
      case GK_ANALYSIS_ADJ:   // g_pm [B Ploc][2C] (M contiguous) x tab_lk [mlim][Lj][kpad] (N contiguous, stacked tables)
        g.b_group_rows = p->Lj;
        return launch_gemm_tc(g, Bsz * Ploc, C2, (long long)p->mlim * p->Lj, p->kpad, 0, st);
      default:                // gYt [B mloc 2C][kpad] x tab_kl [mlim][nlat][Lj] (N contiguous, stacked tables)
        g.b_group_rows = p->nlat;
        return launch_gemm_tc(g, Bsz * mloc * C2, p->kpad, (long long)p->mlim * p->nlat, p->Lj, 0, st);
    }
  }
  return launch_gemm_ffma(g, st);
}

// ---- lat <-> m exchange buffers of the spatially sharded transform (one launch per direction) -----------------------
// `flat` is what NCCL all_to_all_single moves: for every peer s a block [rows][pad32(n_s)] (rows = local orders x 2C, n_s =
// the peer's latitude count, zero-padded pitch = the peer's own lat-contiguous pitch); `full` is the Legendre stage's
// operand [rows][pad_full] with all nlat latitudes.  gather: blocks -> full (columns [nlat, pad_full) zeroed);
// scatter: full -> blocks (pitch tails zeroed).
struct LatSegs {
  int n;
  int lo[MSFNO_MAX_LAT_SEGMENTS];        // first latitude of peer s
  int cnt[MSFNO_MAX_LAT_SEGMENTS];       // its latitude count
  long long base[MSFNO_MAX_LAT_SEGMENTS];  // float offset of its block in `flat`
};

__global__ void lat_segments_kernel(int gather, float* __restrict__ flat, float* __restrict__ full, long long rows, int pad_full,
                                    int nlat, LatSegs sg) {
  const int s = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  for (long long r = (long long)blockIdx.x * nw + warp; r < rows; r += (long long)gridDim.x * nw) {
    float* frow = full + r * pad_full;
    if (s == sg.n) {   // pseudo-segment: the zero tail of `full`
      if (gather)
        for (int j = nlat + lane; j < pad_full; j += 32) frow[j] = 0.0f;
      continue;
    }
    const int n = sg.cnt[s], pitch = (n + 31) & ~31;
    float* brow = flat + sg.base[s] + r * pitch;
    if (gather) {
      for (int j = lane; j < n; j += 32) frow[sg.lo[s] + j] = brow[j];
    } else {
      for (int j = lane; j < pitch; j += 32) brow[j] = j < n ? frow[sg.lo[s] + j] : 0.0f;
    }
  }
}

extern "C" {

size_t msfno_sht_ws_floats(const msfno_plan* p, int B, int C) {
  if (!p) return 0;
  return (size_t)B * p->mlim * 2 * C * p->kpad;
}

int msfno_sht_fwd(msfno_plan* p, const float* x, const float* in_scale, const float* in_shift, float* coef_pm, float* ws,
                  int B, int C, void* stream) {
  if (!p || !x || !coef_pm || !ws || B < 1 || C < 1) return record_error(MSFNO_ERR_BAD_SHAPE, "sht_fwd: bad argument");
  if (!p->d_tab_lk) return record_error(MSFNO_ERR_BAD_STATE, "sht_fwd: analysis table (RealSHT.weights) not set");
  cudaStream_t st = (cudaStream_t)stream;
  // longitude transform: DFT GEMM on the tensor cores in the TF32 tier (dft_tc.cu), FFT kernels otherwise
  const bool dft_f = p->precision == MSFNO_PREC_TF32 && dft_tc_supported(p) && ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(ws)) & 15) == 0;
  int rc = dft_f ? launch_dft_fwd(p, x, ws, in_scale, in_shift, B, C, st)
                                                                      : launch_rfft_trunc(p, x, ws, p->d_scale_rfft, 0, in_scale, in_shift, B, C, st);
  if (rc) return rc;
  // coef_pm[b][poff[m]+j][ch] = sum_k tab_lk[m][j][k] * Xt[b][m][ch][k]
  return legendre_gemm(p, GK_ANALYSIS, p->d_tab_lk, p->kpad, 1, ws, p->kpad, 1, coef_pm, 2 * C, p->h_plen4[0], 2 * C, B, C, st);
}

int msfno_sht_bwd(msfno_plan* p, const float* g_pm, const float* in_scale, float* gx, float* ws, int B, int C, void* stream) {
  if (!p || !g_pm || !gx || !ws || B < 1 || C < 1) return record_error(MSFNO_ERR_BAD_SHAPE, "sht_bwd: bad argument");
  if (!p->d_tab_lk) return record_error(MSFNO_ERR_BAD_STATE, "sht_bwd: analysis table (RealSHT.weights) not set");
  cudaStream_t st = (cudaStream_t)stream;
  // gXt[b][m][ch][k] = sum_j g_pm[b][poff[m]+j][ch] * tab_lk[m][j][k]
  int rc = legendre_gemm(p, GK_ANALYSIS_ADJ, g_pm, 2 * C, 0, p->d_tab_lk, p->kpad, 0, ws, p->kpad, 2 * C, p->nlat, B, C, st);
  if (rc) return rc;
  return launch_irfft_trunc(p, ws, gx, p->d_scale_rfft_adj, nullptr, in_scale, 0, nullptr, B, C, st);
}

int msfno_isht_fwd(msfno_plan* p, const float* coef_cm, float* y, float* ws, int B, int C, const float* skip_add,
                   int act_gelu, double* stats, void* stream) {
  if (!p || !coef_cm || !y || !ws || B < 1 || C < 1) return record_error(MSFNO_ERR_BAD_SHAPE, "isht_fwd: bad argument");
  if (!p->d_tab_kl) return record_error(MSFNO_ERR_BAD_STATE, "isht_fwd: synthesis table (InverseRealSHT.pct) not set");
  cudaStream_t st = (cudaStream_t)stream;
  // Yt[b][m][ch][k] = sum_j coef_cm[b][ch][poff[m]+j] * tab_kl[m][k][j]
  const bool dft = p->precision == MSFNO_PREC_TF32 && dft_tc_supported(p) &&
                   ((reinterpret_cast<uintptr_t>(y) | reinterpret_cast<uintptr_t>(ws) | reinterpret_cast<uintptr_t>(skip_add)) & 15) == 0;
  // (the DFT GEMM reads Yt as a tensor-core operand: round it to TF32 where it is produced)
  int rc = legendre_gemm(p, GK_SYNTHESIS, coef_cm, p->P, 1, p->d_tab_kl, p->Lj, 1, ws, p->kpad, 2 * C, p->nlat, B, C, st, 0, -1,
                         dft ? 1 : 0);
  if (rc) return rc;
  if (stats) MSFNO_CUDA_OK(cudaMemsetAsync(stats, 0, sizeof(double) * 2 * (size_t)B * C, st));
  if (dft) return launch_dft_inv(p, ws, y, skip_add, act_gelu, stats, B, C, st);
  return launch_irfft_trunc(p, ws, y, p->d_scale_irfft, skip_add, nullptr, act_gelu, stats, B, C, st);
}

int msfno_isht_bwd(msfno_plan* p, const float* gy, float* g_cm, float* ws, int B, int C, void* stream) {
  if (!p || !gy || !g_cm || !ws || B < 1 || C < 1) return record_error(MSFNO_ERR_BAD_SHAPE, "isht_bwd: bad argument");
  if (!p->d_tab_kl) return record_error(MSFNO_ERR_BAD_STATE, "isht_bwd: synthesis table (InverseRealSHT.pct) not set");
  cudaStream_t st = (cudaStream_t)stream;
  int rc = launch_rfft_trunc(p, gy, ws, p->d_scale_irfft_adj, 1, nullptr, nullptr, B, C, st);
  if (rc) return rc;
  // g_cm[b][ch][poff[m]+j] = sum_k gYt[b][m][ch][k] * tab_kl[m][k][j]
  return legendre_gemm(p, GK_SYNTHESIS_ADJ, ws, p->kpad, 1, p->d_tab_kl, p->Lj, 0, g_cm, p->P, 2 * C, p->h_plen4[0], B, C, st);
}

int msfno_plan_set_precision(msfno_plan* p, int precision) {
  if (!p || (precision != MSFNO_PREC_FP32 && precision != MSFNO_PREC_TF32)) return record_error(MSFNO_ERR_BAD_SHAPE, "plan_set_precision: bad argument");
  p->precision = precision;
  return MSFNO_OK;
}

// ---- stage-level entry points (spatially sharded SHT) -------------------------------------
int msfno_fft_stage(msfno_plan* p, int inverse, int adjoint, const float* src, float* dst, int B, int C, void* stream) {
  if (!p || !src || !dst || B < 1 || C < 1) return record_error(MSFNO_ERR_BAD_SHAPE, "fft_stage: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  if (!inverse)  // x -> Xt : RealSHT longitude stage, or (adjoint) the adjoint of irfft
    return launch_rfft_trunc(p, src, dst, adjoint ? p->d_scale_irfft_adj : p->d_scale_rfft, adjoint ? 1 : 0, nullptr, nullptr,
                             B, C, st);
  return launch_irfft_trunc(p, src, dst, adjoint ? p->d_scale_rfft_adj : p->d_scale_irfft, nullptr, nullptr, 0, nullptr, B,
                            C, st);
}

int msfno_fft_stage_peer(msfno_plan* p, int inverse, const float* x, float* y, const msfno_peer_map* map, int C, void* stream) {
  if (!p || !map || C < 1 || (inverse ? !y : !x) || map->world < 1 || map->world > MSFNO_MAX_PEERS || map->pitch < 1 || map->lat_lo < 0 ||
      map->lat_lo + p->nlat > map->pitch)
    return record_error(MSFNO_ERR_BAD_SHAPE, "fft_stage_peer: bad argument");
  PeerMapDev pm{};
  pm.world = map->world; pm.pitch = map->pitch; pm.lat_lo = map->lat_lo;
  for (int s = 0; s <= map->world; ++s) pm.mb[s] = map->m_bounds[s];
  if (pm.mb[0] != 0 || pm.mb[map->world] != p->mlim) return record_error(MSFNO_ERR_BAD_SHAPE, "fft_stage_peer: order bounds do not cover [0, mlim)");
  for (int s = 0; s < map->world; ++s) {
    if (pm.mb[s + 1] < pm.mb[s] || (pm.mb[s + 1] > pm.mb[s] && !map->buf[s])) return record_error(MSFNO_ERR_BAD_SHAPE, "fft_stage_peer: bad peer entry");
    pm.buf[s] = map->buf[s];
  }
  // the four-step kernels only (nlon 240 / 1440 / 2880): the generic radix kernels keep the local intermediate
  if (!fft2d_supported(p->nlon) || (reinterpret_cast<uintptr_t>(inverse ? (const void*)y : (const void*)x) & 15))
    return record_error(MSFNO_ERR_UNSUPPORTED, "fft_stage_peer: needs the four-step FFT kernels (nlon 240 / 1440 / 2880, 16-byte aligned grid)");
  cudaStream_t st = (cudaStream_t)stream;
  if (!inverse) return launch_rfft2d(p, x, nullptr, p->d_scale_rfft, 0, nullptr, nullptr, 1, C, st, &pm);
  return launch_irfft2d(p, nullptr, y, p->d_scale_irfft, nullptr, nullptr, 0, nullptr, 1, C, st, &pm);
}

int msfno_legendre_stage(msfno_plan* p, int kind, const float* src, float* dst, int m_lo, int m_hi, int B, int C,
                         void* stream) {
  if (!p || !src || !dst || B < 1 || C < 1 || m_lo < 0 || m_hi > p->mlim || m_lo >= m_hi)
    return record_error(MSFNO_ERR_BAD_SHAPE, "legendre_stage: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  const long long p0 = p->h_poff[m_lo];
  const long long p1 = (m_hi < p->mlim) ? p->h_poff[m_hi] : p->P;
  const long long Ploc = p1 - p0;
  const int maxlen = p->h_plen4[m_lo];
  switch (kind) {
    case GK_ANALYSIS:
      if (!p->d_tab_lk) return record_error(MSFNO_ERR_BAD_STATE, "legendre_stage: analysis table not set");
      return legendre_gemm(p, kind, p->d_tab_lk, p->kpad, 1, src, p->kpad, 1, dst, 2 * C, maxlen, 2 * C, B, C, st, m_lo, m_hi);
    case GK_ANALYSIS_ADJ:
      if (!p->d_tab_lk) return record_error(MSFNO_ERR_BAD_STATE, "legendre_stage: analysis table not set");
      return legendre_gemm(p, kind, src, 2 * C, 0, p->d_tab_lk, p->kpad, 0, dst, p->kpad, 2 * C, p->nlat, B, C, st, m_lo, m_hi);
    case GK_SYNTHESIS:
      if (!p->d_tab_kl) return record_error(MSFNO_ERR_BAD_STATE, "legendre_stage: synthesis table not set");
      return legendre_gemm(p, kind, src, Ploc, 1, p->d_tab_kl, p->Lj, 1, dst, p->kpad, 2 * C, p->nlat, B, C, st, m_lo, m_hi);
    case GK_SYNTHESIS_ADJ:
      if (!p->d_tab_kl) return record_error(MSFNO_ERR_BAD_STATE, "legendre_stage: synthesis table not set");
      return legendre_gemm(p, kind, src, p->kpad, 1, p->d_tab_kl, p->Lj, 0, dst, Ploc, 2 * C, maxlen, B, C, st, m_lo, m_hi);
    default:
      return record_error(MSFNO_ERR_BAD_SHAPE, "legendre_stage: bad kind");
  }
}

int msfno_lat_segments(int gather, float* flat, float* full, long rows, int pad_full, int nlat, int nseg, const int* seg_lo,
                       const int* seg_n, void* stream) {
  if (!flat || !full || rows < 1 || nlat < 1 || pad_full < nlat || nseg < 1 || nseg > MSFNO_MAX_LAT_SEGMENTS || !seg_lo || !seg_n)
    return record_error(MSFNO_ERR_BAD_SHAPE, "lat_segments: bad argument");
  LatSegs sg{};
  sg.n = nseg;
  long long base = 0;
  for (int s = 0; s < nseg; ++s) {
    if (seg_n[s] < 0 || seg_lo[s] < 0 || seg_lo[s] + seg_n[s] > nlat) return record_error(MSFNO_ERR_BAD_SHAPE, "lat_segments: bad segment");
    sg.lo[s] = seg_lo[s]; sg.cnt[s] = seg_n[s]; sg.base[s] = base;
    base += (long long)rows * ((seg_n[s] + 31) & ~31);
  }
  long long blocks = (rows + 7) / 8;
  if (blocks > 148 * 8) blocks = 148 * 8;
  lat_segments_kernel<<<dim3((unsigned)blocks, nseg + (gather ? 1 : 0)), 256, 0, (cudaStream_t)stream>>>(gather, flat, full, rows, pad_full,
                                                                                                       nlat, sg);
  count_launch();
  MSFNO_CUDA_OK(cudaGetLastError());
  return MSFNO_OK;
}

}  // extern "C"
