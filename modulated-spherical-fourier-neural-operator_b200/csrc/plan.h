// Plan: everything that depends only on the transform geometry (and the caller's Legendre table):
// twiddles, per-order scale vectors, packed-position maps, re-laid tables, GEMM group lists.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <map>
#include <mutex>
#include <utility>
#include <vector>

#include "../../include/msfno_b200.h"
#include "fft_core.cuh"

namespace msfno {

// One member of a grouped GEMM: offsets (in floats) into the three base pointers + extents.
struct GemmGroup {
  long long a_off, b_off, d_off;
  int M, N, K, pad;
};

}  // namespace msfno

struct msfno_plan {
  int nlat, nlon, lmax, mmax;
  int device;
  int precision = 0;  // MSFNO_PREC_*: tier of the Legendre contractions (forward transforms only)
  int mlim;   // min(mmax, lmax): orders with at least one degree
  int kpad;   // nlat rounded up to 32
  int Lj;     // padded per-order degree extent (ceil4(lmax))
  int P;      // packed positions
  int ntril;  // reference tril count
  msfno::FftSchedule sched;
  std::vector<int32_t> h_poff, h_plen4, h_n2p;
  // device buffers
  float* d_tw = nullptr;    // [H] cf exp(-2 pi i t / H)
  float* d_tw2 = nullptr;   // [mlim+1] cf exp(-2 pi i m / nlon)
  float* d_scale_rfft = nullptr;      // forward kernel, RealSHT.forward: 2 pi / nlon
  float* d_scale_irfft_adj = nullptr; // forward kernel, adjoint of irfft: c_m
  float* d_scale_irfft = nullptr;     // inverse kernel, irfft(norm="forward"): 1
  float* d_scale_rfft_adj = nullptr;  // inverse kernel, adjoint of 2 pi rfft(norm="forward")
  int32_t* d_poff = nullptr;          // [mmax]
  int32_t* d_n2p = nullptr;           // [ntril]
  int32_t* d_p2lm = nullptr;          // [P] l * mmax + m, or -1 for pad slots
  float* d_tab_lk = nullptr;          // analysis table  [mlim][Lj][kpad]   (contract over lat)
  float* d_tab_kl = nullptr;          // synthesis table [mlim][nlat][Lj]   (contract over degree)
  int32_t* d_flag = nullptr;          // table validation flag
  float* d_dft_fwd = nullptr;         // tensor-core tier: DFT matrix  [2 mlim][nlon]  (dft_tc.cu, built on first use)
  float* d_dft_inv = nullptr;         //                   inverse DFT [nlon][2 mlim]
  float* d_dft_inv_eo = nullptr;      //                   inverse DFT split by the parity of m: [2][nlon / 2][128] (idft_eo_kernel)
  // grouped-GEMM descriptors cached per (kind, B, C)
  std::mutex mu;
  std::map<std::vector<int>, msfno::GemmGroup*> groups;  // key: kind, B, C, m_lo, m_hi
};

namespace msfno {

enum GroupKind { GK_ANALYSIS = 0, GK_ANALYSIS_ADJ = 1, GK_SYNTHESIS = 2, GK_SYNTHESIS_ADJ = 3 };
// returns a device array of B*mlim groups (cached)
int plan_groups(msfno_plan* p, int kind, int B, int C, const GemmGroup** out, int* ngroups, int m_lo = 0, int m_hi = -1);

int launch_rfft_trunc(const msfno_plan* p, const float* x, float* Xt, const float* mscale, int zero_imag,
                      const float* in_scale, const float* in_shift, int B, int C, cudaStream_t st);
int launch_irfft_trunc(const msfno_plan* p, const float* Yt, float* y, const float* mscale, const float* skip,
                       const float* out_scale, int act_gelu, double* stats, int B, int C, cudaStream_t st);

// Grouped fp32 GEMM on CUDA cores.  D = opA(A) * opB(B)^T, row-major D.
//   a_kmajor: A element (m,k) at A[m*lda + k]; otherwise at A[k*lda + m]   (same for B with n)
struct GemmLaunch {
  const float* A; const float* B; float* D;
  long long lda, ldb, ldd;
  int a_kmajor, b_kmajor;
  int relu_even;            // ReLU on even output columns (real parts of interleaved complex)
  const float* mask;        // optional, same layout/offsets as D: zero D[m][n] (even n) where mask[m][n] <= 0 (ReLU backward)
  long long ldmask;
  int accumulate;           // D += result
  // fused epilogue of the 1x1-conv form: D = act(acc + bias[row]) + add[row][col]
  const float* bias;        // per output row (+ group_index * sbias)
  long long sbias;
  const float* add;         // same row/col indexing as D with leading dimension ldadd (+ group_index * sadd)
  long long ldadd, sadd;
  int act_gelu;
  // optional second operand pair accumulated into the same D (K-concatenation without a concat copy)
  const float* A2; const float* B2;
  long long lda2, ldb2, sa2, sb2;
  int K2;
  const GemmGroup* groups;  // device array
  int ngroups;
  int maxM, maxN;           // over groups (grid sizing)
  int use_single;           // 1: ignore `groups`; group y = `single` shifted by y * (sa, sb, sd) (strided batch)
  GemmGroup single;
  long long sa, sb, sd;
  // tensor-core engine only (gemm_tc.cu)
  int x3;                   // fp32 tier: three TF32 MMAs per k-step on hi / lo operand splits ("3xTF32")
  int b_group_rows;         // > 0 with an MN-major B: B is a stack of tables of this many rows each; K rows past the end of
                            // a table read zeros (3-D tensor map) instead of the next table
};
int launch_gemm_ffma(const GemmLaunch& g, cudaStream_t st);
int launch_gemm_single(const float* A, long long lda, int a_kmajor, const float* B, long long ldb, int b_kmajor,
                       float* D, long long ldd, int M, int N, int K, int relu_even, const float* mask,
                       long long ldmask, int accumulate, cudaStream_t st);

// tcgen05 / TMEM / TMA path (gemm_tc.cu): TF32 (one MMA per k-step) or, with g.x3, fp32-grade 3xTF32.  Supported when the
// operands are 16-byte aligned with leading dimensions that are multiples of 4; either majorness, ReLU-mask and
// accumulate epilogues included.  a_rows/a_cols (b_rows/b_cols) describe the whole 2-D buffer behind A (B).
// fp32_engine_x3(): true unless msfno_set_fp32_engine(MSFNO_FP32_ENGINE_FFMA) selected the CUDA-core engine for the
// fp32 tier (kept for shapes the tensor-core path rejects, and as the cross-check in the tests).
bool gemm_tc_supported(const GemmLaunch& g);
bool fp32_engine_x3();
int launch_gemm_tc(const GemmLaunch& g, long long a_rows, long long a_cols, long long b_rows, long long b_cols,
                   int round_tf32, cudaStream_t st, long long a2_rows = 0, long long a2_cols = 0, long long b2_rows = 0,
                   long long b2_cols = 0);

// four-step in-register FFT kernels (fft2d.cu) for nlon in {240, 1440, 2880}
bool fft2d_supported(int nlon);
// tensor-core tier longitude transforms (dft_tc.cu)
bool dft_tc_supported(const msfno_plan* p);
int launch_dft_fwd(msfno_plan* p, const float* x, float* Xt, const float* in_scale, const float* in_shift, int B, int C,
                   cudaStream_t st);
int launch_dft_inv(msfno_plan* p, const float* Yt, float* y, const float* skip, int act_flags, double* stats, int B, int C,
                   cudaStream_t st);
// Where the lat-contiguous intermediate of a SHARDED transform lives (spatial decomposition, B = 1): order m belongs to
// rank s with mb[s] <= m < mb[s+1], whose operand buffer buf[s] is [mb[s+1]-mb[s]][2C][pitch] over ALL latitudes (a CUDA
// IPC mapping for s != this rank); this rank's latitudes start at column lat_lo.  world == 0: the local [mlim][2C][kpad].
struct PeerMapDev {
  int world;
  int mb[MSFNO_MAX_PEERS + 1];
  float* buf[MSFNO_MAX_PEERS];
  int pitch;
  int lat_lo;
};
int launch_rfft2d(const msfno_plan* p, const float* x, float* Xt, const float* mscale, int zero_imag,
                  const float* in_scale, const float* in_shift, int B, int C, cudaStream_t st, const PeerMapDev* pm = nullptr);
int launch_irfft2d(const msfno_plan* p, const float* Yt, float* y, const float* mscale, const float* skip,
                   const float* out_scale, int act_gelu, double* stats, int B, int C, cudaStream_t st,
                   const PeerMapDev* pm = nullptr);

// persistent weight-stationary conv kernel (conv_tc.cu); *handled = 0 -> caller falls back to launch_gemm_tc
int launch_conv_tc(const GemmLaunch& g, long long a_rows, long long a_cols, long long b_rows, long long b_cols,
                   long long a2_rows, long long a2_cols, long long b2_rows, long long b2_cols, int* handled,
                   cudaStream_t st, int round_tf32 = 0);

}  // namespace msfno
