// SpectralAttentionS2 mode-shared complex MLP (forward + backward) as a chain of real GEMMs.
//
// replaces: SpectralAttentionS2.forward_mlp (/root/reference MSFNO/Models/sfno/layers.py:604-620):
//   spectral_layers x [ einsum("bixy,io->boxy") on complex64 (contractions.py:132-137)
//                       + ComplexReLU mode "real" (activations.py:42-46: clone + slice-assign) ]
//   followed by the `wout` product, and the autograd graph PyTorch builds for them.
//
// A complex product with interleaved (re, im) activations is one real GEMM against the packed weight
//   Wbig[2o+0][2i+0] = wr   Wbig[2o+0][2i+1] = -wi
//   Wbig[2o+1][2i+0] = wi   Wbig[2o+1][2i+1] =  wr
// so ComplexReLU("real") is "ReLU on even output columns" in the GEMM epilogue, and the structurally
// zero l < m positions are never touched because the activations live in the packed PM layout.
#include "common.cuh"
#include "plan.h"

namespace msfno {

// w [Ci][Co][2] -> wbig [2Co][2Ci]
__device__ __forceinline__ float rna_tf32(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;\n" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}

__global__ void pack_cweight_kernel(const float* __restrict__ w, float* __restrict__ wbig, int Ci, int Co, int round_tf32) {
  const int total = Ci * Co;
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
    const int i = idx % Ci, o = idx / Ci;  // consecutive threads -> consecutive i (coalesced writes)
    float2 v = *reinterpret_cast<const float2*>(w + ((size_t)i * Co + o) * 2);
    if (round_tf32) { v.x = rna_tf32(v.x); v.y = rna_tf32(v.y); }
    float* r0 = wbig + (size_t)(2 * o) * (2 * Ci) + 2 * i;
    float* r1 = r0 + 2 * Ci;
    r0[0] = v.x; r0[1] = -v.y;
    r1[0] = v.y; r1[1] = v.x;
  }
}
// gwbig [2Co][2Ci] -> gw [Ci][Co][2]
__global__ void unpack_cweight_grad_kernel(const float* __restrict__ g, float* __restrict__ gw, int Ci, int Co) {
  const int total = Ci * Co;
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
    const int o = idx % Co, i = idx / Co;
    const float* r0 = g + (size_t)(2 * o) * (2 * Ci) + 2 * i;
    const float* r1 = r0 + 2 * Ci;
    float2 v;
    v.x = r0[0] + r1[1];
    v.y = r1[0] - r0[1];
    *reinterpret_cast<float2*>(gw + ((size_t)i * Co + o) * 2) = v;
  }
}

// engine of a GEMM of this file: 0 = CUDA-core FFMA, 1 = plain TF32 MMA, 3 = 3xTF32 (fp32 grade)
static int engine_of(int tf32) { return tf32 ? 1 : (fp32_engine_x3() ? 3 : 0); }

// One GEMM (single problem or strided batch described by g) on the selected engine; falls back to FFMA when TMA cannot
// address the operands.  a_rows .. b_cols: the 2-D buffers behind A and B (launch_gemm_tc).
static int gemm_on(int engine, GemmLaunch& g, long long a_rows, long long a_cols, long long b_rows, long long b_cols,
                   int round_tf32, cudaStream_t st) {
  if (engine && gemm_tc_supported(g)) {
    g.x3 = engine == 3;
    return launch_gemm_tc(g, a_rows, a_cols, b_rows, b_cols, engine == 1 ? round_tf32 : 0, st);
  }
  return launch_gemm_ffma(g, st);
}

static GemmLaunch single_problem(const float* A, long long lda, int a_k, const float* Bm, long long ldb, int b_k, float* D,
                                 long long ldd, int M, int N, int K) {
  GemmLaunch g{};
  g.A = A; g.B = Bm; g.D = D;
  g.lda = lda; g.ldb = ldb; g.ldd = ldd;
  g.a_kmajor = a_k; g.b_kmajor = b_k;
  g.ngroups = 1; g.maxM = M; g.maxN = N;
  g.use_single = 1;
  g.single = GemmGroup{0, 0, 0, M, N, K, 0};
  return g;
}

// D[M][N] = A[M][K] * B[N][K]^T on the selected engine
static int gemm_nt_any(int engine, const float* A, long long lda, const float* Bm, long long ldb, float* D, long long ldd, int M,
                       int N, int K, int relu_even, int round_tf32, cudaStream_t st) {
  GemmLaunch g = single_problem(A, lda, 1, Bm, ldb, 1, D, ldd, M, N, K);
  g.relu_even = relu_even;
  return gemm_on(engine, g, M, K, N, K, round_tf32, st);
}

struct AttnWs {
  size_t wbig[16];  // offsets of packed hidden-layer weights
  size_t wbig_out;
  size_t h[16];     // offsets of post-activation hidden states
  size_t total;
};
static AttnWs attn_layout(int P, int B, int C, int hid, int nl) {
  AttnWs L{};
  size_t o = 0;
  for (int l = 0; l < nl; ++l) {
    L.wbig[l] = o;
    o += (size_t)(2 * hid) * (2 * (l == 0 ? C : hid));
  }
  L.wbig_out = o;
  o += (size_t)(2 * C) * (2 * hid);
  o = (o + 63) & ~(size_t)63;
  for (int l = 0; l < nl; ++l) {
    L.h[l] = o;
    o += (size_t)B * P * 2 * hid;
  }
  L.total = o;
  return L;
}

}  // namespace msfno

using namespace msfno;

extern "C" {

size_t msfno_specattn_ws_floats(const msfno_plan* p, int B, int C, int hidden, int nlayers) {
  if (!p || nlayers < 1 || nlayers > 16) return 0;
  return attn_layout(p->P, B, C, hidden, nlayers).total;
}

size_t msfno_specattn_bwd_scratch_floats(const msfno_plan* p, int B, int C, int hidden, int nlayers) {
  if (!p) return 0;
  const size_t act = (size_t)B * p->P * 2 * hidden;
  const size_t wmax = (size_t)(2 * hidden) * (2 * (hidden > C ? hidden : C));
  return 2 * act + wmax + 64;
}

int msfno_specattn_fwd(const msfno_plan* p, const float* a_pm, const float* const* w, int nl, const float* wout,
                       float* out_cm, float* ws, int B, int C, int hid, int precision, void* stream) {
  if (!p || !a_pm || !w || !wout || !out_cm || !ws || nl < 1 || nl > 16) return record_error(MSFNO_ERR_BAD_SHAPE, "specattn_fwd: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  const AttnWs L = attn_layout(p->P, B, C, hid, nl);
  const int rows = B * p->P;
  const int skip_pack = (precision >> 2) & 1;   // bit 2: ws already holds the packed weights of these parameters
  const int tc = ((precision & 1) == MSFNO_PREC_TF32);   // TF32-round the packed weights and the hidden states
  const int engine = engine_of(tc);
  const float* in = a_pm;
  int cin = C;
  for (int l = 0; l < nl; ++l) {
    if (!skip_pack) {
      pack_cweight_kernel<<<(cin * hid + 255) / 256, 256, 0, st>>>(w[l], ws + L.wbig[l], cin, hid, tc);
      count_launch();
    }
    int rc = gemm_nt_any(engine, in, 2 * cin, ws + L.wbig[l], 2 * cin, ws + L.h[l], 2 * hid, rows, 2 * hid, 2 * cin,
                         /*relu_even=*/1, /*round_tf32=*/tc, st);
    if (rc) return rc;
    in = ws + L.h[l];
    cin = hid;
  }
  if (!skip_pack) {
    pack_cweight_kernel<<<(hid * C + 255) / 256, 256, 0, st>>>(wout, ws + L.wbig_out, hid, C, tc);
    count_launch();
  }
  // out_cm[b][ch][p] = sum_k wbig_out[ch][k] * h[b*P + p][k]   (strided batch over b)
  GemmLaunch g{};
  g.A = ws + L.wbig_out; g.B = in; g.D = out_cm;
  g.lda = 2 * hid; g.ldb = 2 * hid; g.ldd = p->P;
  g.a_kmajor = 1; g.b_kmajor = 1;
  g.ngroups = B; g.maxM = 2 * C; g.maxN = p->P;
  g.use_single = 1;
  g.single = GemmGroup{0, 0, 0, 2 * C, p->P, 2 * hid, 0};
  g.sa = 0; g.sb = (long long)p->P * 2 * hid; g.sd = (long long)2 * C * p->P;
  int rc = gemm_on(engine, g, 2 * C, 2 * hid, (long long)B * p->P, 2 * hid, /*round_tf32=*/1, st);
  if (rc) return rc;
  MSFNO_CUDA_OK(cudaGetLastError());
  return MSFNO_OK;
}

int msfno_specattn_bwd(const msfno_plan* p, const float* a_pm, const float* g_cm, const float* ws, float* ga_pm,
                       float* const* gw, float* gwout, float* scratch, int nl, int B, int C, int hid, void* stream) {
  // gwout and the entries of gw may be NULL: that weight gradient is not wanted (frozen backbone) and its GEMM is skipped
  if (!p || !a_pm || !g_cm || !ws || !ga_pm || !gw || !scratch || nl < 1 || nl > 16)
    return record_error(MSFNO_ERR_BAD_SHAPE, "specattn_bwd: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  const AttnWs L = attn_layout(p->P, B, C, hid, nl);
  const int P = p->P, rows = B * P;
  const size_t act = (size_t)rows * 2 * hid;
  float* gz[2] = {scratch, scratch + act};
  float* gwbig = scratch + 2 * act;
  const float* hlast = ws + L.h[nl - 1];

  // every GEMM of the backward runs at fp32 grade: 3xTF32 on the tensor cores, or FFMA (msfno_set_fp32_engine)
  const int engine = engine_of(0);

  // ---- output layer: out[b][ch][p] = sum_k Wout_big[ch][k] h[b*P+p][k]
  // gWout_big[ch][k] = sum_{b,p} g_cm[b][ch][p] * h[b*P+p][k]
  if (gwout) {
    for (int b = 0; b < B; ++b) {
      GemmLaunch g = single_problem(g_cm + (size_t)b * 2 * C * P, P, 1, hlast + (size_t)b * P * 2 * hid, 2 * hid, 0, gwbig,
                                    2 * hid, 2 * C, 2 * hid, P);
      g.accumulate = b > 0;
      int rc = gemm_on(engine, g, 2 * C, P, P, 2 * hid, 0, st);
      if (rc) return rc;
    }
    unpack_cweight_grad_kernel<<<(hid * C + 255) / 256, 256, 0, st>>>(gwbig, gwout, hid, C);
    count_launch();
  }
  // gz_last[b*P+p][k] = relu'(h) * sum_ch g_cm[b][ch][p] * Wout_big[ch][k]
  {
    GemmLaunch g{};
    g.A = g_cm; g.B = ws + L.wbig_out; g.D = gz[0];
    g.lda = P; g.ldb = 2 * hid; g.ldd = 2 * hid;
    g.a_kmajor = 0; g.b_kmajor = 0;
    g.mask = hlast; g.ldmask = 2 * hid;
    g.ngroups = B; g.maxM = P; g.maxN = 2 * hid;
    g.use_single = 1;
    g.single = GemmGroup{0, 0, 0, P, 2 * hid, 2 * C, 0};
    g.sa = (long long)2 * C * P; g.sb = 0; g.sd = (long long)P * 2 * hid;
    int rc = gemm_on(engine, g, (long long)B * 2 * C, P, 2 * C, 2 * hid, 0, st);
    if (rc) return rc;
  }
  int cur = 0;
  for (int l = nl - 1; l >= 0; --l) {
    const int cin = (l == 0) ? C : hid;
    const float* in = (l == 0) ? a_pm : ws + L.h[l - 1];
    // gWbig_l[hidcol][cincol] = sum_rows gz[row][hidcol] * in[row][cincol]
    int rc = MSFNO_OK;
    if (gw[l]) {
      GemmLaunch g = single_problem(gz[cur], 2 * hid, 0, in, 2 * cin, 0, gwbig, 2 * cin, 2 * hid, 2 * cin, rows);
      rc = gemm_on(engine, g, rows, 2 * hid, rows, 2 * cin, 0, st);
      if (rc) return rc;
      unpack_cweight_grad_kernel<<<(cin * hid + 255) / 256, 256, 0, st>>>(gwbig, gw[l], cin, hid);
      count_launch();
    }
    // g_in[row][cincol] = sum_hidcol gz[row][hidcol] * Wbig_l[hidcol][cincol]  (masked by ReLU of the layer below)
    float* dst = (l == 0) ? ga_pm : gz[cur ^ 1];
    GemmLaunch g = single_problem(gz[cur], 2 * hid, 1, ws + L.wbig[l], 2 * cin, 0, dst, 2 * cin, rows, 2 * cin, 2 * hid);
    if (l > 0) { g.mask = in; g.ldmask = 2 * cin; }
    rc = gemm_on(engine, g, rows, 2 * hid, 2 * hid, 2 * cin, 0, st);
    if (rc) return rc;
    cur ^= 1;
  }
  MSFNO_CUDA_OK(cudaGetLastError());
  return MSFNO_OK;
}

int msfno_gemm_nt(const float* A, long lda, const float* Bm, long ldb, float* D, long ldd, int M, int N, int K,
                  int relu_even_cols, int precision, void* stream) {
  if (!A || !Bm || !D || M < 0 || N < 0 || K < 0) return record_error(MSFNO_ERR_BAD_SHAPE, "gemm_nt: bad argument");
  return gemm_nt_any(engine_of(precision == MSFNO_PREC_TF32), A, lda, Bm, ldb, D, ldd, M, N, K, relu_even_cols, 0, (cudaStream_t)stream);
}

int msfno_gemm_ex(const float* A, long lda, int a_kmajor, const float* Bm, long ldb, int b_kmajor, float* D, long ldd, int M,
                  int N, int K, int relu_even_cols, const float* mask, long ldmask, int accumulate, int engine, void* stream) {
  if (!A || !Bm || !D || M < 0 || N < 0 || K < 0 || (engine != 0 && engine != 1 && engine != 3))
    return record_error(MSFNO_ERR_BAD_SHAPE, "gemm_ex: bad argument");
  GemmLaunch g = single_problem(A, lda, a_kmajor, Bm, ldb, b_kmajor, D, ldd, M, N, K);
  g.relu_even = relu_even_cols; g.mask = mask; g.ldmask = ldmask; g.accumulate = accumulate;
  return gemm_on(engine, g, a_kmajor ? M : K, a_kmajor ? K : M, b_kmajor ? N : K, b_kmajor ? K : N, 0, (cudaStream_t)stream);
}

}  // extern "C"
