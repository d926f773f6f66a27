// 1x1 convolution in NCHW as one GEMM with a fused epilogue -- the channel MLPs either side of the spectral path
// (SURVEY.md 8(f) N2): encoder / decoder / block MLP / inner skip.
//
// replaces: nn.Conv2d(cin, cout, 1) [+ bias] [+ nn.GELU] [+ residual / pos_embed add] [+ torch.cat of the big skip]
//   (/root/reference MSFNO/Models/sfno/layers.py:161-168; sfnonet.py:184-185,232,249,671,682-684): cuDNN/cuBLAS GEMM plus
//   up to three separate full-tensor elementwise passes and a 1.4 GB concat copy per call at full resolution.
//
//   y[b][o][p] = act( sum_c w[o][c] x[b][c][p] + sum_c w2[o][c] x2[b][c][p] + bias[o] ) + add[b][o][p]
// The activation tensor is used in place as the MN-major GEMM operand ([K = channels][N = pixels], pixels contiguous),
// so no layout change is needed; the optional second pair (w2, x2) accumulates into the same tile (big-skip concat).
#include "common.cuh"
#include "plan.h"

using namespace msfno;

extern "C" int msfno_conv1x1_fwd(const float* x, long x_bstride, int Cin, const float* w, long ldw, long w_bstride,
                                 const float* x2, long x2_bstride, int Cin2, const float* w2, long ldw2,
                                 const float* bias, long bias_bstride, const float* add, long add_bstride, float* y,
                                 int B, int Cout, long HW, int act_gelu, int precision, void* stream) {
  // bit 1 of `precision`: round the outputs to TF32 (they feed another tensor-core GEMM)
  const int round_out = (precision >> 1) & 1;
  precision &= 1;
  if (!x || !w || !y || B < 1 || Cin < 1 || Cout < 1 || HW < 1 || ldw < Cin || (x2 && (!w2 || Cin2 < 1 || ldw2 < Cin2)) ||
      act_gelu < 0 || act_gelu > 2 || (act_gelu == 2 && !add))
    return record_error(MSFNO_ERR_BAD_SHAPE, "conv1x1_fwd: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  GemmLaunch g{};
  g.A = w; g.lda = ldw; g.a_kmajor = 1;
  g.B = x; g.ldb = HW; g.b_kmajor = 0;
  g.D = y; g.ldd = HW;
  g.bias = bias; g.sbias = bias_bstride;
  g.add = add; g.ldadd = HW; g.sadd = add_bstride;
  g.act_gelu = act_gelu;
  g.ngroups = B; g.maxM = Cout; g.maxN = (int)HW;
  g.use_single = 1;
  g.single = GemmGroup{0, 0, 0, Cout, (int)HW, Cin, 0};
  g.sa = w_bstride; g.sb = x_bstride; g.sd = (long long)Cout * HW;
  if (x2) {
    g.A2 = w2; g.B2 = x2; g.lda2 = ldw2; g.ldb2 = HW; g.sa2 = 0; g.sb2 = x2_bstride; g.K2 = Cin2;
  }
  if (HW > 0x7fffffffL) return record_error(MSFNO_ERR_UNSUPPORTED, "conv1x1_fwd: plane too large");
  const bool tf32 = precision == MSFNO_PREC_TF32;
  if ((tf32 || fp32_engine_x3()) && (HW % 4 == 0) && (x_bstride % HW == 0) && (!x2 || x2_bstride % HW == 0) &&
      (w_bstride % ldw == 0) && gemm_tc_supported(g)) {
    const long long a_rows = (long long)(B - 1) * (w_bstride / ldw) + Cout;
    const long long b_rows = (long long)(B - 1) * (x_bstride / HW) + Cin;
    const long long b2_rows = x2 ? (long long)(B - 1) * (x2_bstride / HW) + Cin2 : 0;
    if (tf32) {   // persistent weight-stationary kernel (plain TF32 only)
      int handled = 0;
      int rc = launch_conv_tc(g, a_rows, ldw, b_rows, HW, Cout, ldw2, b2_rows, HW, &handled, st, round_out);
      if (rc || handled) return rc;
    }
    g.x3 = tf32 ? 0 : 1;   // fp32 tier: 3xTF32 on the grouped tensor-core GEMM
    return launch_gemm_tc(g, a_rows, ldw, b_rows, HW, tf32 ? round_out : 0, st, Cout, ldw2, b2_rows, HW);
  }
  return launch_gemm_ffma(g, st);
}
