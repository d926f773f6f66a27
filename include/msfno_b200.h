/* msfno_b200.h -- C ABI of the B200-native MSFNO spectral hot path (libmsfno_b200.so).
 *
 * The reference (Slusny/Modulated-Spherical-Fourier-Neural-Operator) has no native code and no
 * FFI: its hot path is a sequence of PyTorch library calls behind torch.nn.Module classes.
 * Each entry point below replaces one such sequence; the "replaces" line gives the reference
 * interface (paths relative to /root/reference).  The Python package
 * modulated-spherical-fourier-neural-operator_b200/ (import name msfno_b200) mirrors the
 * reference classes and calls these functions through ctypes with raw device pointers.
 *
 * Conventions
 *  - all pointers are DEVICE pointers (fp32 unless stated) owned by the caller; the library
 *    never frees caller memory and allocates only plan-lifetime buffers (tables, twiddles);
 *  - `stream` is a cudaStream_t passed as void*; every call is asynchronous on it, performs
 *    no allocation or synchronisation and is CUDA-graph capturable (plan_* calls excepted);
 *  - return value 0 = MSFNO_OK, otherwise an MSFNO_ERR_* code; msfno_last_error() returns a
 *    thread-local human-readable message (includes cudaGetErrorString for CUDA errors);
 *  - complex data are interleaved (re, im) pairs; "channel index" ch = 2*c + ri.
 *
 * Internal coefficient layouts (B = batch, C = channels, P = packed triangular positions):
 *  - STD : [B][C][lmax][mmax][2]      torch complex64 layout of the reference
 *  - PM  : [B][P][2C]                 position-major (channels contiguous)
 *  - CM  : [B][2C][P]                 channel-major  (positions contiguous)
 *  Packed position p = poff[m] + (l - m) for l >= m; each order m owns ceil4(lmax - m) slots
 *  (pad slots hold zeros).
 */
#ifndef MSFNO_B200_H
#define MSFNO_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MSFNO_OK 0
#define MSFNO_ERR_BAD_SHAPE 1   /* shape / argument rejected (mirrors the asserts of torch_harmonics) */
#define MSFNO_ERR_UNSUPPORTED 2 /* size this build does not support (e.g. nlon with a prime factor > 5) */
#define MSFNO_ERR_CUDA 3        /* a CUDA runtime call failed; see msfno_last_error() */
#define MSFNO_ERR_BAD_STATE 4   /* plan used before its table was set */

#define MSFNO_LAYOUT_STD 0
#define MSFNO_LAYOUT_PM 1
#define MSFNO_LAYOUT_CM 2

/* msfno_plan_query selectors */
#define MSFNO_Q_KPAD 0   /* padded latitude pitch of the lat<->m intermediates (multiple of 32) */
#define MSFNO_Q_MLIM 1   /* number of azimuthal orders carrying data: min(mmax, lmax) */
#define MSFNO_Q_NPACK 2  /* P: packed positions per (b) plane */
#define MSFNO_Q_NTRIL 3  /* n: len(torch.tril_indices(lmax, mmax)[0]) of the reference */
#define MSFNO_Q_LJ 4     /* padded degree extent of the re-laid tables */

/* precision tiers */
#define MSFNO_PREC_FP32 0 /* the 1e-5 tier: fp32-grade products, fp32 accumulate (engine: msfno_set_fp32_engine) */
#define MSFNO_PREC_TF32 1 /* the 2e-3 tier: tcgen05 kind::tf32, one MMA per k-step on TF32-rounded operands */
/* engines of the fp32 tier's GEMM-shaped stages */
#define MSFNO_FP32_ENGINE_TC3X 0 /* tcgen05 kind::tf32, three MMAs per k-step on hi / lo operand splits ("3xTF32") */
#define MSFNO_FP32_ENGINE_FFMA 1 /* CUDA-core FFMA kernels */

typedef struct msfno_plan msfno_plan;

const char* msfno_last_error(void);
/* compile-time facts: "sm_100a", has_tcgen05 etc. as a static JSON string */
const char* msfno_build_info(void);
/* number of kernels this library has launched so far in this process (bench.py's gpu_launches) */
unsigned long long msfno_launch_count(void);
/* Process-wide engine of the fp32 tier (default MSFNO_FP32_ENGINE_TC3X).  The reference forces fp32 arithmetic inside
 * the transforms and the spectral MLP (MSFNO/Models/sfno/layers.py:403-407,418-422,627-639); both engines meet that
 * tier's 1e-5 rel-L2, the FFMA one is kept as the cross-check and for operands the TMA path cannot address. */
int msfno_set_fp32_engine(int engine);
int msfno_get_fp32_engine(void);

/* ---- plans ------------------------------------------------------------------------------
 * replaces: torch_harmonics.RealSHT.__init__ / InverseRealSHT.__init__ device-side state
 *           (call sites MSFNO/Models/sfno/sfnonet.py:537-548); the Legendre tables themselves
 *           stay caller-owned torch buffers (`weights`, `pct`) so the reference's in-place
 *           rescale (sfnonet.py:551-555) keeps working: call msfno_plan_set_table again
 *           whenever the buffer is reassigned or mutated. */
int msfno_plan_create(msfno_plan** plan, int nlat, int nlon, int lmax, int mmax);
int msfno_plan_destroy(msfno_plan* plan);
long msfno_plan_query(const msfno_plan* plan, int what);
/* table: [mmax][lmax][nlat] fp32 (reference layout).  analysis != 0 -> used by sht_fwd/bwd
 * (RealSHT.weights); analysis == 0 -> used by isht_fwd/bwd (InverseRealSHT.pct).
 * Entries with l < m must be zero (they are by construction); otherwise MSFNO_ERR_UNSUPPORTED. */
int msfno_plan_set_table(msfno_plan* plan, const float* table, int analysis, void* stream);
/* precision tier of the longitude transforms and Legendre contractions of the FORWARD transforms: MSFNO_PREC_FP32
 * (default) or MSFNO_PREC_TF32.  The adjoints always run at fp32 grade (3xTF32 tensor-core GEMMs or FFMA). */
int msfno_plan_set_precision(msfno_plan* plan, int precision);
/* host copies of the packed-position maps: poff[mmax] and n2p[ntril] (reference tril order) */
int msfno_plan_get_maps(const msfno_plan* plan, int32_t* poff_host, int32_t* n2p_host);

/* ---- a2 / a14: forward SHT and its adjoint ------------------------------------------------
 * replaces: torch_harmonics.RealSHT.forward  (2*pi*rfft(norm="forward") + 2 einsum
 *           '...km,mlk->...lm'), called from MSFNO/Models/sfno/layers.py:405 and :629.
 * x [B][C][nlat][nlon] -> coef_pm (PM layout).  Optional fused prologue x*in_scale[b*C+c] +
 * in_shift[b*C+c] (InstanceNorm folded into the transform, sfnonet.py:224/362).
 * ws: >= msfno_sht_ws_floats() floats of scratch (the lat<->m intermediate). */
size_t msfno_sht_ws_floats(const msfno_plan* plan, int B, int C);
int msfno_sht_fwd(msfno_plan* plan, const float* x, const float* in_scale, const float* in_shift,
                  float* coef_pm, float* ws, int B, int C, void* stream);
/* adjoint: g_pm (PM) -> gx [B][C][nlat][nlon]; in_scale (optional) multiplies gx per plane. */
int msfno_sht_bwd(msfno_plan* plan, const float* g_pm, const float* in_scale, float* gx, float* ws,
                  int B, int C, void* stream);

/* ---- a10 / a14: inverse SHT and its adjoint -----------------------------------------------
 * replaces: torch_harmonics.InverseRealSHT.forward (2 einsum '...lm,mlk->...km' + stack +
 *           irfft(n=nlon, norm="forward")), called from layers.py:421 and :638.
 * coef_cm (CM layout) -> y [B][C][nlat][nlon].  Fused epilogue, all optional:
 *   y = act( isht(coef) + skip_add ),  act = exact GELU when (act_gelu & 1); (act_gelu & 2) additionally rounds y to
 *   TF32 (round-to-nearest) because the next consumer is a tensor-core GEMM (tensor-core tier only)
 *   stats[b*C+c] += (sum y, sum y^2) in fp64 (feeds InstanceNorm norm1, sfnonet.py:237/376)
 * replaces additionally: `x + inner_skip(residual)` (sfnonet.py:232/371), act_layer (:235/374)
 * and the statistics pass of nn.InstanceNorm2d. */
int msfno_isht_fwd(msfno_plan* plan, const float* coef_cm, float* y, float* ws, int B, int C,
                   const float* skip_add, int act_gelu, double* stats, void* stream);
/* adjoint: gy [B][C][nlat][nlon] -> g_cm (CM layout). */
int msfno_isht_bwd(msfno_plan* plan, const float* gy, float* g_cm, float* ws, int B, int C, void* stream);

/* ---- stage-level entry points for the spatially sharded SHT (BASELINE config 5) ----------
 * The reference has no distributed transform (DDP only, main.py:39-49); these expose the two halves of
 * msfno_sht_* / msfno_isht_* separately so a lat<->m all-to-all can sit between them.
 * msfno_fft_stage: longitude transform only, on a plan whose nlat is the LOCAL latitude count.
 *   inverse = 0: x [B][C][nlat][nlon] -> Xt [B][mlim][2C][kpad]   (adjoint != 0: adjoint of irfft instead)
 *   inverse = 1: Yt -> y                                           (adjoint != 0: adjoint of rfft instead)
 * msfno_legendre_stage: Legendre contraction for the azimuthal orders [m_lo, m_hi) only; the lat<->m buffer
 *   holds just those orders ([B][m_hi-m_lo][2C][kpad]) and the PM / CM coefficient buffers just their packed
 *   positions [poff[m_lo], poff[m_hi]).  kind: 0 analysis (Xt -> PM), 1 its adjoint (PM -> Xt),
 *   2 synthesis (CM -> Yt), 3 its adjoint (Yt -> CM). */
int msfno_fft_stage(msfno_plan* plan, int inverse, int adjoint, const float* src, float* dst, int B, int C,
                    void* stream);
int msfno_legendre_stage(msfno_plan* plan, int kind, const float* src, float* dst, int m_lo, int m_hi, int B,
                         int C, void* stream);
/* The two ends of the lat<->m all-to-all, one launch each (they replaced one strided PyTorch copy per peer).
 * `flat`: the buffer all_to_all_single moves -- for every peer s a block [rows][pad32(seg_n[s])], rows = local orders x 2C,
 * seg_n[s] = the peer's latitude count (the pitch is the peer's own lat-contiguous pitch), blocks back to back;
 * `full`: the Legendre stage's operand [rows][pad_full] holding all nlat latitudes, peer s at columns seg_lo[s].
 * gather != 0: blocks -> full (columns >= nlat zeroed);  gather == 0: full -> blocks (pitch tails zeroed).
 * seg_lo / seg_n are HOST arrays of nseg <= MSFNO_MAX_LAT_SEGMENTS entries. */
#define MSFNO_MAX_LAT_SEGMENTS 16
int msfno_lat_segments(int gather, float* flat, float* full, long rows, int pad_full, int nlat, int nseg,
                       const int* seg_lo, const int* seg_n, void* stream);

/* ---- lat<->m exchange over NVLink peer memory (one process per GPU, CUDA IPC) -----------------------------------
 * msfno_peer_alloc: cudaMalloc'ed, zero-filled buffer plus its 64-byte IPC handle (ship the handle to the other ranks by
 * any means, e.g. torch.distributed.all_gather_object); msfno_peer_open maps a peer's buffer into this process
 * (msfno_peer_close unmaps it, msfno_peer_free releases an own buffer).
 * msfno_peer_block_copy: ONE launch that copies, for every block i, rows x cols floats from
 * src[(src_row0 + r) * src_pitch + src_col0 + j] to dst[(dst_row0 + r) * dst_pitch + dst_col0 + j] and zero-fills the
 * zero_tail columns behind them; dst is typically a peer mapping, so the transpose between the latitude-sharded longitude
 * stage and the order-sharded Legendre stage is a direct NVLink store into the consumer's operand buffer.
 * msfno_peer_barrier: all `world` ranks call it the same number of times; flags[r] is (this rank's mapping of) rank r's
 * flag array of `world` unsigned ints, zero at start; `state` is this rank's own device array of two unsigned ints, zero
 * at start: state[1] counts the barriers (the epoch is kept on the device, so a captured CUDA graph replays correctly),
 * state[0] is set if a peer never arrived (bounded spin instead of a hung GPU).  Completes when every rank has arrived
 * AND the peer stores issued before the barrier on each rank are visible.  `blocks` and `flags` are HOST arrays. */
#define MSFNO_MAX_PEERS 16
typedef struct msfno_peer_block {
  float* dst;
  long long rows;
  int cols, zero_tail;
  long long src_row0, src_col0, src_pitch;
  long long dst_row0, dst_col0, dst_pitch;
} msfno_peer_block;
/* The exchange FUSED into the longitude stage (B = 1, four-step FFT sizes): msfno_fft_stage_peer runs msfno_fft_stage on
 * this rank's latitudes with the lat-contiguous intermediate living in the ranks' peer-mapped operand buffers --
 * inverse = 0: x [C][nlat_loc][nlon] -> every (m, re/im) row segment is stored straight into buf[owner of m] (NVLink
 * stores from the FFT kernel's epilogue; no local intermediate, no copy kernel, no collective);
 * inverse = 1: the staging fill of the inverse FFT loads its segments from buf[owner of m] -> y [C][nlat_loc][nlon].
 * buf[s]: rank s's buffer [m_bounds[s+1] - m_bounds[s]][2C][pitch] over all latitudes, this rank's at column lat_lo.
 * Bracket with msfno_peer_barrier exactly like msfno_peer_block_copy. */
typedef struct msfno_peer_map {
  int world;
  int m_bounds[MSFNO_MAX_PEERS + 1];
  float* buf[MSFNO_MAX_PEERS];
  int pitch;
  int lat_lo;
} msfno_peer_map;
int msfno_fft_stage_peer(msfno_plan* plan, int inverse, const float* x, float* y, const msfno_peer_map* map, int C,
                         void* stream);
int msfno_peer_alloc(size_t bytes, void** ptr, void* handle64);
int msfno_peer_free(void* ptr);
int msfno_peer_open(const void* handle64, void** ptr);
int msfno_peer_close(void* ptr);
int msfno_peer_block_copy(const float* src, int nblocks, const msfno_peer_block* blocks, void* stream);
int msfno_peer_barrier(unsigned int* const* flags, int rank, int world, unsigned int* state, void* stream);

/* ---- coefficient layout changes (public boundary of RealSHT / InverseRealSHT) ------------
 * replaces: the zeros()+slice-assign in RealSHT.forward, view_as_real/complex shuffles and the
 * tril gather/scatter of SpectralConvS2.forward (layers.py:406-413). */
int msfno_coef_relayout(const msfno_plan* plan, const float* src, int src_layout, float* dst,
                        int dst_layout, int B, int C, void* stream);

/* ---- a4 / a5 / a14: SpectralConvS2 per-mode channel contraction ---------------------------
 * replaces: compl_contract_fwd_c  einsum("bin,kin->bkn") on complex64
 *           (MSFNO/Models/sfno/contractions.py:37-41 via layers.py:411) and its autograd.
 * w is the reference parameter layout [Co][Ci][n][2], n in torch.tril_indices order; it is
 * streamed exactly once per call.  Activations in and out are PM layout (a mode's channels
 * contiguous): a_pm (Ci channels) -> out_pm (Co); pad slots of the output are zero-filled.
 * msfno_coef_relayout(PM -> CM) feeds the result to msfno_isht_fwd.
 * ws: msfno_specconv_ws_floats() floats of scratch (the activations re-ordered to tril order so that every
 * operand is a contiguous run for the TMA engine); NULL selects the slower register-load kernel. */
size_t msfno_specconv_ws_floats(const msfno_plan* plan, int B, int Ci, int Co);
int msfno_specconv_fwd(const msfno_plan* plan, const float* a_pm, const float* w, float* out_pm, float* ws,
                       int B, int Ci, int Co, void* stream);
/* ga[b,i,n] = sum_k conj(w[k,i,n]) g[b,k,n] :  g_pm (PM, Co) -> ga_pm (PM, Ci) */
int msfno_specconv_bwd_x(const msfno_plan* plan, const float* g_pm, const float* w, float* ga_pm, float* ws,
                         int B, int Ci, int Co, void* stream);
/* gw[k,i,n] = sum_b conj(a[b,i,n]) g[b,k,n] :  a_pm, g_pm (PM) -> gw [Co][Ci][n][2] */
int msfno_specconv_bwd_w(const msfno_plan* plan, const float* a_pm, const float* g_pm, float* gw, float* ws,
                         int B, int Ci, int Co, void* stream);

/* ---- a6 / a7 / a8 / a14: SpectralAttentionS2 mode-shared complex MLP ----------------------
 * replaces: SpectralAttentionS2.forward_mlp (layers.py:604-620): spectral_layers x
 *           [compl_mul2d_fwd_c einsum("bixy,io->boxy") (contractions.py:132-137) +
 *           ComplexReLU mode "real" (activations.py:42-46)] + the `wout` product.
 * w[l]: [Cin_l][hidden][2] (l = 0: Cin = C, else hidden); wout: [hidden][C][2].
 * a_pm (PM, C channels) -> out_cm (CM, C channels).
 * ws: msfno_specattn_ws_floats() floats; after the call it holds the packed real weights and
 * the post-activation hidden states needed by msfno_specattn_bwd (keep it alive until then).
 * precision: MSFNO_PREC_*, optionally | 4 when `ws` still holds the packed weights of the same parameters from a
 * previous call (inference with frozen weights: skips the re-pack kernels). */
size_t msfno_specattn_ws_floats(const msfno_plan* plan, int B, int C, int hidden, int nlayers);
int msfno_specattn_fwd(const msfno_plan* plan, const float* a_pm, const float* const* w, int nlayers,
                       const float* wout, float* out_cm, float* ws, int B, int C, int hidden,
                       int precision, void* stream);
/* g_cm (CM) -> ga_pm (PM); gw[l] / gwout receive the weight gradients in the parameter layout; a NULL entry
 * (or gwout == NULL) skips that weight gradient -- frozen-backbone training needs only ga_pm.
 * scratch: msfno_specattn_bwd_scratch_floats() floats. */
size_t msfno_specattn_bwd_scratch_floats(const msfno_plan* plan, int B, int C, int hidden, int nlayers);
int msfno_specattn_bwd(const msfno_plan* plan, const float* a_pm, const float* g_cm, const float* ws,
                       float* ga_pm, float* const* gw, float* gwout, float* scratch, int nlayers,
                       int B, int C, int hidden, void* stream);

/* ---- a11 / a14: FiLM and the InstanceNorm it follows --------------------------------------
 * replaces: FiLM.forward (sfnonet.py:689-697): (1 + gamma*scale) * x + beta*scale with
 *           gamma, beta [B][C] broadcast over HW (einops.repeat materialisation removed). */
int msfno_film_affine_fwd(const float* x, const float* gamma, const float* beta, float scale, float* y,
                          int B, int C, long HW, void* stream);
/* gx = (1+gamma*scale)*gy; ggamma[b,c] = scale*sum(gy*x); gbeta[b,c] = scale*sum(gy). */
int msfno_film_affine_bwd(const float* gy, const float* x, const float* gamma, float scale, float* gx,
                          float* ggamma, float* gbeta, int B, int C, long HW, void* stream);
/* per-plane (sum, sum of squares) in fp64: stats[plane][2] (overwritten).
 * replaces: the reduction pass of nn.InstanceNorm2d (sfnonet.py:491-499). */
int msfno_plane_stats(const float* x, double* stats, int planes, long HW, void* stream);
/* Collapse InstanceNorm(eps, affine nw/nb) followed by optional FiLM(gamma, beta, scale) into one
 * per-plane affine y = A*x + S.  gamma/beta may be NULL (no FiLM). */
int msfno_norm_film_coeffs(const double* stats, const float* nw, const float* nb, const float* gamma,
                           const float* beta, float scale, float eps, float* A, float* S, int B, int C,
                           long HW, void* stream);
/* Fold the per-plane affine into the 1x1 conv that consumes it: Wb[b][o][c] = W[o][c] A[b][c] (TF32-rounded when
 * round_tf32), bb[b][o] = sum_c W[o][c] S[b][c] + bias[o] (bias may be NULL).  W: [O][ld] zero-padded rows. */
int msfno_fold_affine(const float* W, const float* A, const float* S, const float* bias, float* Wb, float* bb,
                      int B, int O, int C, int ld, int round_tf32, void* stream);
/* msfno_norm_film_coeffs + msfno_fold_affine in one launch (same arithmetic, bit-identical Wb / bb): the
 * InstanceNorm -> FiLM -> fc1 hand-over of every block (sfnonet.py:380-386) costs one small kernel instead of two. */
int msfno_fold_norm_affine(const float* W, const double* stats, const float* nw, const float* nb, const float* gamma,
                           const float* beta, float scale, float eps, long HW, const float* bias, float* Wb,
                           float* bb, int B, int O, int C, int ld, int round_tf32, void* stream);
/* Mean-carrying residual stream of the tensor-core tier (the stream stored between blocks is X = x - mu, mu [B][C] carried
 * beside it; InstanceNorm, sfnonet.py:221-251, is shift-invariant): from the plane sums `stats` [B C][2] of X (may be
 * NULL with b2 = mu_out = NULL) and the incoming offset mu (may be NULL = 0):
 *   b2[b][o] = bias2[o] - mean(X[b][o]);  mu_out[b][c] = mu[b][c] + mean(X[b][c]);
 *   sb[b][o] = skip_bias[o] + sum_c Wskip[o][c] mu[b][c]   (sb may be NULL; needs mu and Wskip [C][ldw]).
 * One launch instead of six library element-wise / gemv launches per block. */
int msfno_mean_carry(const double* stats, long HW, const float* mu, const float* bias2, const float* Wskip, long ldw,
                     const float* skip_bias, float* b2, float* mu_out, float* sb, int B, int C, void* stream);
/* out = g * gelu'(h), exact (erf) GELU: activation adjoint of the frozen-weight channel-MLP backward (a14);
 * round_tf32 != 0: rounded to TF32 (nearest) where it is made, as the tensor-core operand of the next GEMM */
int msfno_gelu_bwd_mul(const float* g, const float* h, float* out, long long n, int round_tf32, void* stream);
/* y[plane][:] = A[plane]*x[plane][:] + S[plane] */
int msfno_plane_affine(const float* x, const float* A, const float* S, float* y, int planes, long HW,
                       void* stream);

/* ---- 8(f) N3: latitude-weighted squared-error reductions of the spherical losses ----------------------------
 * replaces: L2Sphere / L2Sphere_noSine / CosineMSELoss.forward (MSFNO/Models/losses.py:6-37,80-155) and their autograd:
 *   out[plane] = ( sum_{h,w} wlat[h] (prd - tar)^2 ,  sum_{h,w} wlat[h] tar^2 )      (fp64, overwritten)
 * prd, tar: [planes][H][W]; wlat: [H] device vector (the reference rebuilds it on the host every call, :90,129). */
int msfno_weighted_sq_sums(const float* prd, const float* tar, const float* wlat, double* out, int planes, int H,
                           int W, void* stream);
/* gprd[plane][h][w] = coef[plane] * wlat[h] * (prd - tar): the adjoint of the reduction with the chain-rule factors of
 * the caller (2, 1/norm, 1/(2 sqrt(.)), upstream gradient) folded into coef. */
int msfno_weighted_diff(const float* prd, const float* tar, const float* wlat, const float* coef, float* gprd,
                        int planes, int H, int W, void* stream);

/* ---- 8(f) N2: 1x1 convolution (NCHW) with fused epilogue ------------------------------------------------
 * replaces: nn.Conv2d(cin, cout, 1) + bias + nn.GELU + residual/pos_embed add + torch.cat of the big skip
 *           (MSFNO/Models/sfno/layers.py:161-168; sfnonet.py:184-185,232,249,671,682-684).
 *   y[b][o][p] = act( sum_c w[o][c] x[b][c][p] + sum_c w2[o][c] x2[b][c][p] + bias[o] ) + add[b][o][p]
 * x: [B][Cin][HW] with batch stride x_bstride (floats); w: [Cout][ldw] row-major, ldw >= Cin, ldw % 4 == 0, columns
 * >= Cin zero; w_bstride != 0 selects per-sample weights (InstanceNorm/FiLM affine folded into the conv).
 * x2 / w2 (optional): second input accumulated into the same output (the decoder's concat of x and the big skip).
 * bias (optional, [Cout], bias_bstride 0 or Cout), add (optional, [Cout][HW], add_bstride 0 or Cout*HW).
 * act_gelu: 0 none, 1 GELU, 2 = activation ADJOINT: y = (conv + bias) * gelu'(add), `add` holding the pre-activation
 * (the W2^T g (.) gelu'(h) step of the frozen-weight MLP backward without a pass of its own).
 * precision: MSFNO_PREC_FP32 or MSFNO_PREC_TF32, optionally | 2 to round the outputs to TF32 (feeds another MMA). */
int msfno_conv1x1_fwd(const float* x, long x_bstride, int Cin, const float* w, long ldw, long w_bstride,
                      const float* x2, long x2_bstride, int Cin2, const float* w2, long ldw2, const float* bias,
                      long bias_bstride, const float* add, long add_bstride, float* y, int B, int Cout, long HW,
                      int act_gelu, int precision, void* stream);

/* ---- N2: fused two-layer 1x1-conv MLP (encoder / decoder at full resolution) ---------------
 * replaces: MLP.fwd = Conv2d(1x1) -> GELU -> Conv2d(1x1) (MSFNO/Models/sfno/layers.py:161-168) together with the
 *           pos_embed add (sfnonet.py:671) and the big-skip torch.cat (sfnonet.py:682-684).
 *   y[b][o][p] = sum_h w2[o][h] * gelu( sum_c w1[h][c] x[b][c][p] + sum_c w1b[h][c] x2[b][c][p] + b1[h] )
 *                + b2[o] + add[b][o][p]
 * The Chid-channel hidden activation stays in tensor memory.  Tensor-core (TF32) tier only: returns
 * MSFNO_ERR_UNSUPPORTED for shapes outside Chid % 32 == 0 (Chid <= 256, or a multiple of 256 up to 1024: processed
 * in 256-channel chunks), Cout <= 256, HW % 4 == 0 (callers then
 * use two msfno_conv1x1_fwd calls).  w1: [Chid][ldw1] (per-sample stride w1_bstride, 0 = shared), w1b: [Chid][ldw1b],
 * w2: [Cout][ldw2], b1: [Chid] (stride b1_bstride), b2: [Cout] (per-sample stride b2_bstride, 0 = shared) or NULL,
 * add: [B or 1][Cout][HW] or NULL.
 * stats: NULL, or [B][Cout][2] doubles that receive the plane (sum, sum of squares) of y -- the InstanceNorm statistics
 * the next block needs (msfno_plane_stats of y) without another pass over y.  flags bit 1: round y to TF32. */
int msfno_mlp1x1_fwd(const float* x, long x_bstride, int Cin, const float* w1, long ldw1, long w1_bstride,
                     const float* x2, long x2_bstride, int Cin2, const float* w1b, long ldw1b, const float* b1,
                     long b1_bstride, int Chid, const float* w2, long ldw2, const float* b2, long b2_bstride, const float* add,
                     long add_bstride, float* y, double* stats, int B, int Cout, long HW, int flags, void* stream);

/* ---- generic K-major batched GEMM used by the Legendre and MLP stages ---------------------
 * D[M][N] = A[M][K] * B[N][K]^T (row-major D, ldd), optional ReLU on even columns.
 * Exposed for tests and for the 1x1-conv MLPs either side of the path. */
int msfno_gemm_nt(const float* A, long lda, const float* Bm, long ldb, float* D, long ldd, int M, int N,
                  int K, int relu_even_cols, int precision, void* stream);

/* General form (the GEMMs of the adjoints; exposed for the tests): D[M][N] (+)= mask(relu(opA(A) opB(B)^T)).
 * a_kmajor / b_kmajor: 1 = element (r, k) at r*ld + k, 0 = at k*ld + r.  mask (optional, indexed like D with ldmask):
 * even columns of D are zeroed where mask <= 0 (ReLU backward on the real parts).  engine: 0 = CUDA-core FFMA,
 * 1 = plain TF32 MMA, 3 = 3xTF32 (fp32 grade); tensor-core engines fall back to FFMA for unaligned operands. */
int msfno_gemm_ex(const float* A, long lda, int a_kmajor, const float* Bm, long ldb, int b_kmajor, float* D, long ldd,
                  int M, int N, int K, int relu_even_cols, const float* mask, long ldmask, int accumulate, int engine,
                  void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MSFNO_B200_H */
