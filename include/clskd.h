/*
 * clskd.h — C ABI of libclskd_sm100.so: the B200 (sm_100a) kernels behind the
 * DCCRN teacher->student distillation step of KhanhNguyen4999/Speech-Enhancement-CLSKD.
 *
 * The reference has no FFI (it is pure PyTorch); every entry point below replaces
 * one or more torch library call sites of the reference, cited as file:line
 * relative to the reference tree.  A maintainer binds them with ctypes
 * (see INTEGRATION.md); there are no torch types in any signature.
 *
 * Conventions
 *  - every pointer is a DEVICE pointer unless the name ends in _host;
 *  - the caller allocates all outputs and workspaces; kernels never allocate or sync;
 *  - all work is enqueued on `stream` (a cudaStream_t passed as void*);
 *  - return 0 on success, <0 on error (CLSKD_ERR_*); clskd_last_error() gives the
 *    thread-local message;
 *  - activations are "channels-last" [B, T, F, C] (C fastest) described by explicit
 *    element strides, so logical NCHW views of the reference ([B, C, F, T]) need no copy;
 *  - dtype tags: CLSKD_F32 / CLSKD_BF16.  Accumulation is always fp32 (fp64 for
 *    BatchNorm / loss reductions).
 */
#ifndef CLSKD_H_
#define CLSKD_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CLSKD_F32 0
#define CLSKD_BF16 1

#define CLSKD_OK 0
#define CLSKD_ERR_ARG (-1)
#define CLSKD_ERR_CUDA (-2)
#define CLSKD_ERR_UNSUPPORTED (-3)

#define CLSKD_MAX_TAPS 16

const char* clskd_last_error(void);
/* ABI version of this header (bumped on any signature change). */
int clskd_abi_version(void);
/* 1 if the library was compiled with the tcgen05/TMA kernels (always 1 for sm_100a builds). */
int clskd_has_tcgen05(void);

/* ------------------------------------------------------------------------------------------
 * Tap-list implicit GEMM ("tapconv"): the one contraction that every convolution-like
 * operator of the path lowers to.
 *
 *   Y[b,to,fo,n] = bias[n] + sum_j sum_c X[b, to+dt[j], fo*sf+df[j], c] * W[j][c][n]
 *
 * with X read as zero outside [0,Ti) x [0,Fi).  X may be the channel-concatenation of two
 * sources (skip connections: replaces complex_cat, tools_for_model.py:181-190).
 * Replaces: F.conv1d (tools_for_model.py:57), F.conv_transpose1d (:100), the four nn.Conv2d of
 * ComplexConv2d (:252-256), the four nn.ConvTranspose2d of ComplexConvTranspose2d (:320-324),
 * nn.Linear (:171-172), the LSTM input projections (:164-167), ABF's 1x1/3x3 convs
 * (framework.py:180,184,189), and all of their data gradients.
 * ------------------------------------------------------------------------------------------ */
typedef struct ClskdTapConv {
  const void* x0;          /* source 0, element (b,t,f,c) at x0[b*x0_sB + t*x0_sT + f*x0_sF + c] */
  const void* x1;          /* source 1 or NULL */
  int64_t x0_sB, x0_sT, x0_sF;
  int64_t x1_sB, x1_sT, x1_sF;
  int32_t c0, c1;          /* channels of source 0 / 1 (c1 = 0 when x1 is NULL) */
  int32_t B, To, Fo;       /* output grid; rows M = B*To*Fo */
  int32_t Ti, Fi;          /* valid input extents */
  int32_t sf;              /* input-f step per output f */
  int32_t ntaps;
  int32_t dt[CLSKD_MAX_TAPS];
  int32_t df[CLSKD_MAX_TAPS];
  const void* w;           /* fwd: [ntaps][c0+c1][N] fp32 (CUDA-core path) or [ntaps][N][c0+c1] bf16 (tcgen05 path) */
  const float* bias;       /* [N] fp32 or NULL */
  int32_t N;
  void* y;                 /* element (b,to,fo,n) at y[b*y_sB + to*y_sT + fo*y_sF + n] */
  int64_t y_sB, y_sT, y_sF;
  int32_t x_dtype, y_dtype;
  int32_t accumulate;      /* 1: Y += result (fp32 outputs only) */
  /* Fused epilogue of the tcgen05 forward kernel (all NULL = plain contraction; the other entry points
   * reject a descriptor that sets any of them).  Applied per output element v = acc + bias[n]:
   *   ep_scale/ep_shift [N] fp32: v = v*scale[n] + shift[n]   (eval-mode nn.BatchNorm2d folded, DCCRN.py:80-81)
   *   ep_slope (device scalar):   v = v > 0 ? v : slope*v      (nn.PReLU, DCCRN.py:82)
   *   stats_sum/stats_sumsq [N] fp64: += column sums of the STORED (rounded) outputs and of their squares
   *   over the valid rows - the batch statistics of the train-mode BatchNorm that follows the conv;
   *   the caller zeroes them; several launches (sub-pixel phases of a transposed conv) accumulate. */
  const float* ep_scale;
  const float* ep_shift;
  const float* ep_slope;
  double* stats_sum;
  double* stats_sumsq;
} ClskdTapConv;

/* fp32 CUDA-core implicit GEMM (exact-fp32 policy, and layers too small for a UMMA tile) */
int clskd_tapconv_fwd(const ClskdTapConv* d, void* stream);
/* bf16 tcgen05/TMA implicit GEMM; requires c0%8==0, c1%8==0, N%8==0, bf16 x/w, dense [B,Ti,Fi,C] sources
 * (see clskd_tapconv_umma_supported).  Packed weight `w`: bf16 [ntaps][Np][c0p + c1p] with every extent rounded up to
 * a multiple of 16 (c0p = ceil16(c0) ...) and zeros in the padding: channel counts that are multiples of 8 but not of
 * 16 (the reference's quarter-width student) run with 16-wide TMA boxes over the 8 channels that exist. */
int clskd_tapconv_fwd_umma(const ClskdTapConv* d, void* stream);
int clskd_tapconv_umma_supported(const ClskdTapConv* d);
/* padded channel extent of source 1 in the packed weight of clskd_tapconv_fwd_umma: ceil16(c1), and - when source 1 is
 * narrower than the K chunk source 0 runs with (64 / 32 / 16: the largest that divides ceil16(c0)) - that chunk: the
 * narrow source is one zero-filled TMA box, the contraction keeps source 0's chunk size */
int clskd_tapconv_umma_c1p(int c0, int c1);
/* Tuning overrides of the tcgen05 forward kernel for A/B measurements (tools/kbench.py); value 0 = automatic.
 *   key 0: operand-reuse mode (1 one TMA box per tap, 2 time-grouped patches, 3 full halo patch where possible)
 *   key 1: 1 = never keep the packed weight resident in shared memory
 *   key 2: 1 = one CTA per SM
 *   key 3: 1 = route clskd_tapconv_fwd_umma to the round-1 kernel (clskd_tapconv_fwd_umma_v1)
 *   key 4: 1 = one TMA box per patch, 2 = one TMA box per time row of the patch
 *   key 5: 1 = disable the per-shape autotuner.  Without a forced configuration the first call for a shape
 *          signature (>= 65536 rows) times a handful of candidate configurations on the caller's tensors
 *          (device-wide synchronisation, once per shape and process) and caches the fastest; also disabled by
 *          the environment variable CLSKD_AUTOTUNE=0.
 *   key 6: weight-gradient kernel (clskd_tapconv_wgrad_umma): 1 = one TMA box per tap, 2 = time-grouped patches
 *          at most (no full halo patch), 3 = patches wherever the geometry allows (automatic: only for N >= 128),
 *          4 = one box per tap inside the patch kernel instead of the round-1 kernel (tests), 5 = never the
 *          tap-stacked kernel (clskd_tapconv_wgrad_umma_stacked)
 *   key 7: 1 = LSTM recurrence on the CUDA-core kernels only (the bf16 policy otherwise runs H = 32/64/128 on
 *          mma.sync tensor-core kernels with W_hh held in registers) */
int clskd_set_tuning(int key, int value);
/* the round-1 forward kernel (one TMA box per tap, weights through the ring): A/B baseline only */
int clskd_tapconv_fwd_umma_v1(const ClskdTapConv* d, void* stream);
/* diagnostics: shapes tuned so far / how many of them kept the round-1 kernel */
int clskd_tuning_stats(int* n_shapes, int* n_v1);

/* weight gradient of the same contraction:
 *   dW[j][c][n] = sum_{b,to,fo} X[b,to+dt[j],fo*sf+df[j],c] * dY[b,to,fo,n]     (fp32 out)
 * `d->y` is dY (read), `d->w` is dW (written, fp32 [ntaps][c0+c1][N]).  dW is zeroed by the call
 * unless d->accumulate. */
int clskd_tapconv_wgrad(const ClskdTapConv* d, void* stream);
/* the same weight gradient on the tcgen05 tensor cores (bf16 x and dY, fp32 dW, split-K over CTAs
 * with fp32 reductions): requires c0%8==0, c1%8==0, N%8==0 (multiples of 8 run padded to 16), power-of-two Fo, 16-byte aligned
 * tensors/strides (see clskd_tapconv_wgrad_umma_supported). */
int clskd_tapconv_wgrad_umma(const ClskdTapConv* d, void* stream);
int clskd_tapconv_wgrad_umma_supported(const ClskdTapConv* d);
/* stride-1 same-size convolutions with N <= 64 and a full (time x frequency) tap grid: the frequency taps are stacked
 * along the MMA's N dimension (one dY patch per time tap, sub-blocks one patch row apart), the X tile is read once per
 * time tap instead of once per tap.  clskd_tapconv_wgrad_umma routes eligible launches here (tuning key 6 = 5: never). */
int clskd_tapconv_wgrad_umma_stacked(const ClskdTapConv* d, void* stream);
int clskd_tapconv_wgrad_umma_stacked_supported(const ClskdTapConv* d);
/* the round-1 weight-gradient kernel (one TMA box per tap); clskd_tapconv_wgrad_umma routes its narrow-N launches here */
int clskd_tapconv_wgrad_umma_v1(const ClskdTapConv* d, void* stream);

/* ------------------------------------------------------------------------------------------
 * Layout / packing helpers
 * ------------------------------------------------------------------------------------------ */
/* dst[i0..i3 @ dst_strides] <- src[i0..i3 @ src_strides] over `shape` (4 dims, element strides,
 * i3 fastest); dtype conversion allowed.  The one layout adapter of the path. */
int clskd_strided_copy4d(const void* src, int src_dtype, const int64_t* src_strides, void* dst,
                         int dst_dtype, const int64_t* dst_strides, const int64_t* shape,
                         void* stream);
/* out[i] = sum over the entry pair (table[2i], table[2i+1]) of (+/-) src_sel[off], with entry
 * e = off*4 + neg*2 + sel (sel 0: a, sel 1: b) and e<0 -> skipped.
 * Builds the 2x-wide block weight [[Wr,-Wi],[Wi,Wr]] of a complex convolution
 * (tools_for_model.py:252-259, :320-327) in whatever order a kernel wants. */
int clskd_pack_gather(const float* a, const float* b, const int32_t* table, int64_t n, void* out,
                      int out_dtype, void* stream);
/* clskd_pack_gather for many weights in one launch (after the optimizer step rewrote the parameters, every packed
 * kernel-side weight of the step is rebuilt at once instead of lazily, one launch each).  desc: device array of
 * n_entries x 6 int64 {a, b (or a), table, out, start, n*2 + (out_dtype == CLSKD_BF16)}; `start` = offset of the entry
 * in the concatenated index space, ascending from 0; total = sum of n. */
int clskd_multi_pack_gather(const int64_t* desc, int n_entries, int64_t total, void* stream);
/* dst[j] = s0*src[i0] + s1*src[i1], table entry pair (e0,e1) with e = idx*2 + neg, e<0 -> skipped.
 * Folds the block-weight gradient back onto the reference's separate real/imag parameters. */
int clskd_unpack_gather2(const float* src, const int32_t* table2, int64_t n, float* dst,
                         int accumulate, void* stream);
/* Split-bf16 operand staging for fp32-accurate GEMMs on the tensor cores (replaces the TF32/fp32
 * library GEMMs behind F.conv1d / F.conv_transpose1d of the STFT front end, tools_for_model.py:57,100,
 * torch.stft of framework.py:27, and the fp32 nn.LSTM / nn.Linear projections, tools_for_model.py:164-172).
 * Row m = (i0*n1 + i1)*n2 + i2 of the fp32 source is x[i0*s0 + i1*s1 + i2*s2 + k*sk], k < K.
 * hi = bf16(x), lo = bf16(x - hi).  out is bf16 [Mp, nseg*Kp] (rows >= n0*n1*n2 and k >= K zero):
 *   order 0 (activation operand): [hi | lo] (nseg 2) or [hi | lo | hi] (nseg 3)
 *   order 1 (weight operand, nseg 3): [hi | hi | lo]
 * so that one bf16 contraction over 3*Kp yields x_hi*w_hi + x_lo*w_hi + x_hi*w_lo. */
int clskd_split_bf16x3(const float* x, int64_t s0, int64_t s1, int64_t s2, int64_t sk, int64_t n0,
                       int n1, int n2, int K, int Kp, int64_t Mp, int order, int nseg, void* out,
                       void* stream);
/* zero / reflect padding of waveforms: dst[b, i] = src[b, map(i - left)], i in [0, L+left+right).
 * mode 0: zeros (ConvSTFT, tools_for_model.py:56); mode 1: reflect (torch.stft center=True,
 * framework.py:27). dst is fp32. */
int clskd_pad1d(const void* src, int src_dtype, int64_t src_sB, int B, int L, int left, int right,
                int mode, float* dst, void* stream);
/* gradient of clskd_pad1d: dsrc[b,i] = sum of ddst entries that read it (fp32). */
int clskd_pad1d_bwd(const float* ddst, int B, int L, int left, int right, int mode, float* dsrc,
                    int accumulate, void* stream);

/* ------------------------------------------------------------------------------------------
 * BatchNorm2d (+PReLU) over channels-last rows.  Replaces nn.BatchNorm2d + nn.PReLU
 * (DCCRN.py:80-82,123-126) and the BatchNorm2d inside ABF (framework.py:181,185).
 * ------------------------------------------------------------------------------------------ */
/* per-channel sums over M rows of a dense [M, C] matrix: sum[c] = sum x, sumsq[c] = sum x*x (fp64);
 * outputs are zeroed by the call. */
int clskd_colstats(const void* x, int dtype, int64_t M, int C, double* sum, double* sumsq,
                   void* stream);
/* finalize training statistics: mean/invstd (fp32 [C]) from the fp64 sums; updates running stats
 * with `momentum` using the unbiased variance exactly like torch (running_* may be NULL). */
int clskd_bn_finalize(const double* sum, const double* sumsq, int64_t M, int C, float eps,
                      float momentum, float* mean, float* invstd, float* running_mean,
                      float* running_var, void* stream);
/* eval statistics: mean = running_mean, invstd = rsqrt(running_var + eps) */
int clskd_bn_eval_stats(const float* running_mean, const float* running_var, int C, float eps,
                        float* mean, float* invstd, void* stream);
/* eval-mode BatchNorm folded to one affine per channel for a conv epilogue:
 * scale = gamma*rsqrt(running_var+eps), shift = beta - running_mean*scale (gamma/beta may be NULL) */
int clskd_bn_fold(const float* running_mean, const float* running_var, const float* gamma,
                  const float* beta, int C, float eps, float* scale, float* shift, void* stream);
/* y = prelu((x-mean)*invstd*gamma+beta); slope is a DEVICE scalar pointer or NULL (identity). */
int clskd_bn_act_fwd(const void* x, int x_dtype, int64_t M, int C, const float* mean,
                     const float* invstd, const float* gamma, const float* beta,
                     const float* slope, void* y, int y_dtype, void* stream);
/* backward pass 1: per-channel sums of dz and dz*xhat where dz = dy * prelu'(bn(x)); also
 * dslope = sum dy*min(0,bn(x)).  sums are fp64 [C], dslope fp64 scalar (NULL ok); zeroed by the call. */
int clskd_bn_act_bwd_stats(const void* x, int x_dtype, const void* dy, int dy_dtype, int64_t M,
                           int C, const float* mean, const float* invstd, const float* gamma,
                           const float* beta, const float* slope, double* sum_dz,
                           double* sum_dz_xhat, double* dslope, void* stream);
/* backward pass 2: dx = gamma*invstd*(dz - mean(dz) - xhat*mean(dz*xhat)) (training) or
 * gamma*invstd*dz (eval: training=0); also writes dgamma/dbeta (fp32 [C], may be NULL) and
 * dslope_out (fp32 scalar, may be NULL). */
int clskd_bn_act_bwd_apply(const void* x, int x_dtype, const void* dy, int dy_dtype, int64_t M,
                           int C, const float* mean, const float* invstd, const float* gamma,
                           const float* beta, const float* slope, const double* sum_dz,
                           const double* sum_dz_xhat, const double* dslope, int training, void* dx,
                           int dx_dtype, float* dgamma, float* dbeta, float* dslope_out,
                           void* stream);

/* ------------------------------------------------------------------------------------------
 * Complex BatchNorm (Trabelsi 2x2 whitening), tools_for_model.py:398-508.  x is [M, 2*Cc]
 * (real half then imag half of the channel axis).
 * ------------------------------------------------------------------------------------------ */
/* moments: s[0..4][Cc] = sum xr, xi, xr*xr, xr*xi, xi*xi (fp64, zeroed by the call) */
int clskd_cbn_moments(const void* x, int dtype, int64_t M, int Cc, double* s, void* stream);
/* training=1: stats from moments + running update (lerp, biased moments like the reference);
 * training=0: stats from running buffers.  coef[7][Cc] = Mr, Mi, Zrr, Zri, Zir, Zii, (unused);
 * also saves whitening terms needed by backward in `saved` [8][Cc]. */
int clskd_cbn_finalize(const double* s, int64_t M, int Cc, float eps, float momentum, int training,
                       const float* Wrr, const float* Wri, const float* Wii, float* RMr, float* RMi,
                       float* RVrr, float* RVri, float* RVii, float* coef, void* stream);
int clskd_cbn_apply(const void* x, int x_dtype, int64_t M, int Cc, const float* coef,
                    const float* Br, const float* Bi, void* y, int y_dtype, void* stream);
/* Backward of the block (autograd of tools_for_model.py:398-508; with batch statistics the gradient
 * flows through the mean and the 2x2 covariance, like the reference's undetached statistics):
 *   pass 1  s6[0..5][Cc] = sum dyr, dyi, dyr*xr, dyr*xi, dyi*xr, dyi*xi        (fp64, zeroed by the call)
 *   pass 2  per channel: dWrr/dWri/dWii/dBr/dBi (fp32 [Cc], each may be NULL) and coefb[11][Cc] such that
 *           dx = Z^T (dy - mean dy) + G (x - mean x)   (G = 0 with running statistics, training=0);
 *           `s` are the forward moments of clskd_cbn_moments (training=1), else the running buffers are used
 *   pass 3  dx from x, dy and coefb (x, dy, dx share one dtype) */
int clskd_cbn_bwd_moments(const void* x, const void* dy, int dtype, int64_t M, int Cc, double* s6,
                          void* stream);
int clskd_cbn_bwd_finalize(const double* s, const double* s6, int64_t M, int Cc, float eps, int training,
                           const float* Wrr, const float* Wri, const float* Wii, const float* RMr,
                           const float* RMi, const float* RVrr, const float* RVri, const float* RVii,
                           float* coefb, float* dWrr, float* dWri, float* dWii, float* dBr, float* dBi,
                           void* stream);
int clskd_cbn_bwd_apply(const void* x, const void* dy, int dtype, int64_t M, int Cc, const float* coefb,
                        void* dx, void* stream);

/* ------------------------------------------------------------------------------------------
 * Mask / spectrum kernels (DCCRN.py:153,159,207-232)
 * spec / out_spec are interleaved complex spectra [B, T, 257, 2]; mask is the last decoder
 * output [B, T(+offset), 256, 2] addressed by strides (DC bin mask is zero, DCCRN.py:209-210).
 * mode: 0 = 'E', 1 = 'C', 2 = 'R'.
 * ------------------------------------------------------------------------------------------ */
int clskd_mask_fwd(const float* spec, const void* mask, int mask_dtype, int64_t m_sB, int64_t m_sT,
                   int B, int T, int nbins, int mode, float* out_spec, float* mask_padded,
                   void* stream);
/* gradient wrt the mask (spec is data): dmask has the mask's strides/dtype */
int clskd_mask_bwd(const float* spec, const void* mask, int mask_dtype, int64_t m_sB, int64_t m_sT,
                   int B, int T, int nbins, int mode, const float* dout_spec, void* dmask,
                   int dmask_dtype, int64_t dm_sB, int64_t dm_sT, void* stream);

/* ------------------------------------------------------------------------------------------
 * iSTFT overlap-add (tools_for_model.py:100-107) + squeeze/clamp (DCCRN.py:235-237)
 * frames [B, T, win] fp32 (synthesis-basis GEMM output) -> wav [B, L], L = (T-1)*hop+win-2*trim
 * wav[b,i] = clamp( sum_t frames[b,t,i+trim-t*hop] / (coff[i+trim]+1e-8), -1, 1 )
 * coff = overlap-added window^2 (computed in-kernel from `window`; window NULL: no normalisation,
 * which is the plain conv_transpose1d used as the data gradient of the framing GEMM).
 * ------------------------------------------------------------------------------------------ */
int clskd_ola_fwd(const float* frames, const float* window, int B, int T, int win, int hop,
                  int trim, int do_clamp, float* wav, void* stream);
/* dframes[b,t,k] = dwav[b, t*hop+k-trim] / (coff+1e-8) * (|wav| < 1 or !do_clamp), 0 outside */
int clskd_ola_bwd(const float* dwav, const float* wav, const float* window, int B, int T, int win,
                  int hop, int trim, int do_clamp, float* dframes, void* stream);

/* ------------------------------------------------------------------------------------------
 * Time-domain objectives (tools_for_loss.py:30-47,83-97), MSE (DCCRN.py:260-261)
 * kind: 0 = si_snr(s1,s2), 1 = sdr(s1,s2), 2 = si_sdr(reference=s1, estimation=s2), 3 = mse.
 * out is one fp32 scalar; `part` is a [B,4] fp64 workspace that bwd re-uses.
 * ------------------------------------------------------------------------------------------ */
int clskd_wave_loss_fwd(const float* s1, const float* s2, int B, int L, int kind, float eps,
                        double* part, float* out, void* stream);
/* ds1 (and/or ds2, either may be NULL) = gout * d out / d s */
int clskd_wave_loss_bwd(const float* s1, const float* s2, int B, int L, int kind, float eps,
                        const double* part, const float* gout, float* ds1, float* ds2,
                        void* stream);

/* ------------------------------------------------------------------------------------------
 * STFT-magnitude losses (framework.py:16-68): inputs are the interleaved spectra
 * [M, nbins, 2] of prediction x and target y.
 *   part[0] = sum |log ymag - log xmag|, part[1] = sum (ymag-xmag)^2, part[2] = sum ymag^2
 *   mag = sqrt(clamp(re^2+im^2, 1e-7));  n = number of (row, bin) pairs
 *   out[0] = sqrt(part[1])/sqrt(part[2]) (spectral convergence), out[1] = part[0]/n (log-mag L1)
 * ------------------------------------------------------------------------------------------ */
int clskd_stftmag_loss_fwd(const float* xs, const float* ys, int64_t n, double* part, float* out,
                           void* stream);
/* dxs = g_mag * d(mean|log y - log x|)/dxs + g_sc * d(||y-x||_F/||y||_F)/dxs; g_* are host floats
 * already containing the upstream gradient and the mean/factor scaling. */
int clskd_stftmag_loss_bwd(const float* xs, const float* ys, int64_t n, const double* part,
                           const float* gout_mag, const float* gout_sc, float scale_mag,
                           float scale_sc, float* dxs, void* stream);

/* ------------------------------------------------------------------------------------------
 * SPKD (framework.py:150-172): Gram G = Z Z^T over the flattened feature axis, L1 row
 * normalisation, squared Frobenius distance, /B^2.
 * ------------------------------------------------------------------------------------------ */
/* G[B,B] (fp32) = Z Z^T (+= when accumulate); Z is [B, K] with row stride ldz (elements). */
int clskd_gram_fwd(const void* z, int dtype, int B, int64_t K, int64_t ldz, float* G,
                   int accumulate, void* stream);
/* the same Gram product / gradient on the tcgen05 tensor cores (bf16 z, B <= 512 - batches above 128 rows run as
 * 128-row blocks: cross blocks read both row blocks and write the mirrored result too - 16-byte aligned
 * rows, K >= 4096): each feature element is read from HBM once, partial Grams stay in TMEM.
 * clskd_gram_bwd_umma always overwrites dz. */
int clskd_gram_umma_supported(const void* z, int dtype, int B, int64_t K, int64_t ldz);
int clskd_gram_fwd_umma(const void* z, int dtype, int B, int64_t K, int64_t ldz, float* G,
                        int accumulate, void* stream);
int clskd_gram_bwd_umma(const void* z, int dtype, int B, int64_t K, int64_t ldz, const float* dG,
                        const float* gout, void* dz, int dz_dtype, int64_t lddz, void* stream);
/* loss (fp32 scalar) = || rownorm1(Gt) - rownorm1(Gs) ||_F^2 * scale; also dGs = dloss/dGs (fp32
 * [B,B], may be NULL) for a unit upstream gradient. */
int clskd_spkd_loss(const float* Gt, const float* Gs, int B, float scale, float* loss, float* dGs,
                    void* stream);
/* dZ[i,:] = gout * sum_j (dG[i,j] + dG[j,i]) Z[j,:] */
int clskd_gram_bwd(const void* z, int dtype, int B, int64_t K, int64_t ldz, const float* dG,
                   const float* gout, void* dz, int dz_dtype, int64_t lddz, int accumulate,
                   void* stream);

/* ------------------------------------------------------------------------------------------
 * LSTM recurrence (nn.LSTM inside NavieComplexLSTM, tools_for_model.py:144-147,164-167).
 * `pre` holds the input projections + both biases (gate order i,f,g,o).  R = P*Bp rows share one
 * weight set (P parts, e.g. {real,imag} inputs, of Bp batch rows each); the pre-activation of
 * (set s, time t, row r=(p,b), gate g) is at
 *     pre[s*pre_set_stride + p*pre_pstride + t*pre_tstride + b*pre_ld + g].
 * `whh_t` is W_hh transposed, [nsets][H][4H] fp32 (w_bf16=1 keeps it as bf16 in shared memory:
 * the bf16 policy for H up to 128).  Outputs are dense: h [nsets][P][T][Bp][H]; for training also
 * gates [nsets][P][T][Bp][4H] (post-activation) and c [nsets][P][T][Bp][H] (NULL to skip).
 * ------------------------------------------------------------------------------------------ */
int clskd_lstm_fwd(const float* pre, const float* whh_t, int T, int R, int Bp, int H, int nsets,
                   int64_t pre_pstride, int64_t pre_tstride, int64_t pre_ld, int64_t pre_set_stride,
                   int64_t whh_set_stride, int w_bf16, float* h, float* gates, float* c,
                   void* stream);
/* The same recurrence with a carried state (time-chunked streaming inference, eval.py:42-60 on long
 * utterances): h0 / c0 [nsets][R][H] initialise the state (NULL = zeros, the reference's default), hN / cN
 * receive the state after the last step (NULL to skip; may alias h0 / c0 for an in-place carry).
 * With T == 0 nothing is launched and hN / cN are left untouched. */
int clskd_lstm_fwd_state(const float* pre, const float* whh_t, int T, int R, int Bp, int H, int nsets,
                         int64_t pre_pstride, int64_t pre_tstride, int64_t pre_ld,
                         int64_t pre_set_stride, int64_t whh_set_stride, int w_bf16, float* h,
                         float* gates, float* c, const float* h0, const float* c0, float* hN,
                         float* cN, void* stream);
/* BPTT: given dh_out (gradient wrt every h_t, layout of h) produces dpre (gradient wrt the
 * pre-activations, addressed like `pre`).  gates/c are the saved forward tensors; `whh` is
 * W_hh [nsets][4H][H] fp32. */
int clskd_lstm_bwd(const float* dh_out, const float* whh, const float* gates, const float* c, int T,
                   int R, int Bp, int H, int nsets, int64_t whh_set_stride, int64_t pre_pstride,
                   int64_t pre_tstride, int64_t pre_ld, int64_t pre_set_stride, float* dpre,
                   void* stream);
/* Same, with the recurrence precision of the forward pass: w_bf16 != 0 (the bf16 policy) contracts bf16-rounded
 * W_hh with the gate gradients as a bf16 hi + lo pair on mma.sync tensor cores (H = 32 / 64 / 128; other sizes and
 * w_bf16 == 0 run the fp32 CUDA-core kernel of clskd_lstm_bwd). */
int clskd_lstm_bwd_policy(const float* dh_out, const float* whh, const float* gates, const float* c, int T,
                          int R, int Bp, int H, int nsets, int64_t whh_set_stride, int64_t pre_pstride,
                          int64_t pre_tstride, int64_t pre_ld, int64_t pre_set_stride, float* dpre,
                          int w_bf16, void* stream);

/* ------------------------------------------------------------------------------------------
 * ABF helpers (framework.py:206-224)
 * ------------------------------------------------------------------------------------------ */
/* nearest resize along F of a dense [B,T,Fi,C] tensor to [B,T,Fo,C] (F.interpolate nearest) */
int clskd_resize_f_fwd(const void* x, int dtype, int64_t BT, int Fi, int Fo, int C, void* y,
                       void* stream);
int clskd_resize_f_bwd(const void* dy, int dtype, int64_t BT, int Fi, int Fo, int C, void* dx,
                       void* stream);
/* out = x*sigmoid(z0) + y*sigmoid(z1); z is [M,2] logits (fp32) */
int clskd_att_blend_fwd(const void* x, const void* y, int dtype, const float* z, int64_t M, int C,
                        void* out, void* stream);
int clskd_att_blend_bwd(const void* x, const void* y, int dtype, const float* z, const void* dout,
                        int64_t M, int C, void* dx, void* dy, float* dz, void* stream);

/* Fused middle stage of an ABF block (framework.py:209-219) on dense channels-last tensors:
 *   xp = BN(z1) with the given per-channel mean / invstd / gamma / beta;  yv = y_prev resized
 *   (nearest) from Fy to F frequency rows, Fy == F or 2*Fy == F;  logit_k = W_att[k] . [xp ; yv] + b_att[k];
 *   xb = xp * sigmoid(logit_0) + yv * sigmoid(logit_1).
 * z1, xb, dz1, gout: [B,T,F,C]; y_prev, dy: [B,T,Fy,C]; logits: fp32 [B,T,F,2] (saved for backward);
 * watt: fp32 [2][2C] (nn.Conv2d(2C, 2, 1) weight, x channels first); batt: fp32 [2] or NULL.
 * clskd_abf_mid_bwd runs the statistics pass and the apply pass: sums[2][C] = (sum dxp, sum dxp*xhat)
 * (= dbeta, dgamma), dwatt[2][2C], dbatt[2] (all fp64, zeroed by the call), dz1 (BatchNorm backward
 * with batch statistics when training != 0) and dy (adjoint of the resize). */
int clskd_abf_mid_supported(int B, int T, int F, int Fy, int C);
int clskd_abf_mid_fwd(const void* z1, const void* y, int dtype, int B, int T, int F, int Fy, int C,
                      const float* mean, const float* invstd, const float* gamma, const float* beta,
                      const float* watt, const float* batt, void* xb, float* logits, void* stream);
int clskd_abf_mid_bwd(const void* gout, const void* z1, const void* y, int dtype, int B, int T, int F,
                      int Fy, int C, const float* mean, const float* invstd, const float* gamma,
                      const float* beta, const float* watt, const float* logits, int training,
                      double* sums, double* dwatt, double* dbatt, void* dz1, void* dy, void* stream);

/* One-pass backward of the middle stage with the BatchNorm backward of conv1 (framework.py:209, nn.BatchNorm2d after
 * the 1x1 nn.Conv2d(in, mid, bias=False)) FOLDED into conv1's own gradients - the batch-mean terms of
 *   dz1 = gi (dxp - mean dxp - xhat mean(dxp xhat)),   gi = gamma invstd,   xhat = (W1 x - mu) invstd
 * are linear in x, so dz1 is never materialised and gout / z1 / y_prev are read once instead of twice:
 *   clskd_abf_mid_bwd_fold: sums / dwatt / dbatt as clskd_abf_mid_bwd, dy, and dxp [B,T,F,C] = the gradient of the
 *     BatchNorm OUTPUT (row-local).
 *   clskd_abf_fold_dgrad: packed weight weff bf16 [Cin][C + c1p] and bias fp32 [Cin] of the two-source 1x1 contraction
 *     dx = [dxp | x] weff^T + bias (clskd_tapconv_fwd_umma with x0 = dxp, x1 = x, c1p = clskd_tapconv_umma_c1p(C, Cin));
 *     w1: fp32 [C][Cin]; sums from the call above; M = B*T*F rows; training = 0: running statistics (no mean terms).
 *   clskd_abf_fold_dw1: dW1 fp32 [C][Cin] from P = x^T dxp (fp32 [Cin][C], clskd_tapconv_wgrad* on dxp),
 *     G = x^T x (fp64 [Cin][Cin]) and sx = column sums of x (fp64 [Cin]) - both from clskd_colgram;
 *     G and sx may be NULL when training = 0.
 *   clskd_colgram: G and sx of a dense bf16 map x [M][C], C in {16, 32, 64, 128}, in one pass over x (mma.sync on
 *     ldmatrix.trans tiles; outputs zeroed by the call).
 *   clskd_abf_fold_stats: column sums / sums of squares (fp64 [C] each) of z1 = bf16(W1) x from G and sx - the batch
 *     statistics of conv1's BatchNorm without the statistics epilogue of the conv (or a pass over z1). */
int clskd_abf_mid_bwd_fold(const void* gout, const void* z1, const void* y, int dtype, int B, int T, int F,
                           int Fy, int C, const float* mean, const float* invstd, const float* gamma,
                           const float* beta, const float* watt, const float* logits, double* sums,
                           double* dwatt, double* dbatt, void* dxp, void* dy, void* stream);
int clskd_abf_fold_dgrad(const float* w1, const float* gamma, const float* mean, const float* invstd,
                         const double* sums, int64_t M, int training, int C, int Cin, int c1p, void* weff,
                         float* bias, void* stream);
int clskd_colgram_supported(int dtype, int C);
int clskd_colgram(const void* x, int dtype, int64_t M, int C, double* G, double* sx, void* stream);
int clskd_abf_fold_stats(const double* G, const double* sx, const float* w1, int C, int Cin, double* sum,
                         double* sumsq, void* stream);
int clskd_abf_fold_dw1(const float* P, const double* G, const double* sx, const float* w1, const float* gamma,
                       const float* mean, const float* invstd, const double* sums, int64_t M, int training,
                       int C, int Cin, float* dw1, void* stream);

/* The same block when the 1x1 conv in front of it has only TWO input channels (the mask-level map
 * of the decoder side, framework.py:209 with in_channel = 2): z1 = W1 x is recomputed per row from
 * x [B,T,F,2] and w1 [C][2] (fp32) instead of being stored and re-read, BatchNorm statistics of z1 come
 * from the 2x2 moments of x (clskd_cbn_moments with Cc = 1 -> clskd_rank2_colstats -> clskd_bn_finalize),
 * and the backward returns dx [B,T,F,2] and dw1 (fp64 [C][2], zeroed by the call) directly - replacing the
 * conv1 forward, its statistics pass, its data- and weight-gradient launches and the dz1 round trip. */
int clskd_abf_mid_xs_fwd(const void* x, const float* w1, const void* y, int dtype, int B, int T, int F,
                         int Fy, int C, const float* mean, const float* invstd, const float* gamma,
                         const float* beta, const float* watt, const float* batt, void* xb, float* logits,
                         void* stream);
int clskd_abf_mid_xs_bwd(const void* gout, const void* x, const float* w1, const void* y, int dtype, int B,
                         int T, int F, int Fy, int C, const float* mean, const float* invstd,
                         const float* gamma, const float* beta, const float* watt, const float* logits,
                         int training, double* sums, double* dwatt, double* dbatt, double* dw1, void* dx,
                         void* dy, void* stream);
/* The same 2-channel middle stage with the rank-2 structure of z1 = W1 x folded through every reduction
 * (the x half of the logits is two scalars per row; the BatchNorm-backward batch sums, dW_att and dW1 follow from
 * three per-channel sums, six scalar sums and the 2x2 moments of x; dx needs two dot products per row): the
 * backward is ONE pass over gout / y_prev + a one-block finalisation + a 24-byte-per-row pass for dx.
 * Same arguments / outputs as clskd_abf_mid_xs_fwd / _bwd; the backward takes a caller-allocated workspace of
 * clskd_abf_xs2_bwd_workspace(B, T, F, C, &bytes) bytes (16-byte aligned). */
int clskd_abf_xs2_fwd(const void* x, const float* w1, const void* y, int dtype, int B, int T, int F, int Fy,
                      int C, const float* mean, const float* invstd, const float* gamma, const float* beta,
                      const float* watt, const float* batt, void* xb, float* logits, void* stream);
int clskd_abf_xs2_bwd_workspace(int B, int T, int F, int C, int64_t* bytes);
int clskd_abf_xs2_bwd(const void* gout, const void* x, const float* w1, const void* y, int dtype, int B, int T,
                      int F, int Fy, int C, const float* mean, const float* invstd, const float* gamma,
                      const float* beta, const float* watt, const float* logits, int training, double* sums,
                      double* dwatt, double* dbatt, double* dw1, void* dx, void* dy, void* workspace,
                      int64_t ws_bytes, void* stream);
/* sum[c] = sum_m (W1 x_m)[c], sumsq[c] = sum_m (W1 x_m)[c]^2 from s5 = moments of the 2-channel x */
int clskd_rank2_colstats(const double* s5, const float* w1, int C, double* sum, double* sumsq,
                         void* stream);

/* Tap-in-channel decomposition of a convolution-like layer with very few output channels (ABF's
 * 3x3 conv onto the 2-channel mask map; the last, mask-producing transposed conv of the decoder):
 * a pointwise GEMM first produces, at every INPUT position, the contribution to each (tap, n) pair
 * as channels z[.., j*N+n]; these kernels then gather-sum the taps into the output grid
 *     y[b,t,f,n] = bias[n] + sum_j z[b, t+dt[j], (f+df[j])/sf, j*N+n]
 * (terms with (f+df[j]) not divisible by sf, or outside [0,Ti)x[0,Fi), are absent: sf = 1 is a
 * stride-1 "same" conv, sf = 2 the sub-pixel structure of a stride-2 transposed conv), and scatter
 * the gradient back  dz[b,ti,fi,j*N+n] = dy[b, ti-dt[j], fi*sf-df[j], n]  (channels >= ntaps*N of dz
 * are zero).  z/dz are dense [B,Ti,Fi,Zc], y/dy dense [B,To,Fo,N]; dt_host/df_host are HOST arrays. */
int clskd_tapsum_fwd(const void* z, int z_dtype, int B, int Ti, int Fi, int To, int Fo, int sf, int Zc,
                     int ntaps, const int32_t* dt_host, const int32_t* df_host, int N,
                     const float* bias, void* y, int y_dtype, void* stream);
int clskd_tapsum_bwd(const void* dy, int dy_dtype, int B, int Ti, int Fi, int To, int Fo, int sf, int Zc,
                     int ntaps, const int32_t* dt_host, const int32_t* df_host, int N, void* dz,
                     int dz_dtype, void* stream);

/* ------------------------------------------------------------------------------------------
 * hcl (framework.py:287-306): adaptive average pooling of a dense [B,T,F,C] map to (l,l) per
 * channel over the logical (H=F, W=T) plane -> out [B, C, l, l] fp32.
 * ------------------------------------------------------------------------------------------ */
int clskd_adaptive_pool_fwd(const void* x, int dtype, int B, int T, int F, int C, int l,
                            float* out, void* stream);
int clskd_adaptive_pool_bwd(const float* dout, int B, int T, int F, int C, int l, void* dx,
                            int dtype, int accumulate, void* stream);
/* out[0] (+)= scale * sum (a-b)^2 ; generic dtype */
int clskd_sqdiff_sum(const void* a, int a_dtype, const void* b, int b_dtype, int64_t n,
                     double* out, void* stream);
/* da (+)= g * 2*(a-b) with g = gout[0]*scale */
int clskd_sqdiff_bwd(const void* a, int a_dtype, const void* b, int b_dtype, int64_t n,
                     const float* gout, float scale, void* da, int da_dtype, int accumulate,
                     void* stream);

/* ------------------------------------------------------------------------------------------
 * Optimizer (optim.Adam, distill.py:202-204): flat fp32 buffers.
 * ------------------------------------------------------------------------------------------ */
/* hyper-parameters are doubles: the bias corrections 1 - beta^step are evaluated in double like
 * torch.optim.Adam (1 - 0.999f in fp32 is already off by 1.3e-5 relative at step 1) */
int clskd_adam_step(float* p, const float* g, float* m, float* v, int64_t n, double lr, double beta1,
                    double beta2, double eps, double weight_decay, int step, double grad_scale,
                    void* stream);

/* gradient bucket: flat[offsets[i]..offsets[i+1]) = fp32 tensor at device address ptrs[i]
 * (zeros when ptrs[i] == 0).  `ptrs` (uint64[n]) and `offsets` (int64[n+1]) are DEVICE arrays;
 * total = offsets[n].  One launch builds the flat bucket that NCCL all-reduces and Adam consumes. */
int clskd_multi_pack_f32(const void* ptrs, const int64_t* offsets, int n, int64_t total,
                         float* flat, void* stream);

/* small utilities */
int clskd_fill_f32(float* p, int64_t n, float v, void* stream);
int clskd_axpy_f32(float* y, const float* x, int64_t n, float a, void* stream);
/* out = a*x + b*y (y may be NULL): the rr-ii / ir+ri combination of the complex LSTM
 * (tools_for_model.py:168-169) and its gradient. */
int clskd_axpby_f32(const float* x, const float* y, float a, float b, float* out, int64_t n,
                    void* stream);
/* out[0] = sum_i w[i]*in[i] over n fp32 scalars given as an array of device pointers is not
 * needed: losses are combined by the autograd graph on the host side. */

/* Per-utterance second moments of two fp32 waveform batches a, b [B, L] (row strides in elements):
 * out[b*5 .. b*5+4] = (sum a, sum b, sum a*a, sum a*b, sum b*b) in fp64.  One pass gives every batched validation
 * metric of the training loop (SI-SDR with mean removal, SNR and their improvements over the mixture), which the
 * reference computes one utterance at a time on the CPU (distill.py:150-200, asteroid get_metrics). */
int clskd_pair_moments(const float* a, const float* b, int B, int L, int64_t a_sB, int64_t b_sB, double* out,
                       void* stream);

/* out = in0 + in1 (+ in2 + in3), k in 2..4 dense same-dtype tensors of n elements, fp32 accumulation: the gradient
 * accumulation of a tensor with several consumers (skip connection + next layer + feature tap), which autograd
 * would otherwise do with its own add kernels.  16-byte aligned tensors. */
int clskd_sum_n(const void* in0, const void* in1, const void* in2, const void* in3, int k, int dtype,
                int64_t n, void* out, void* stream);

/* fp64 -> fp32 scalar conversion with scaling: out[i] = (float)(in[i]*scale) */
int clskd_f64_to_f32(const double* in, int n, double scale, float* out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* CLSKD_H_ */
