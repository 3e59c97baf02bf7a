/*
 * clpk.h — C ABI of libclpk.so, the B200 (sm_100a) decode hot path of the CLIP-feature image codec.
 *
 * The reference (lionl1106/Clip-Neural-image-conpression, package clip_feature_codec) is pure Python and has
 * NO FFI/plugin layer: its boundary is a Python class surface.  This header is therefore the
 * boundary a maintainer would bind with ctypes; every entry point names the reference
 * code (file:line, relative to the reference root, PKG = src/clip_feature_codec) whose arithmetic it replaces.
 *
 * Conventions
 *   - every pointer marked "dev" is a caller-owned DEVICE pointer; "host" pointers are plain host memory;
 *   - `stream` is a cudaStream_t passed as void* (0 = legacy default stream);
 *   - return value: 0 = CLPK_OK, otherwise an error code; clpk_last_error() gives the message (thread local);
 *   - no entry point allocates device memory except clpk_plan_create / clpk_plan_prepare_ddim;
 *   - no CPU fallback exists: without a CUDA device every compute entry point returns CLPK_ERR_CUDA.
 *   - activations inside the library are NHWC; the reference-facing tensors (x_t, eps, images) are NCHW fp32.
 */
#ifndef CLPK_H_
#define CLPK_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CLPK_OK 0
#define CLPK_ERR_ARG 1     /* bad argument / unsupported shape */
#define CLPK_ERR_CUDA 2    /* CUDA runtime / driver failure    */
#define CLPK_ERR_STATE 3   /* call order (plan not prepared …) */

#define CLPK_MAX_LEVELS 8

/* 16-bit tensor-core operand formats (both run at the same tcgen05 kind::f16 rate, fp32 accumulation in TMEM).
 * F16 (10 mantissa bits) is the plan default: with BF16 (7 bits) the per-step epsilon error of narrow UNets
 * (base=32) exceeds the 1e-2 parity bar (measured 1.6e-2) — see DESIGN.md "Precision". */
#define CLPK_OP_BF16 0
#define CLPK_OP_F16 1

const char* clpk_last_error(void);
int clpk_version(void);
/* number of kernels this library has launched since load (bench.py's gpu_launches). */
uint64_t clpk_launch_count(void);

/* ---------------------------------------------------------------------------------------------------------------
 * Codec side: PKG/codecs/quantizer.py, PKG/cli/reconstruct_diffusion.py:43-44, PKG/cli/eval.py:58-59
 * ------------------------------------------------------------------------------------------------------------- */

/* z[b,d] = float(q[b,d]) * scale[d] + zero[d]   (separately rounded mul and add: numpy semantics,
 *          reconstruct_diffusion.py:43 / quantizer.py:38-39; bit exact)
 * then, if l2norm != 0:  z[b,:] /= max(||z[b,:]||_2, 1e-9)   (reconstruct_diffusion.py:21-23,44)
 * If z_raw (dev, may be NULL) is given it receives the un-normalised dequantised values. */
int clpk_dequant_l2norm_u8(const uint8_t* q_dev, const float* scale_dev, const float* zero_dev,
                           float* z_dev, float* z_raw_dev, int batch, int dim, int l2norm, void* stream);

/* q[b,d] = uint8(clamp(rint((x[b,d] - zero[d]) / scale[d]), 0, 255))  — quantizer.py:29-33 (round half even). */
int clpk_quant_encode_u8(const float* x_dev, const float* scale_dev, const float* zero_dev,
                         uint8_t* q_dev, int batch, int dim, void* stream);

/* per-channel fit: zero[d] = min_n x[n,d]; scale[d] = max(max_n x[n,d] - zero[d], 1e-8) / 255 — quantizer.py:22-27 */
int clpk_quant_fit(const float* x_dev, float* scale_dev, float* zero_dev, int n, int dim, void* stream);

/* ---------------------------------------------------------------------------------------------------------------
 * Sampler side: PKG/diffusion/ddim.py:34-45 (one DDIM update, elementwise)
 *   x0 = clamp((x - c[0]*eps) / c[1], -1, 1);  x' = c[2]*x0 + c[3]*eps  [+ c[4]*noise]
 *   c = {sqrt(1-a_t), sqrt(a_t), sqrt(a_s), sqrt(a_s - sigma^2), sigma}  (host computed with the reference's ops)
 *   noise_dev may be NULL (eta == 0 or sigma == 0).  All products/sums are individually rounded like the ATen
 *   kernels, so the result is bit-identical to the reference for identical eps.  x_out may alias x.
 * ------------------------------------------------------------------------------------------------------------- */
int clpk_ddim_step(const float* x_dev, const float* eps_dev, const float* noise_dev, const float* coef5_host,
                   float* x_out_dev, int64_t n, void* stream);

/* ---------------------------------------------------------------------------------------------------------------
 * Model leaf ops (used by the plan; exported so each one can be parity-tested alone)
 * ------------------------------------------------------------------------------------------------------------- */

/* timestep_embedding (PKG/models/unet.py:22-39): out[b, 0:half] = cos(t*f_k), out[b, half:2*half] = sin(t*f_k) */
int clpk_timestep_embedding(const int64_t* t_dev, float* out_dev, int batch, int dim, float max_period, void* stream);
/* Same with a caller-evaluated frequency table freqs_dev[dim / 2] (= exp(-ln(max_period) * k / half), unet.py:34): pass the
 * table of the reference's own torch.exp to reproduce its embedding to ~1e-6 on any host (a 1-ulp difference between two
 * expf implementations is amplified by t <= 999 to ~1e-4 otherwise). */
int clpk_timestep_embedding_table(const int64_t* t_dev, const float* freqs_dev, float* out_dev, int batch, int dim,
                                  void* stream);

/* y[m,n] = act(sum_k x[m,k]*w[n,k] + b[n]) (+ add[m,n]); act: 0 none, 1 SiLU.  fp32.  nn.Linear of
 * unet.py:47-53 and blocks.py:19-20. */
int clpk_linear(const float* x_dev, const float* w_dev, const float* b_dev, const float* add_dev, float* y_dev,
                int m, int n, int k, int act, void* stream);

/* FiLM.forward standalone (blocks.py:22-25): y[b,c,:] = x[b,c,:] * scale1p[b,c] + shift[b,c], NCHW fp32, scale1p = 1+s
 * (two separately rounded ops like ATen).  Inside the UNet this is fused into the conv1 epilogue instead. */
int clpk_film_apply(const float* x_nchw_dev, const float* scale1p_dev, const float* shift_dev, float* y_nchw_dev,
                    int batch, int ch, int hw, void* stream);

/* GroupNorm(groups, C) over NHWC fp32 input, optional SiLU, 16-bit (op_dtype) NHWC output = conv A operand
 * (blocks.py:33-36,41,43; unet.py:78,105).  ws_dev: scratch of clpk_groupnorm_ws_bytes(batch, hw, c, groups) bytes. */
int64_t clpk_groupnorm_ws_bytes(int batch, int hw, int c, int groups);
int clpk_groupnorm_silu(const float* x_nhwc_dev, const float* gamma_dev, const float* beta_dev, void* y_op_nhwc_dev,
                        void* ws_dev, int batch, int hw, int c, int groups, float eps, int silu, int op_dtype,
                        void* stream);

/* The two halves of clpk_groupnorm_silu for producers that already emitted per-tile statistics (conv epilogue):
 * finalize: the gn_partial buffer of clpk_conv_epilogue (layout there) -> stats[b][groups] float2 (mean, rstd), biased
 * variance, tiles combined with the parallel-variance formula (robust for |mean| >> std like torch's Welford pass;
 * n_per_group is informational, the counts come from the buffer);
 * apply: y = (x - mean) * rstd * gamma + beta [SiLU] in the 16-bit operand format. */
int clpk_groupnorm_finalize(const void* partial_dev, void* stats_dev, int batch, int slots, int groups,
                            double n_per_group, float eps, void* stream);
int clpk_groupnorm_apply(const float* x_nhwc_dev, const float* gamma_dev, const float* beta_dev, const void* stats_dev,
                         void* y_op_nhwc_dev, int batch, int hw, int c, int groups, int silu, int op_dtype, void* stream);

/* Convolution kinds understood by the implicit-GEMM kernel. */
#define CLPK_CONV_3X3_S1 0   /* Conv2d 3x3 stride 1 pad 1  (blocks.py:34,36; unet.py:79)   */
#define CLPK_CONV_3X3_S2 1   /* Conv2d 3x3 stride 2 pad 1  (unet.py:63)                    */
#define CLPK_CONVT_4X4_S2 2  /* ConvTranspose2d 4x4 stride 2 pad 1 (unet.py:75)            */
#define CLPK_CONV_1X1 3      /* pointwise conv / plain GEMM over NHWC pixels: the stem conv (unet.py:55) runs as this
                                over 32-wide im2col columns written by clpk_stem_im2col                          */

/* Repack a reference-layout fp32 weight (Conv2d: [Cout,Cin,kh,kw]; ConvTranspose2d: [Cin,Cout,4,4]) into the 16-bit
 * K-major GEMM layout the tcgen05 kernel reads: Conv: [Cout_pad][tap][Cin]; ConvT: [phase(4)][Cout_pad][tap(4)][Cin].
 * Returns the number of 16-bit elements written (or needed, when out_dev == NULL) or <0 on error. */
int64_t clpk_pack_conv_weight(const float* w_dev, void* out_op_dev, int kind, int cin, int cout, int op_dtype,
                              void* stream);

/* Epilogue description of one implicit-GEMM convolution launch.
 *   y = conv(x) + bias;  if film: y = y * film_scale1p[b,c] + film_shift[b,c]   (blocks.py:22-25, scale1p = 1+s)
 *   if resid: y += resid (same NHWC shape as the output; may alias out_f32)       (blocks.py:44, unet.py:104)
 *   outputs (any subset): out_f32 NHWC fp32, out_op NHWC in the operand format (feeds the next conv), out_nchw fp32
 *   NCHW restricted to cout_valid channels. */
typedef struct clpk_conv_epilogue {
  const float* bias;          /* dev [cout]                       */
  const float* film_scale1p;  /* dev, element (b,c) at [b*film_stride + c], or NULL */
  const float* film_shift;    /* dev, same indexing, or NULL      */
  int64_t film_stride;
  const float* resid;         /* dev NHWC fp32 or NULL            */
  float* out_f32;             /* dev NHWC fp32 or NULL            */
  void* out_op;               /* dev NHWC 16-bit (op_dtype) or NULL */
  float* out_nchw;            /* dev NCHW fp32 or NULL            */
  int cout_valid;             /* channels really present (<= cout_pad) */
  /* Fused GroupNorm statistics of the FINAL output values (after bias / FiLM / residual), accumulated in the epilogue as
   * shifted sums.  gn_partial = float buffer of 2 * batch * slots * groups + slots elements, slots =
   * clpk_conv_gn_slots(...), groups = cout / gn_cpg:  [b][slot][g] float2 (tile mean, tile M2 = sum (x - mean)^2) followed
   * by [slot] float element counts (geometry only, identical for every image).  Fold with clpk_groupnorm_finalize or
   * clpk_groupnorm_affine.  NULL = off.  Needs an NHWC output and gn_cpg in {4, 8, 16} or a multiple of 32. */
  void* gn_partial;
  int gn_cpg;
  /* Input transform fused into the A-operand path — the consumer-side half of GroupNorm [+ SiLU] (blocks.py:41,43;
   * unet.py:105) without a stand-alone normalisation pass over HBM:
   *   a[b,h,w,c] <- act(x[b,h,w,c] * in_scale[b*cin + c] + in_shift[b*cin + c]),  act = SiLU when in_silu != 0,
   * applied in shared memory (fp32 math) to every operand slab before the tensor cores read it; the conv's zero padding
   * stays zero.  Tables from clpk_groupnorm_affine.  Only for geometries with clpk_conv_in_affine_supported(...) == 1
   * (3x3 stride-1 convs on rows of >= 128 pixels, cout <= 128, cin % 64 == 0).  NULL = off. */
  const float* in_scale;
  const float* in_shift;
  int in_silu;
  /* Residual in the 16-bit operand format (NHWC, op_dtype) instead of fp32: y += resid_op, for a residual stream that is
   * kept in 16 bits only (blocks.py:44, unet.py:104 with the running sum rounded to fp16 once per block).  Exclusive with
   * `resid`; needs out_op and no out_f32 (may alias out_op: in-place update).  NULL = off. */
  const void* resid_op;
} clpk_conv_epilogue;

/* 1 when clpk_conv_igemm accepts in_scale / in_shift for this geometry, else 0. */
int clpk_conv_in_affine_supported(int kind, int h_in, int w_in, int cin, int cout);

/* Folds the per-tile statistics a conv epilogue wrote (gn_partial, `pieces` triples per slot: pieces = c / gn_cpg) into
 * the per-(image, channel) affine form of GroupNorm: scale[b*c + ch] = rstd * gamma[ch],
 * shift[b*c + ch] = beta[ch] - mean * rstd * gamma[ch]  (biased variance, eps inside the sqrt; blocks.py:33,35). */
int clpk_groupnorm_affine(const void* partial_dev, const float* gamma_dev, const float* beta_dev, float* scale_dev,
                          float* shift_dev, int batch, int slots, int pieces, int groups, int c, float eps, void* stream);

/* Number of partial-sum slots per image a conv of this geometry writes for a consumer GroupNorm with gn_cpg channels
 * per group (<= 0: the fused statistics are not available for this shape). */
int clpk_conv_gn_slots(int kind, int h_in, int w_in, int cout, int gn_cpg);

/* The UNet head out(out_norm(x)) (unet.py:78-79,105) as ONE kernel: GroupNorm (no activation) applied to the 16-bit NHWC
 * activation in shared memory (in_scale / in_shift [batch][c] from clpk_groupnorm_affine), then the 3x3 conv to 3
 * channels evaluated as a pointwise tcgen05 GEMM (N = 27 (tap, channel) columns) + a 9-point shift-add, fp32 NCHW output.
 * w_packed: clpk_pack_head_weight of the reference-layout weight [3][c][3][3] -> 16-bit [32][c]. */
int clpk_head_conv_supported(int h, int w, int c, int cout);
int clpk_pack_head_weight(const float* w_dev, void* out_op_dev, int c, int op_dtype, void* stream);
int clpk_head_conv(const void* x_op_nhwc_dev, const float* in_scale_dev, const float* in_shift_dev, const void* w_packed_dev,
                   const float* bias_dev, float* out_nchw_dev, int batch, int h, int w, int c, int op_dtype, void* stream);

/* Implicit-GEMM convolution on the 5th-gen tensor cores (tcgen05.mma, TMEM accumulators, TMA-fed).
 * x: 16-bit (op_dtype) NHWC [batch, h_in, w_in, cin]; w_packed from clpk_pack_conv_weight with the same op_dtype.
 * Output spatial size: S1: (h_in, w_in); S2: (h_in/2, w_in/2); ConvT: (2*h_in, 2*w_in). */
int clpk_conv_igemm(const void* x_op_nhwc_dev, const void* w_packed_dev, int kind, int batch, int h_in, int w_in,
                    int cin, int cout, int op_dtype, const clpk_conv_epilogue* ep, void* stream);

/* Same contract evaluated by a plain CUDA-core kernel (one thread per output element, fp32 accumulation of the same
 * 16-bit operands).  On-device cross-check for the tensor-core kernel in tests; never used by the plan. */
int clpk_conv_direct(const void* x_op_nhwc_dev, const void* w_packed_dev, int kind, int batch, int h_in, int w_in,
                     int cin, int cout, int op_dtype, const clpk_conv_epilogue* ep, void* stream);

/* Stem im2col: fp32 NCHW [B,cin,H,W] (cin*9 <= 32) -> 16-bit NHWC [B,H,W,32] with column k = c*9 + r*3 + s holding
 * x[b,c,h+r-1,w+s-1] (zero padded), columns >= cin*9 zero.  Feeds clpk_conv_igemm(kind = CLPK_CONV_1X1, cin = 32). */
int clpk_stem_im2col(const float* x_nchw_dev, void* cols_op_dev, int batch, int cin, int h, int w, int op_dtype,
                     void* stream);

/* in_conv (unet.py:55,88) on the CUDA cores: fp32 NCHW [B,cin,H,W] -> fp32 NHWC [B,H,W,cout], 3x3 s1 p1, fp32
 * arithmetic.  Kept as an exact-fp32 leaf op; the plan runs the stem on the tensor cores (im2col + CLPK_CONV_1X1). */
int clpk_conv_in(const float* x_nchw_dev, const float* w_dev /*[cout,cin,3,3]*/, const float* b_dev, float* y_nhwc_dev,
                 int batch, int cin, int h, int w, int cout, void* stream);

/* ---------------------------------------------------------------------------------------------------------------
 * Plan level: the whole CLIPCondUNet forward (unet.py:81-106) and the DDIM loop (ddim.py:21-46)
 * ------------------------------------------------------------------------------------------------------------- */

typedef struct clpk_unet_config {
  int z_dim;      /* 512 */
  int base;       /* 128 */
  int n_levels;   /* len(ch_mult) */
  int ch_mult[CLPK_MAX_LEVELS];
  int time_dim;   /* 256 */
  int img_ch;     /* 3 */
  int groups;     /* 8 */
  int op_dtype;   /* CLPK_OP_F16 (default) or CLPK_OP_BF16: tensor-core operand format */
} clpk_unet_config;

typedef struct clpk_plan clpk_plan;

/* Creates a plan for fixed (batch, H, W).  `names`/`ptrs`/`numels` describe the state dict (unet.py:45-79,
 * SURVEY Appendix B): n_params entries, names are the reference's parameter names, ptrs are DEVICE fp32 pointers that
 * only need to stay valid during this call (weights are repacked into plan-owned memory). */
int clpk_plan_create(const clpk_unet_config* cfg, int batch, int height, int width, int n_params,
                     const char* const* names, const float* const* ptrs_dev, const int64_t* numels,
                     clpk_plan** out_plan);
void clpk_plan_destroy(clpk_plan* plan);
/* Optional: the plan evaluates the timestep embedding with this frequency table [time_dim / 2] (copied) instead of the
 * device expf (see clpk_timestep_embedding_table).  Call before clpk_plan_prepare_ddim (a prepared loop is invalidated). */
int clpk_plan_set_time_freqs(clpk_plan* plan, const float* freqs_dev);
int64_t clpk_plan_device_bytes(const clpk_plan* plan);
/* algorithmic FLOPs of one forward at the plan's batch (conv + linear, 2*MAC) */
double clpk_plan_flops_per_forward(const clpk_plan* plan);
/* kernels launched by one forward / one DDIM step */
int clpk_plan_launches_per_forward(const clpk_plan* plan);

/* eps = CLIPCondUNet(x_t, z_clip, t):  x_t fp32 NCHW [B,img_ch,H,W], z_clip fp32 [B,z_dim], t int64 [B]. */
int clpk_unet_forward(clpk_plan* plan, const float* x_nchw_dev, const float* z_clip_dev, const int64_t* t_dev,
                      float* eps_nchw_dev, void* stream);

/* Prepare a DDIM run: `steps` sub-sampled timesteps ts[] (host int64, ddim.py:25) and coef[steps][5] (host fp32,
 * see clpk_ddim_step).  Builds the per-step conditioning tables and captures the CUDA graph of one step. */
int clpk_plan_prepare_ddim(clpk_plan* plan, int steps, const int64_t* ts_host, const float* coef_host, int use_graph,
                           void* stream);

/* Runs the prepared loop.  x_dev: fp32 NCHW, holds x_T on entry and the result on exit (unclamped, ddim.py:46).
 * noise_dev: NULL, or fp32 [steps][B*img_ch*H*W] of pre-drawn N(0,1) (parity runs);  if NULL and any coef sigma > 0
 * the noise is generated in-kernel with Philox4x32-10 keyed by (seed, step, element).
 * eps_trace_dev / x_trace_dev: NULL or [steps][...] buffers receiving every step's eps and input x (parity). */
int clpk_ddim_sample(clpk_plan* plan, const float* z_clip_dev, float* x_dev, const float* noise_dev, uint64_t seed,
                     float* eps_trace_dev, float* x_trace_dev, void* stream);

/* Measurement hooks (bench.py): eager DDIM steps with CUDA events around every launch.  ms_out6 / count_out6 are HOST
 * arrays of 6 entries, one per kernel class: 0 ResBlock 3x3 convs (tcgen05), 1 other tcgen05 convs, 2 GroupNorm
 * (stats + apply), 3 stem conv, 4 conditioning, 5 DDIM update.  Needs clpk_plan_prepare_ddim first. */
int clpk_plan_profile_steps(clpk_plan* plan, int iters, float* ms_out6, int* count_out6, void* stream);
/* Algorithmic work of one forward at the plan's batch: FLOPs of conv classes 0 and 1, fp32 elements read by GroupNorm. */
int clpk_plan_work_breakdown(const clpk_plan* plan, double* conv_res_flops, double* conv_other_flops,
                             double* gn_elements);
/* Algorithmic HBM bytes of all GroupNorm(+SiLU) applies of one forward: input read once (2 B / element when the input is
 * a 16-bit tensor, 4 B when fp32) + 16-bit operand written once. */
int clpk_plan_groupnorm_bytes(const clpk_plan* plan, double* bytes);

/* ---------------------------------------------------------------------------------------------------------------
 * Output side: reconstruct_diffusion.py:55-56, PKG/eval/metrics.py:16-29
 * ------------------------------------------------------------------------------------------------------------- */

/* u8[b,h,w,c] = uint8((clamp(x[b,c,h,w],-1,1) + 1) * 127.5)  (truncation) — reconstruct_diffusion.py:55-56 */
int clpk_to_uint8_hwc(const float* x_nchw_dev, uint8_t* out_hwc_dev, int batch, int ch, int h, int w, void* stream);

/* The BICUBIC-resized original of the eval loop (PKG/cli/eval.py:66-67: `Image.open(p).convert("RGB").resize((S,S),
 * Image.BICUBIC)`, then `float32(u8) / 127.5 - 1` as CHW) on the device, bit-identical to Pillow's 8-bit resampler.
 * clpk_resample_u8 is ONE separable pass over a uint8 tensor viewed as [outer][in_size][inner] -> [outer][out_size][inner]
 * (horizontal pass of an HWC image: outer = H, in = W, inner = C; vertical: outer = 1, in = H, inner = W*C; Pillow runs the
 * horizontal pass first and skips a pass whose size does not change).  bounds [out_size][2] = (first input index, count),
 * kk [out_size][ksize] = int32 coefficients with 22 fractional bits, exactly Pillow's precompute_coeffs +
 * normalize_coeffs_8bpc (host: eval/resample.py::bicubic_coeffs).  out = clip8((2^21 + sum kk*in) >> 22). */
int clpk_resample_u8(const uint8_t* src_dev, uint8_t* dst_dev, const int32_t* bounds_dev, const int32_t* kk_dev, int ksize,
                     int64_t outer, int in_size, int out_size, int inner, void* stream);
int clpk_u8_hwc_to_float_chw(const uint8_t* src_hwc_dev, float* dst_chw_dev, int h, int w, int c, void* stream);

/* per-image PSNR in the uint8 domain (metrics.py:16-29): sq_err_sum[b] = sum((u8(a)-u8(b))^2) as exact int64;
 * the host finishes 20*log10(255/sqrt(mse)).  Inputs fp32 in [-1,1], any layout as long as both agree. */
int clpk_psnr_sqerr_u8(const float* a_dev, const float* b_dev, int64_t* sq_err_sum_dev, int batch, int64_t per_image,
                       void* stream);

/* DDPM helpers (PKG/diffusion/scheduler.py:46-68): out[b,i] = (ca[b]*x[b,i] + cb[b]*y[b,i]) [/ cdiv[b]] [clamped to
 * [-1,1]], every operation individually rounded like the reference's ATen expressions.  q_sample: ca = sqrt_ac[t],
 * cb = sqrt_1m_ac[t]; predict_x0_from_eps: ca = 1, cb = -sqrt_1m_ac[t], cdiv = sqrt_ac[t]; posterior mean of
 * p_mean_variance: ca = coef1[t], cb = coef2[t].  ca / cb / cdiv: device fp32 [batch] (cdiv may be NULL). */
int clpk_ddpm_combine(const float* x_dev, const float* y_dev, const float* ca_dev, const float* cb_dev, const float* cdiv_dev,
                      float* out_dev, int batch, int64_t per_image, int clamp, void* stream);

/* per-image SSIM in the uint8 domain (metrics.py:32-46: skimage structural_similarity(HWC uint8, data_range=255,
 * channel_axis=-1) with its defaults — 7x7 uniform window, K1 0.01, K2 0.03, sample covariance, float64, 3-pixel border
 * cropped, mean over pixels then over channels).  a, b: fp32 NCHW [batch,ch,h,w] in [-1,1]; out: fp64 [batch];
 * ws: clpk_ssim_ws_bytes(...) bytes of device scratch.  h, w < 7 is an error (skimage raises ValueError). */
int64_t clpk_ssim_ws_bytes(int batch, int ch, int h, int w);
int clpk_ssim_u8(const float* a_nchw_dev, const float* b_nchw_dev, double* ssim_dev, void* ws_dev, int batch, int ch, int h,
                 int w, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* CLPK_H_ */
