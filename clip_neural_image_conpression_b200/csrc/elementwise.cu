// elementwise.cu — library plumbing + the small bandwidth/latency-bound kernels of the decode path:
// dequantise + L2 renorm, quantiser fit/encode, DDIM update, timestep embedding, small fp32 Linear, uint8 post-process,
// uint8-domain squared error (PSNR).  sm_100a.
#include "common.cuh"
#include "kernels.cuh"
#include "ddim_math.cuh"

#include <stdlib.h>

#include <atomic>
#include <stdarg.h>
#include <string.h>

namespace clpk {

static thread_local char g_err[512] = "";
static std::atomic<uint64_t> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
void count_launch(int n) { g_launches.fetch_add((uint64_t)n, std::memory_order_relaxed); }

bool pdl_enabled() {
  static const bool on = [] { const char* e = getenv("CLPK_PDL"); return e && atoi(e) != 0; }();
  return on;
}

// SM count of the CURRENT device (cached per device: one process may drive several GPUs)
int num_sms() {
  static std::atomic<int> cache[64];
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 148;
  if (dev < 0 || dev >= 64) dev = 0;
  int sms = cache[dev].load(std::memory_order_relaxed);
  if (sms == 0) {
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) sms = 148;
    cache[dev].store(sms, std::memory_order_relaxed);
  }
  return sms;
}

// ------------------------------------------------------------------------------------------------ dequant + L2 norm
// One block per row (image).  The dequantisation is __fmul_rn then __fadd_rn (numpy does two separately rounded ops,
// reconstruct_diffusion.py:43); the norm is an fp32 tree sum -> sqrt -> max(.,1e-9) -> true division (:21-23).
__global__ void dequant_l2norm_kernel(const uint8_t* __restrict__ q, const float* __restrict__ scale,
                                      const float* __restrict__ zero, float* __restrict__ z, float* __restrict__ z_raw,
                                      int dim, int l2norm) {
  __shared__ float red[32];
  const int b = blockIdx.x;
  const uint8_t* qr = q + (long long)b * dim;
  float ss = 0.f;
  for (int d = threadIdx.x; d < dim; d += blockDim.x) {
    const float v = __fadd_rn(__fmul_rn((float)qr[d], scale[d]), zero[d]);
    if (z_raw) z_raw[(long long)b * dim + d] = v;
    if (!l2norm) z[(long long)b * dim + d] = v;
    ss = __fadd_rn(ss, __fmul_rn(v, v));
  }
  if (!l2norm) return;
  ss = warp_sum(ss);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = ss;
  __syncthreads();
  if (threadIdx.x < 32) {
    float t = (threadIdx.x < (blockDim.x >> 5)) ? red[threadIdx.x] : 0.f;
    t = warp_sum(t);
    if (threadIdx.x == 0) red[0] = fmaxf(__fsqrt_rn(t), 1e-9f);
  }
  __syncthreads();
  const float nrm = red[0];
  for (int d = threadIdx.x; d < dim; d += blockDim.x) {
    const float v = __fadd_rn(__fmul_rn((float)qr[d], scale[d]), zero[d]);
    z[(long long)b * dim + d] = __fdiv_rn(v, nrm);
  }
}

// q = uint8(clamp(rint((x - zero) / scale), 0, 255)) — torch.round is round-half-even == rintf (quantizer.py:32)
__global__ void quant_encode_kernel(const float* __restrict__ x, const float* __restrict__ scale,
                                    const float* __restrict__ zero, uint8_t* __restrict__ q, long long total, int dim) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int d = (int)(i % dim);
    float v = rintf(__fdiv_rn(__fsub_rn(x[i], zero[d]), scale[d]));
    v = fminf(fmaxf(v, 0.f), 255.f);
    q[i] = (uint8_t)v;
  }
}

// per-channel min / max over n rows (quantizer.py:23-26); one thread per channel, coalesced across channels
__global__ void quant_fit_kernel(const float* __restrict__ x, float* __restrict__ scale, float* __restrict__ zero, int n,
                                 int dim) {
  const int d = blockIdx.x * blockDim.x + threadIdx.x;
  if (d >= dim) return;
  float mn = x[d], mx = x[d];
  for (int i = 1; i < n; ++i) {
    const float v = x[(long long)i * dim + d];
    mn = fminf(mn, v);
    mx = fmaxf(mx, v);
  }
  zero[d] = mn;
  scale[d] = __fdiv_rn(fmaxf(__fsub_rn(mx, mn), 1e-8f), 255.f);
}

// ------------------------------------------------------------------------------------------------ DDIM update
// (element math + noise source: ddim_math.cuh)
// coef_tab: [steps][5] on the device; the step index, noise source and seed come from the device-resident DdimRun so
// that one captured graph serves every step of every run.
__global__ void ddim_step_kernel(const float* __restrict__ x, const float* __restrict__ eps,
                                 const float* __restrict__ coef_tab, const DdimRun* __restrict__ run,
                                 float* __restrict__ x_out, long long n) {
  pdl_prologue_done();
  const int step = run->step;
  const float* noise = run->noise;
  const long long noise_step_stride = run->noise_step_stride;
  const unsigned long long seed = run->seed;
  const DdimCoef k = ddim_load_coef(coef_tab, step);
  const bool stochastic = k.sigma > 0.f;  // ddim.py:44 `if eta > 0 and sigma_t > 0`
  const float* nz = (noise && stochastic) ? noise + (long long)step * noise_step_stride : nullptr;
  const long long n4 = n >> 2;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const float4 xv = reinterpret_cast<const float4*>(x)[i];
    const float4 ev = reinterpret_cast<const float4*>(eps)[i];
    const float4 o = ddim_update4(xv, ev, k, stochastic, nz, i, step, seed);
    reinterpret_cast<float4*>(x_out)[i] = o;
  }
  // tail (n % 4)
  for (long long i = (n4 << 2) + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    float o = ddim_update(x[i], eps[i], k);
    if (stochastic) {
      float z;
      if (nz) {
        z = nz[i];
      } else {
        uint32_t ctr[4] = {(uint32_t)i, (uint32_t)(i >> 32), (uint32_t)step, 0x7461696cu};
        philox4x32_10(ctr, (uint32_t)seed, (uint32_t)(seed >> 32));
        float z1;
        box_muller(ctr[0], ctr[1], z, z1);
      }
      o = __fadd_rn(o, __fmul_rn(k.sigma, z));
    }
    x_out[i] = o;
  }
}

int launch_ddim_step(const float* x, const float* eps, const float* coef_tab_dev, const DdimRun* run_dev, float* x_out,
                     long long n, cudaStream_t stream) {
  const long long work = (n + 3) / 4;
  const int blocks = (int)std::min<long long>((work + 255) / 256, (long long)num_sms() * 8);
  CLPK_CHECK_CUDA(launch_kernel_pdl(ddim_step_kernel, dim3(std::max(blocks, 1)), dim3(256), 0, stream, x, eps, coef_tab_dev, run_dev,
                                    x_out, n));
  CLPK_CHECK_LAUNCH();
  return CLPK_OK;
}

__global__ void ddim_advance_kernel(DdimRun* run) {
  pdl_prologue_done();
  run->step += 1;
}
int launch_ddim_advance(DdimRun* run_dev, cudaStream_t stream) {
  CLPK_CHECK_CUDA(launch_kernel_pdl(ddim_advance_kernel, dim3(1), dim3(1), 0, stream, run_dev));
  CLPK_CHECK_LAUNCH();
  return CLPK_OK;
}

// DDPM helpers (PKG/diffusion/scheduler.py:46-68: q_sample, predict_x0_from_eps, the posterior mean of p_mean_variance):
//   out[b, i] = clamp?( (ca[b] * x[b, i] + cb[b] * y[b, i]) / cdiv[b]? )
// with per-sample coefficients gathered from the schedule tables on the host side.  Every multiply / add / divide is
// individually rounded (no FMA contraction), so the result is bit-identical to the reference's ATen expression:
//   q_sample            : ca = sqrt_ac[t], cb = sqrt_1m_ac[t]                     (a*x0 + b*noise)
//   predict_x0_from_eps : ca = 1, cb = -sqrt_1m_ac[t], cdiv = sqrt_ac[t]           ((x_t - b*eps) / a; 1*x and -(b*eps) exact)
//   posterior mean      : ca = coef1, cb = coef2                                   (c1*x0_pred + c2*x_t)
__global__ void ddpm_combine_kernel(const float* __restrict__ x, const float* __restrict__ y, const float* __restrict__ ca,
                                    const float* __restrict__ cb, const float* __restrict__ cdiv, float* __restrict__ out,
                                    long long per_image, int clamp) {
  const int b = blockIdx.y;
  const float a = ca[b], bb = cb[b];
  const float d = cdiv ? cdiv[b] : 1.0f;
  const float* px = x + (long long)b * per_image;
  const float* py = y + (long long)b * per_image;
  float* po = out + (long long)b * per_image;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < per_image; i += (long long)gridDim.x * blockDim.x) {
    float v = __fadd_rn(__fmul_rn(a, px[i]), __fmul_rn(bb, py[i]));
    if (cdiv) v = __fdiv_rn(v, d);
    if (clamp && v == v) v = fminf(fmaxf(v, -1.0f), 1.0f);  // torch.clamp(-1, 1) propagates NaN; fmaxf/fminf would drop it
    po[i] = v;
  }
}

// Holds the stream for `ns` nanoseconds (one thread).  Used by the profiling pass: while it runs, the host enqueues a whole
// DDIM step (launches + event records), so the bracketed kernels then execute back to back and the event timestamps do
// not contain host launch latency.
__global__ void delay_kernel(long long ns) {
  long long t0;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
  for (;;) {
    long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    if (t - t0 >= ns) break;
    __nanosleep(1000);
  }
}
int launch_delay(long long ns, cudaStream_t stream) {
  delay_kernel<<<1, 1, 0, stream>>>(ns);
  CLPK_CHECK_LAUNCH();
  return CLPK_OK;
}

// h[b,:] = zemb[b,:] + ht_tab[run->step,:]   (unet.py:86 with the step-invariant / batch-invariant halves hoisted)
__global__ void cond_combine_kernel(const float* __restrict__ zemb, const float* __restrict__ ht_tab,
                                    const DdimRun* __restrict__ run, float* __restrict__ h, int batch, int dim) {
  pdl_prologue_done();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= batch * dim) return;
  const int d = i % dim;
  h[i] = __fadd_rn(ht_tab[(long long)run->step * dim + d], zemb[i]);
}
int launch_cond_combine(const float* zemb, const float* ht_tab, const DdimRun* run, float* h, int batch, int dim,
                        cudaStream_t stream) {
  CLPK_CHECK_CUDA(launch_kernel_pdl(cond_combine_kernel, dim3((batch * dim + 255) / 256), dim3(256), 0, stream, zemb, ht_tab, run, h,
                                    batch, dim));
  CLPK_CHECK_LAUNCH();
  return CLPK_OK;
}

__global__ void add_const_kernel(float* p, float v, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] += v;
}
int launch_add_const(float* p, float v, int n, cudaStream_t stream) {
  add_const_kernel<<<(n + 255) / 256, 256, 0, stream>>>(p, v, n);
  CLPK_CHECK_LAUNCH();
  return CLPK_OK;
}

// ------------------------------------------------------------------------------------------------ timestep embedding
// unet.py:33-36: freqs = exp(float32(-ln(max_period)) * k / half) evaluated in fp32; args = float(t) * freqs.
// freqs != NULL: the caller's frequency table [half] (the host evaluates the expression above with its own expf, e.g. the
// reference's torch-CPU exp: a 1-ulp difference in a frequency is amplified by t <= 999 to ~1e-4 in the angle).
__global__ void timestep_embedding_kernel(const int64_t* __restrict__ t, float* __restrict__ out, int batch, int dim,
                                          float neg_log_period, const float* __restrict__ freqs) {
  const int half = dim / 2;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= batch * half) return;
  const int b = i / half, k = i - b * half;
  const float f = freqs ? __ldg(freqs + k) : expf(__fdiv_rn(__fmul_rn(neg_log_period, (float)k), (float)half));
  const float a = __fmul_rn((float)t[b], f);
  out[(long long)b * dim + k] = cosf(a);
  out[(long long)b * dim + half + k] = sinf(a);
  if ((dim & 1) && k == 0) out[(long long)b * dim + dim - 1] = 0.f;
}

int launch_timestep_embedding(const int64_t* t, float* out, int batch, int dim, float max_period, cudaStream_t stream,
                              const float* freqs) {
  const int total = batch * (dim / 2);
  // float32(-math.log(max_period)): the reference evaluates the log in double and rounds once (unet.py:33)
  const float neg_log_period = (float)(-log((double)max_period));
  timestep_embedding_kernel<<<(total + 127) / 128, 128, 0, stream>>>(t, out, batch, dim, neg_log_period, freqs);
  CLPK_CHECK_LAUNCH();
  return CLPK_OK;
}

// ------------------------------------------------------------------------------------------------ small fp32 Linear
// y[m,n] = act(dot(x[m,:], w[n,:]) + b[n]) (+ add[m % add_rows, n]).  One warp per output column n, looping over rows
// in groups of 8 so each weight row is read once per 8 rows.  M is the batch (<= a few hundred): FLOPs are negligible.
__global__ void linear_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ b,
                              const float* __restrict__ add, int add_rows, float* __restrict__ y, int m, int n, int k,
                              int act) {
  pdl_prologue_done();
  const int col = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (col >= n) return;
  const float* wr = w + (long long)col * k;
  const float bias = b ? b[col] : 0.f;
  for (int m0 = blockIdx.y * 8; m0 < m; m0 += gridDim.y * 8) {
    float acc[8];
#pragma unroll
    for (int r = 0; r < 8; ++r) acc[r] = 0.f;
    for (int kk = lane; kk < k; kk += 32) {
      const float wv = wr[kk];
#pragma unroll
      for (int r = 0; r < 8; ++r)
        if (m0 + r < m) acc[r] = fmaf(x[(long long)(m0 + r) * k + kk], wv, acc[r]);
    }
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      const float s = warp_sum(acc[r]);
      if (lane == 0 && m0 + r < m) {
        float v = s + bias;
        if (act == 1) v = v / (1.0f + expf(-v));
        if (add) v += add[(long long)((m0 + r) % add_rows) * n + col];
        y[(long long)(m0 + r) * n + col] = v;
      }
    }
  }
}

int launch_linear(const float* x, const float* w, const float* b, const float* add, int add_rows, float* y, int m, int n,
                  int k, int act, cudaStream_t stream) {
  dim3 grid((n + 7) / 8, std::min((m + 7) / 8, 64));
  CLPK_CHECK_CUDA(launch_kernel_pdl(linear_kernel, grid, dim3(256), 0, stream, x, w, b, add, add_rows > 0 ? add_rows : m, y, m, n, k,
                                    act));
  CLPK_CHECK_LAUNCH();
  return CLPK_OK;
}

// ------------------------------------------------------------------------------------------------ standalone FiLM
// y = x * scale1p[b,c] + shift[b,c] over NCHW (blocks.py:22-25); inside the UNet this lives in the conv1 epilogue.
__global__ void film_apply_kernel(const float* __restrict__ x, const float* __restrict__ sc, const float* __restrict__ sh,
                                  float* __restrict__ y, long long total, int hw) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long bc = i / hw;
    y[i] = __fadd_rn(__fmul_rn(x[i], sc[bc]), sh[bc]);
  }
}

// ------------------------------------------------------------------------------------------------ output post-process
__device__ __forceinline__ uint8_t to_u8_trunc(float v) {
  // ((clamp(v,-1,1) + 1) * 127.5).astype(uint8): separately rounded add and mul, C truncation
  const float c = fminf(fmaxf(v, -1.f), 1.f);
  return (uint8_t)(int)__fmul_rn(__fadd_rn(c, 1.0f), 127.5f);
}

__global__ void to_uint8_hwc_kernel(const float* __restrict__ x, uint8_t* __restrict__ out, int batch, int ch, int h,
                                    int w) {
  const long long total = (long long)batch * h * w * ch;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % ch);
    long long r = i / ch;
    const int xx = (int)(r % w); r /= w;
    const int yy = (int)(r % h);
    const int b = (int)(r / h);
    out[i] = to_u8_trunc(x[(((long long)b * ch + c) * h + yy) * w + xx]);
  }
}

// metrics.py:16-19 _to_uint8: ((img+1)*127.5).clip(0,255).astype(uint8);  :26 squared difference summed exactly in int64
__device__ __forceinline__ int metric_u8(float v) {
  const float s = __fmul_rn(__fadd_rn(v, 1.0f), 127.5f);
  return (int)fminf(fmaxf(s, 0.f), 255.f);
}
__global__ void psnr_sqerr_kernel(const float* __restrict__ a, const float* __restrict__ b,
                                  unsigned long long* __restrict__ out, long long per_image) {
  const int img = blockIdx.y;
  const float* pa = a + (long long)img * per_image;
  const float* pb = b + (long long)img * per_image;
  unsigned long long acc = 0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < per_image;
       i += (long long)gridDim.x * blockDim.x) {
    const int d = metric_u8(pa[i]) - metric_u8(pb[i]);
    acc += (unsigned long long)(d * d);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0 && acc) atomicAdd(out + img, acc);  // integer atomics: order independent, exact
}

// ------------------------------------------------------------------------------------------------ SSIM
// metrics.py:32-46 -> skimage.metrics.structural_similarity(x1, x2, data_range=255, channel_axis=-1) on the uint8 images
// of _to_uint8 (metrics.py:16-19).  scikit-image is neither vendored nor installed (parity unpinned); this
// follows its published defaults: 7x7 uniform window (mean filter), K1 = 0.01, K2 = 0.03, sample covariance
// (NP / (NP - 1)), float64 arithmetic, S = ((2 ux uy + C1)(2 vxy + C2)) / ((ux^2 + uy^2 + C1)(vx + vy + C2)), the
// (win - 1) / 2 border cropped, mean over the remaining pixels per channel, then the mean over channels.
// The window sums are exact integers here (skimage's two separable float64 passes round each 1-D mean: <= 1e-15 rel).
constexpr int kSsimWin = 7, kSsimPad = 3, kSsimTileW = 32, kSsimTileH = 8;

__global__ void __launch_bounds__(kSsimTileW * kSsimTileH)
ssim_partial_kernel(const float* __restrict__ a, const float* __restrict__ b, double* __restrict__ partial, int h, int w,
                    int tiles_x) {
  __shared__ uint8_t sa[kSsimTileH + 2 * kSsimPad][kSsimTileW + 2 * kSsimPad];
  __shared__ uint8_t sb[kSsimTileH + 2 * kSsimPad][kSsimTileW + 2 * kSsimPad];
  __shared__ double wsum[kSsimTileH];
  const int tile = blockIdx.x, plane = blockIdx.y;  // plane = image * channels + channel (NCHW)
  const int ty = tile / tiles_x, tx = tile - ty * tiles_x;
  const int oy0 = kSsimPad + ty * kSsimTileH, ox0 = kSsimPad + tx * kSsimTileW;  // first output pixel of the tile
  const float* pa = a + (long long)plane * h * w;
  const float* pb = b + (long long)plane * h * w;
  const int tid = threadIdx.y * kSsimTileW + threadIdx.x;
  constexpr int kHaloW = kSsimTileW + 2 * kSsimPad, kHaloH = kSsimTileH + 2 * kSsimPad;
  for (int i = tid; i < kHaloW * kHaloH; i += kSsimTileW * kSsimTileH) {
    const int ly = i / kHaloW, lx = i - ly * kHaloW;
    const int gy = oy0 - kSsimPad + ly, gx = ox0 - kSsimPad + lx;
    const bool in = gy < h && gx < w;
    sa[ly][lx] = in ? (uint8_t)metric_u8(pa[(long long)gy * w + gx]) : (uint8_t)0;
    sb[ly][lx] = in ? (uint8_t)metric_u8(pb[(long long)gy * w + gx]) : (uint8_t)0;
  }
  __syncthreads();
  const int oy = oy0 + threadIdx.y, ox = ox0 + threadIdx.x;
  double s = 0.0;
  if (oy < h - kSsimPad && ox < w - kSsimPad) {
    int sx = 0, sy = 0, sxx = 0, syy = 0, sxy = 0;
#pragma unroll
    for (int dy = 0; dy < kSsimWin; ++dy)
#pragma unroll
      for (int dx = 0; dx < kSsimWin; ++dx) {
        const int x = sa[threadIdx.y + dy][threadIdx.x + dx], y = sb[threadIdx.y + dy][threadIdx.x + dx];
        sx += x; sy += y; sxx += x * x; syy += y * y; sxy += x * y;
      }
    constexpr double np_ = (double)(kSsimWin * kSsimWin), cov_norm = np_ / (np_ - 1.0);
    const double c1 = (0.01 * 255.0) * (0.01 * 255.0), c2 = (0.03 * 255.0) * (0.03 * 255.0);
    const double ux = sx / np_, uy = sy / np_;
    const double vx = cov_norm * (sxx / np_ - ux * ux), vy = cov_norm * (syy / np_ - uy * uy);
    const double vxy = cov_norm * (sxy / np_ - ux * uy);
    s = ((2.0 * ux * uy + c1) * (2.0 * vxy + c2)) / ((ux * ux + uy * uy + c1) * (vx + vy + c2));
  }
  // fixed-order block sum: shuffle tree per warp (= tile row), then the rows in order
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (threadIdx.x == 0) wsum[threadIdx.y] = s;
  __syncthreads();
  if (tid == 0) {
    double t = 0.0;
    for (int r = 0; r < kSsimTileH; ++r) t += wsum[r];
    partial[(long long)plane * gridDim.x + tile] = t;
  }
}

// one warp per image: per-channel mean of S over the cropped pixels, then the mean over channels
__global__ void ssim_fold_kernel(const double* __restrict__ partial, double* __restrict__ out, int ch, int tiles,
                                 double pixels_per_channel) {
  const int img = blockIdx.x, lane = threadIdx.x;
  double acc = 0.0;
  for (int c = 0; c < ch; ++c) {
    const double* pp = partial + ((long long)img * ch + c) * tiles;
    double s = 0.0;
    for (int k = lane; k < tiles; k += 32) s += pp[k];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    acc += s / pixels_per_channel;
  }
  if (lane == 0) out[img] = acc / (double)ch;
}

}  // namespace clpk

using namespace clpk;

extern "C" int clpk_ddpm_combine(const float* x, const float* y, const float* ca, const float* cb, const float* cdiv,
                                 float* out, int batch, int64_t per_image, int clamp, void* stream) {
  CLPK_REQUIRE(x && y && ca && cb && out && batch > 0 && per_image > 0, "clpk_ddpm_combine: bad arguments");
  CLPK_REQUIRE(batch <= 65535, "clpk_ddpm_combine: batch too large");
  dim3 grid((unsigned)std::min<long long>((per_image + 255) / 256, (long long)num_sms() * 4), (unsigned)batch);
  ddpm_combine_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(x, y, ca, cb, cdiv, out, per_image, clamp);
  CLPK_CHECK_LAUNCH();
  return CLPK_OK;
}

extern "C" int64_t clpk_ssim_ws_bytes(int batch, int ch, int h, int w) {
  if (batch <= 0 || ch <= 0 || h < kSsimWin || w < kSsimWin) return -1;
  const long long tiles = (long long)((w - 2 * kSsimPad + kSsimTileW - 1) / kSsimTileW) *
                          ((h - 2 * kSsimPad + kSsimTileH - 1) / kSsimTileH);
  return (int64_t)(tiles * batch * ch * (long long)sizeof(double));
}

extern "C" int clpk_ssim_u8(const float* a, const float* b, double* out, void* ws, int batch, int ch, int h, int w,
                            void* stream) {
  CLPK_REQUIRE(a && b && out && ws && batch > 0 && ch > 0, "clpk_ssim_u8: bad arguments");
  CLPK_REQUIRE(h >= kSsimWin && w >= kSsimWin, "clpk_ssim_u8: win_size 7 exceeds the image extent (%d x %d)", h, w);
  const int tiles_x = (w - 2 * kSsimPad + kSsimTileW - 1) / kSsimTileW;
  const int tiles_y = (h - 2 * kSsimPad + kSsimTileH - 1) / kSsimTileH;
  CLPK_REQUIRE((long long)batch * ch <= 65535, "clpk_ssim_u8: too many image planes");
  dim3 grid((unsigned)(tiles_x * tiles_y), (unsigned)(batch * ch));
  dim3 block(kSsimTileW, kSsimTileH);
  ssim_partial_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(a, b, reinterpret_cast<double*>(ws), h, w, tiles_x);
  CLPK_CHECK_LAUNCH();
  ssim_fold_kernel<<<batch, 32, 0, (cudaStream_t)stream>>>(reinterpret_cast<const double*>(ws), out, ch, tiles_x * tiles_y,
                                                           (double)(h - 2 * kSsimPad) * (double)(w - 2 * kSsimPad));
  CLPK_CHECK_LAUNCH();
  return CLPK_OK;
}

extern "C" const char* clpk_last_error(void) { return g_err; }
extern "C" int clpk_version(void) { return 100; }
extern "C" uint64_t clpk_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

extern "C" int clpk_dequant_l2norm_u8(const uint8_t* q, const float* scale, const float* zero, float* z, float* z_raw,
                                      int batch, int dim, int l2norm, void* stream) {
  CLPK_REQUIRE(batch >= 0 && dim > 0, "clpk_dequant_l2norm_u8: bad shape");
  if (batch == 0) return CLPK_OK;
  CLPK_REQUIRE(q && scale && zero && z, "clpk_dequant_l2norm_u8: null pointer");
  dequant_l2norm_kernel<<<batch, 256, 0, (cudaStream_t)stream>>>(q, scale, zero, z, z_raw, dim, l2norm);
  CLPK_CHECK_LAUNCH();
  return CLPK_OK;
}

extern "C" int clpk_quant_encode_u8(const float* x, const float* scale, const float* zero, uint8_t* q, int batch, int dim,
                                    void* stream) {
  CLPK_REQUIRE(batch >= 0 && dim > 0, "clpk_quant_encode_u8: bad shape");
  if (batch == 0) return CLPK_OK;
  CLPK_REQUIRE(x && scale && zero && q, "clpk_quant_encode_u8: null pointer");
  const long long total = (long long)batch * dim;
  quant_encode_kernel<<<(int)std::min<long long>((total + 255) / 256, 4096), 256, 0, (cudaStream_t)stream>>>(
      x, scale, zero, q, total, dim);
  CLPK_CHECK_LAUNCH();
  return CLPK_OK;
}

extern "C" int clpk_quant_fit(const float* x, float* scale, float* zero, int n, int dim, void* stream) {
  CLPK_REQUIRE(n > 0 && dim > 0 && x && scale && zero, "clpk_quant_fit: bad arguments");
  quant_fit_kernel<<<(dim + 127) / 128, 128, 0, (cudaStream_t)stream>>>(x, scale, zero, n, dim);
  CLPK_CHECK_LAUNCH();
  return CLPK_OK;
}

extern "C" int clpk_ddim_step(const float* x, const float* eps, const float* noise, const float* coef5_host,
                              float* x_out, int64_t n, void* stream) {
  CLPK_REQUIRE(x && eps && coef5_host && x_out && n >= 0, "clpk_ddim_step: bad arguments");
  if (n == 0) return CLPK_OK;
  CLPK_REQUIRE((((uintptr_t)x | (uintptr_t)eps | (uintptr_t)x_out | (uintptr_t)noise) & 15) == 0,
               "clpk_ddim_step: pointers must be 16-byte aligned");
  // coefficients + run state travel through a small device block so the kernel is the same one the graph replays
  struct { float coef[8]; DdimRun run; } host_blk;
  memset(&host_blk, 0, sizeof(host_blk));
  memcpy(host_blk.coef, coef5_host, 5 * sizeof(float));
  host_blk.run.step = 0;
  host_blk.run.noise = noise;
  char* blk = nullptr;
  CLPK_CHECK_CUDA(cudaMalloc(&blk, sizeof(host_blk)));
  cudaError_t e = cudaMemcpyAsync(blk, &host_blk, sizeof(host_blk), cudaMemcpyHostToDevice, (cudaStream_t)stream);
  int rc = CLPK_OK;
  if (e != cudaSuccess) { set_error("clpk_ddim_step: memcpy failed: %s", cudaGetErrorString(e)); rc = CLPK_ERR_CUDA; }
  if (!rc)
    rc = launch_ddim_step(x, eps, reinterpret_cast<const float*>(blk),
                          reinterpret_cast<const DdimRun*>(blk + 8 * sizeof(float)), x_out, n, (cudaStream_t)stream);
  cudaStreamSynchronize((cudaStream_t)stream);  // host_blk is on this stack frame; blk is freed right below
  cudaFree(blk);
  return rc;
}

extern "C" int clpk_timestep_embedding(const int64_t* t, float* out, int batch, int dim, float max_period,
                                       void* stream) {
  CLPK_REQUIRE(t && out && batch > 0 && dim >= 2, "clpk_timestep_embedding: bad arguments");
  return launch_timestep_embedding(t, out, batch, dim, max_period, (cudaStream_t)stream);
}

extern "C" int clpk_timestep_embedding_table(const int64_t* t, const float* freqs, float* out, int batch, int dim,
                                             void* stream) {
  CLPK_REQUIRE(t && freqs && out && batch > 0 && dim >= 2, "clpk_timestep_embedding_table: bad arguments");
  return launch_timestep_embedding(t, out, batch, dim, 10000.f, (cudaStream_t)stream, freqs);
}

extern "C" int clpk_linear(const float* x, const float* w, const float* b, const float* add, float* y, int m, int n,
                           int k, int act, void* stream) {
  CLPK_REQUIRE(x && w && y && m > 0 && n > 0 && k > 0, "clpk_linear: bad arguments");
  return launch_linear(x, w, b, add, m, y, m, n, k, act, (cudaStream_t)stream);
}

extern "C" int clpk_film_apply(const float* x, const float* scale1p, const float* shift, float* y, int batch, int ch,
                               int hw, void* stream) {
  CLPK_REQUIRE(x && scale1p && shift && y && batch > 0 && ch > 0 && hw > 0, "clpk_film_apply: bad arguments");
  const long long total = (long long)batch * ch * hw;
  film_apply_kernel<<<(int)std::min<long long>((total + 255) / 256, (long long)num_sms() * 16), 256, 0,
                      (cudaStream_t)stream>>>(x, scale1p, shift, y, total, hw);
  CLPK_CHECK_LAUNCH();
  return CLPK_OK;
}

// ------------------------------------------------------------------------------------------------ BICUBIC originals
// PIL's 8-bit resampling (Pillow src/libImaging/Resample.c, what `Image.resize(size, Image.BICUBIC)` of
// PKG/cli/eval.py:66 runs): one separable pass = out[o] = clip8((2^21 + sum_k kk[o][k] * in[xmin[o] + k]) >> 22) with
// int32 fixed-point coefficients (22 fractional bits) precomputed per (input size, output size) on the host
// (eval/resample.py, bit-identical to PIL's doubles).  The tensor is viewed as [outer][in_size][inner] uint8:
// horizontal pass of an HWC image = (H, W, C), vertical pass = (1, H, W*C).  Integer arithmetic -> bit exact.
namespace clpk {
__global__ void resample_u8_kernel(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst, const int* __restrict__ bounds,
                                   const int* __restrict__ kk, int ksize, long long outer, int in_size, int out_size, int inner) {
  const long long total = outer * out_size * inner;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int i = (int)(idx % inner);
    const long long r = idx / inner;
    const int o = (int)(r % out_size);
    const long long u = r / out_size;
    const int xmin = __ldg(bounds + 2 * o), n = __ldg(bounds + 2 * o + 1);
    const uint8_t* sp = src + (u * in_size + xmin) * inner + i;
    const int* kp = kk + (long long)o * ksize;
    int acc = 1 << 21;
    for (int k = 0; k < n; ++k) acc += __ldg(kp + k) * (int)sp[(long long)k * inner];
    acc >>= 22;   // arithmetic shift: floor, like PIL's clip8 lookup index
    dst[idx] = (uint8_t)min(max(acc, 0), 255);
  }
}

// float CHW in [-1, 1] from uint8 HWC: (float32(u8) / 127.5f) - 1.0f, two separately rounded ops (eval.py:67 numpy)
__global__ void u8_hwc_to_float_chw_kernel(const uint8_t* __restrict__ src, float* __restrict__ dst, int h, int w, int c) {
  const long long total = (long long)h * w * c;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int x = (int)(idx % w);
    const long long r = idx / w;
    const int y = (int)(r % h);
    const int ch = (int)(r / h);
    dst[idx] = __fsub_rn(__fdiv_rn((float)src[((long long)y * w + x) * c + ch], 127.5f), 1.0f);
  }
}
}  // namespace clpk

extern "C" int clpk_resample_u8(const uint8_t* src, uint8_t* dst, const int32_t* bounds, const int32_t* kk, int ksize,
                                int64_t outer, int in_size, int out_size, int inner, void* stream) {
  CLPK_REQUIRE(src && dst && bounds && kk && ksize > 0 && outer > 0 && in_size > 0 && out_size > 0 && inner > 0,
               "clpk_resample_u8: bad arguments");
  const long long total = outer * out_size * inner;
  resample_u8_kernel<<<(int)std::min<long long>((total + 255) / 256, (long long)num_sms() * 16), 256, 0, (cudaStream_t)stream>>>(
      src, dst, bounds, kk, ksize, outer, in_size, out_size, inner);
  CLPK_CHECK_LAUNCH();
  return CLPK_OK;
}

extern "C" int clpk_u8_hwc_to_float_chw(const uint8_t* src, float* dst, int h, int w, int c, void* stream) {
  CLPK_REQUIRE(src && dst && h > 0 && w > 0 && c > 0, "clpk_u8_hwc_to_float_chw: bad arguments");
  const long long total = (long long)h * w * c;
  u8_hwc_to_float_chw_kernel<<<(int)std::min<long long>((total + 255) / 256, (long long)num_sms() * 16), 256, 0,
                               (cudaStream_t)stream>>>(src, dst, h, w, c);
  CLPK_CHECK_LAUNCH();
  return CLPK_OK;
}

extern "C" int clpk_to_uint8_hwc(const float* x, uint8_t* out, int batch, int ch, int h, int w, void* stream) {
  CLPK_REQUIRE(x && out && batch > 0 && ch > 0 && h > 0 && w > 0, "clpk_to_uint8_hwc: bad arguments");
  const long long total = (long long)batch * ch * h * w;
  to_uint8_hwc_kernel<<<(int)std::min<long long>((total + 255) / 256, (long long)num_sms() * 16), 256, 0,
                        (cudaStream_t)stream>>>(x, out, batch, ch, h, w);
  CLPK_CHECK_LAUNCH();
  return CLPK_OK;
}

extern "C" int clpk_psnr_sqerr_u8(const float* a, const float* b, int64_t* out, int batch, int64_t per_image,
                                  void* stream) {
  CLPK_REQUIRE(a && b && out && batch > 0 && per_image > 0, "clpk_psnr_sqerr_u8: bad arguments");
  CLPK_CHECK_CUDA(cudaMemsetAsync(out, 0, sizeof(int64_t) * batch, (cudaStream_t)stream));
  dim3 grid((unsigned)std::min<long long>((per_image + 255) / 256, 256), (unsigned)batch);
  psnr_sqerr_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(a, b, reinterpret_cast<unsigned long long*>(out),
                                                           per_image);
  CLPK_CHECK_LAUNCH();
  return CLPK_OK;
}
