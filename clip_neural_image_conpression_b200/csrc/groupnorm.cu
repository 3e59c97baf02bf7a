// groupnorm.cu — GroupNorm(groups, C) [+ SiLU] over NHWC fp32 activations, bf16 NHWC output (the conv A operand).
//
// Replaces nn.GroupNorm + nn.SiLU of the reference ResBlock (PKG/models/blocks.py:33,35,41,43) and out_norm
// (PKG/models/unet.py:78,105 — no SiLU there).  Statistics are per (image, group) over (C/groups)*H*W elements,
// biased variance, eps inside the sqrt (torch semantics).
//
// Two memory-bound passes (HBM roofline): (1) gn_stats: one read of x, fp32 per-thread sums, deterministic block
// reduction, per-block partials; the LAST block of each image folds the partials in fp64 and publishes (mean, rstd)
// — no float atomics, so results are run-to-run identical.  (2) gn_apply: one read of x, one bf16 write.
#include "kernels.cuh"

namespace clpk {

constexpr int kGnThreads = 256;
constexpr int kGnMaxJ = 4;

GnShape gn_shape(int batch, int hw, int c, int groups) {
  GnShape s;
  s.batch = batch; s.hw = hw; s.c = c; s.groups = groups;
  // enough blocks to fill the machine a few times, but >= 32 pixels per block
  const int want = std::max(1, (num_sms() * 4 + batch - 1) / batch);
  s.chunks = std::max(1, std::min(want, (hw + 31) / 32));
  return s;
}

// ws layout: int counter[B] (always at offset 0 and always left at zero, so one ws can serve every shape of a plan)
//            | float2 stats[B][G] | float2 partial[B][chunks][G]
static inline long long ws_counter_bytes(const GnShape& s) { return ((long long)s.batch * 4 + 255) / 256 * 256; }
static inline long long ws_stats_bytes(const GnShape& s) { return ((long long)s.batch * s.groups * 8 + 255) / 256 * 256; }
static inline long long ws_partial_bytes(const GnShape& s) { return (long long)s.batch * s.chunks * s.groups * 8; }
long long gn_ws_bytes(const GnShape& s) {
  return ((ws_counter_bytes(s) + ws_stats_bytes(s) + ws_partial_bytes(s) + 255) / 256) * 256;
}

template <int J>
__global__ void __launch_bounds__(kGnThreads)
gn_stats_kernel(const float* __restrict__ x, float2* __restrict__ partial, float2* __restrict__ stats,
                int* __restrict__ counter, int hw, int c, int groups, int chunks, int pix_per_chunk, float eps) {
  __shared__ float2 red[kGnThreads * J];
  __shared__ int is_last;
  pdl_prologue_done();
  const int b = blockIdx.y, chunk = blockIdx.x;
  const int qc = c >> 2;           // float4 quads per pixel
  const int cpg = c / groups;
  const int tid = threadIdx.x, T = blockDim.x;
  int ppb, poff, q[J];
  if (J == 1) {
    // ppb whole pixels per block iteration; threads beyond ppb*qc (block rounded up to a warp multiple) idle
    ppb = kGnThreads / qc; poff = tid / qc; q[0] = (poff < ppb) ? tid - poff * qc : qc;
  } else {
    ppb = 1; poff = 0;
#pragma unroll
    for (int j = 0; j < J; ++j) q[j] = tid + j * T;
  }
  // SHIFTED sums (robust against |mean| >> std): every thread subtracts K[g] = the image's first element of its
  // quad's group — one value per (image, group), identical in every block, so partials add up without re-basing
  float s[J], ss[J], kq[J];
#pragma unroll
  for (int j = 0; j < J; ++j) {
    s[j] = 0.f; ss[j] = 0.f;
    kq[j] = (q[j] < qc) ? __ldg(x + (long long)b * hw * c + ((q[j] * 4) / cpg) * cpg) : 0.f;
  }
  const int p0 = chunk * pix_per_chunk;
  const int p1 = min(hw, p0 + pix_per_chunk);
  const float4* xb = reinterpret_cast<const float4*>(x) + (long long)b * hw * qc;
  int p = p0 + poff;
  // 4 pixels in flight per thread
  for (; p + 3 * ppb < p1; p += 4 * ppb) {
#pragma unroll
    for (int j = 0; j < J; ++j) {
      if (q[j] < qc) {
        float4 v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) v[u] = __ldg(xb + (long long)(p + u * ppb) * qc + q[j]);
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const float a0 = v[u].x - kq[j], a1 = v[u].y - kq[j], a2 = v[u].z - kq[j], a3 = v[u].w - kq[j];
          s[j] += (a0 + a1) + (a2 + a3);
          ss[j] += (a0 * a0 + a1 * a1) + (a2 * a2 + a3 * a3);
        }
      }
    }
  }
  for (; p < p1; p += ppb) {
#pragma unroll
    for (int j = 0; j < J; ++j) {
      if (q[j] < qc) {
        const float4 v = __ldg(xb + (long long)p * qc + q[j]);
        const float a0 = v.x - kq[j], a1 = v.y - kq[j], a2 = v.z - kq[j], a3 = v.w - kq[j];
        s[j] += (a0 + a1) + (a2 + a3);
        ss[j] += (a0 * a0 + a1 * a1) + (a2 * a2 + a3 * a3);
      }
    }
  }
#pragma unroll
  for (int j = 0; j < J; ++j) red[j * T + tid] = make_float2(s[j], ss[j]);
  __syncthreads();
  // fixed-order reduction: warp w folds the entries of groups w, w+nwarps, ...
  const int lane = tid & 31, warp = tid >> 5, nwarps = T >> 5;
  for (int g = warp; g < groups; g += nwarps) {
    float gs = 0.f, gss = 0.f;
    for (int e = lane; e < J * T; e += 32) {
      const int et = e % T;
      const int eq = (J == 1) ? (et % qc) : e;  // J>1: entry index == quad index
      const bool live = (J == 1) ? (et < ppb * qc) : (eq < qc);
      if (live && (eq * 4) / cpg == g) { gs += red[e].x; gss += red[e].y; }
    }
    gs = warp_sum(gs); gss = warp_sum(gss);
    if (lane == 0) partial[((long long)b * chunks + chunk) * groups + g] = make_float2(gs, gss);
  }
  // last block of this image publishes (mean, rstd)
  __threadfence();
  __syncthreads();
  if (tid == 0) is_last = (atomicAdd(counter + b, 1) == chunks - 1);
  __syncthreads();
  if (is_last) {
    __threadfence();
    // warp w folds the partials of groups w, w+nwarps, ...: lanes stride over the chunks (independent loads in
    // flight), then a fixed-order fp64 shuffle tree -> deterministic and latency-parallel
    for (int g = warp; g < groups; g += nwarps) {
      double S = 0.0, SS = 0.0;
      const float2* pp = partial + (long long)b * chunks * groups + g;
      for (int k = lane; k < chunks; k += 32) {
        const float2 v = __ldcg(pp + (long long)k * groups);
        S += (double)v.x; SS += (double)v.y;
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        S += __shfl_xor_sync(0xffffffffu, S, o);
        SS += __shfl_xor_sync(0xffffffffu, SS, o);
      }
      if (lane == 0) {
        const double n = (double)hw * (double)cpg;
        const double dm = S / n;  // mean - K
        double var = SS / n - dm * dm;
        if (var < 0.0) var = 0.0;
        const double k = (double)__ldg(x + (long long)b * hw * c + g * cpg);
        stats[(long long)b * groups + g] = make_float2((float)(k + dm), (float)(1.0 / sqrt(var + (double)eps)));
      }
    }
    if (tid == 0) counter[b] = 0;  // ready for the next launch / graph replay
  }
}

// y = (x - mean) * rstd * gamma + beta, optional SiLU, 8 channels per thread and access (32 B fp32 or 16 B 16-bit in,
// 16 B out), evaluated as one FMA per element with (scale, shift) = (rstd*gamma, beta - mean*rstd*gamma).  Statistics
// come either from `stats` (mean, rstd) or — when `partial` is given — are folded in-kernel from the per-tile partial
// sums a conv epilogue wrote (same fixed-order fp64 fold as gn_finalize_kernel, one warp per group, redundantly per
// block: a few KB of L2 reads instead of a separate launch).
// HBM-bound: every thread keeps kGnUnroll independent 16/32-byte loads in flight (2048 threads/SM x 16 B alone is below
// the ~44 KB/SM that 6.5 TB/s x ~1 us of loaded latency needs), SiLU costs ONE MUFU op per element
// (x*sigmoid(x) = h + h*tanh(h), h = x/2, tanh.approx.f32) instead of ex2 + rcp, and when the grid stride is a multiple
// of the octets per pixel (kHoist) each thread owns one channel octet, so its 8 (scale, shift) pairs live in registers.
constexpr int kGnUnroll = 4;

template <bool kIn16>
__device__ __forceinline__ void gn_load8(const void* __restrict__ xin, long long idx, uint4& a, uint4& b) {
  if (kIn16) {
    a = __ldcs(reinterpret_cast<const uint4*>(xin) + idx);
  } else {
    const uint4* xb = reinterpret_cast<const uint4*>(xin) + 2 * idx;
    a = __ldcs(xb);
    b = __ldcs(xb + 1);
  }
}

template <bool kIn16>
__device__ __forceinline__ uint4 gn_transform8(const uint4& a, const uint4& b, const float (&sc)[8], const float (&sh)[8],
                                               bool silu, bool f16) {
  float v[8];
  if (kIn16) {
    const uint32_t w4[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      v[2 * k] = from_op((uint16_t)(w4[k] & 0xffffu), f16);
      v[2 * k + 1] = from_op((uint16_t)(w4[k] >> 16), f16);
    }
  } else {
    v[0] = __uint_as_float(a.x); v[1] = __uint_as_float(a.y); v[2] = __uint_as_float(a.z); v[3] = __uint_as_float(a.w);
    v[4] = __uint_as_float(b.x); v[5] = __uint_as_float(b.y); v[6] = __uint_as_float(b.z); v[7] = __uint_as_float(b.w);
  }
  float r[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    r[k] = fmaf(v[k], sc[k], sh[k]);
    if (silu) r[k] = silu_tanh(r[k]);
  }
  return make_uint4(pack_op2(r[0], r[1], f16), pack_op2(r[2], r[3], f16), pack_op2(r[4], r[5], f16),
                    pack_op2(r[6], r[7], f16));
}

// (scale, shift) of the 8 channels starting at `ch`
__device__ __forceinline__ void gn_affine8(const float* __restrict__ gamma, const float* __restrict__ beta,
                                           const float2* st_s, int ch, int cpg, float (&sc)[8], float (&sh)[8]) {
  const float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma + ch)), g1 = __ldg(reinterpret_cast<const float4*>(gamma + ch) + 1);
  const float4 b0 = __ldg(reinterpret_cast<const float4*>(beta + ch)), b1 = __ldg(reinterpret_cast<const float4*>(beta + ch) + 1);
  const float2 s0 = st_s[ch / cpg], s1 = st_s[(ch + 4) / cpg];
  const float g[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
  const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const float2 s = k < 4 ? s0 : s1;
    sc[k] = s.y * g[k];
    sh[k] = fmaf(-s.x, sc[k], bb[k]);
  }
}


// Folds the per-tile statistics a conv epilogue wrote (conv_igemm.cu, "Fused GroupNorm statistics") into (mean, rstd)
// of group g of image b; executed by one whole warp (lanes stride over the entries, fixed-order shuffle tree ->
// deterministic).  partial[b][slot][piece] = (tile mean, tile M2): `pieces` pairs per slot, group g owns pieces
// [g*m, (g+1)*m), m = pieces / groups (m > 1: group sizes such as 24 or 48 that the epilogue sums in 8- or 16-channel
// pieces); counts[slot] = elements behind every pair of that slot (geometry only, shared by all images).
// Combination (Chan et al.) relative to the FIRST pair's mean m0, so every accumulated term is of the order of the spread
// of the tile means (fp32 is ample):
//   N = sum n, A = sum n (mean_i - m0), Q = sum n (mean_i - m0)^2, M = sum M2_i
//   mean = m0 + A / N,   var = (M + Q - A^2 / N) / N
__device__ __forceinline__ float2 gn_fold_partials(const float2* __restrict__ partial, const float* __restrict__ counts,
                                                   int b, int g, int slots, int pieces, int groups, float eps, int lane) {
  const int m = pieces / groups;
  const float2* pp = partial + (long long)b * slots * pieces + g * m;
  const float m0 = __ldg(pp).x;
  float A = 0.f, Q = 0.f, M = 0.f, N = 0.f;
#pragma unroll 4
  for (int k = lane; k < slots * m; k += 32) {
    const int slot = k / m, j = k - slot * m;
    const float2 v = __ldg(pp + (long long)slot * pieces + j);
    const float n = __ldg(counts + slot);
    const float d = v.x - m0, nd = n * d;
    A += nd; Q = fmaf(nd, d, Q); M += v.y; N += n;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    A += __shfl_xor_sync(0xffffffffu, A, o);
    Q += __shfl_xor_sync(0xffffffffu, Q, o);
    M += __shfl_xor_sync(0xffffffffu, M, o);
    N += __shfl_xor_sync(0xffffffffu, N, o);
  }
  const float inv_n = 1.0f / N, dm = A * inv_n;
  const float var = fmaxf((M + fmaf(-A, dm, Q)) * inv_n, 0.f);
  return make_float2(m0 + dm, 1.0f / sqrtf(var + eps));
}

template <bool kIn16, bool kHoist>
__global__ void __launch_bounds__(kGnThreads, 4)
gn_apply_kernel(const void* __restrict__ xin, const float* __restrict__ gamma, const float* __restrict__ beta,
                const float2* __restrict__ stats, const float2* __restrict__ partial, int slots, int pieces, double n_per_group,
                float eps, uint16_t* __restrict__ y, int hw, int c, int groups, int silu, int op_f16) {
  __shared__ float2 st_s[32];
  pdl_prologue_done();
  const bool f16 = op_f16 != 0;
  const int b = blockIdx.y;
  if (partial) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int g = warp; g < groups; g += kGnThreads / 32) {
      const float* counts = reinterpret_cast<const float*>(partial + (long long)gridDim.y * slots * pieces);
      const float2 st = gn_fold_partials(partial, counts, b, g, slots, pieces, groups, eps, lane);
      if (lane == 0) st_s[g] = st;
    }
  } else {
    if (threadIdx.x < groups) st_s[threadIdx.x] = stats[(long long)b * groups + threadIdx.x];
  }
  __syncthreads();
  const unsigned oc = (unsigned)c >> 3;  // 8-channel octets per pixel
  const int cpg = c / groups;
  const unsigned total = (unsigned)hw * oc;
  const unsigned stride = gridDim.x * blockDim.x;
  const long long base = (long long)b * total;
  uint4* yb = reinterpret_cast<uint4*>(y) + base;
  const bool act = silu != 0;
  unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
  float sc[8], sh[8];
  if (kHoist) gn_affine8(gamma, beta, st_s, (int)(i % oc) * 8, cpg, sc, sh);
  // main loop: kGnUnroll independent loads in flight, then transform + store
  for (; (unsigned long long)i + (unsigned long long)(kGnUnroll - 1) * stride < total; i += kGnUnroll * stride) {
    uint4 a[kGnUnroll], bq[kGnUnroll];
#pragma unroll
    for (int u = 0; u < kGnUnroll; ++u) gn_load8<kIn16>(xin, base + i + (long long)u * stride, a[u], bq[u]);
#pragma unroll
    for (int u = 0; u < kGnUnroll; ++u) {
      if (!kHoist) gn_affine8(gamma, beta, st_s, (int)((i + u * stride) % oc) * 8, cpg, sc, sh);
      yb[i + u * stride] = gn_transform8<kIn16>(a[u], bq[u], sc, sh, act, f16);
    }
  }
  for (; i < total; i += stride) {
    uint4 a, bq;
    gn_load8<kIn16>(xin, base + i, a, bq);
    if (!kHoist) gn_affine8(gamma, beta, st_s, (int)(i % oc) * 8, cpg, sc, sh);
    yb[i] = gn_transform8<kIn16>(a, bq, sc, sh, act, f16);
  }
}

// Folds per-tile partial sums produced by a conv epilogue: partial[b][slots][groups] (sum, sum of squares) ->
// stats[b][g] = (mean, rstd).  One warp per (image, group); lanes stride over the slots, fp64 shuffle tree: deterministic.
__global__ void gn_finalize_kernel(const float2* __restrict__ partial, float2* __restrict__ stats, int slots, int groups,
                                   float eps) {
  const int b = blockIdx.x;
  const int g = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (g >= groups) return;
  const float* counts = reinterpret_cast<const float*>(partial + (long long)gridDim.x * slots * groups);
  const float2 st = gn_fold_partials(partial, counts, b, g, slots, groups, groups, eps, lane);
  if (lane == 0) stats[(long long)b * groups + g] = st;
}

// partial -> per-(image, channel) affine form of GroupNorm for kernels that normalise their own A operand
// (clpk_conv_epilogue.in_scale / in_shift, clpk_head_conv): one block per (group, image); its 4 warps fold a quarter of
// the slots each (all relative to the first pair's mean, so the quarter sums simply add), warp 0 combines in fixed order.
__global__ void __launch_bounds__(128)
gn_affine_kernel(const float2* __restrict__ partial, const float* __restrict__ gamma, const float* __restrict__ beta,
                 float* __restrict__ scale, float* __restrict__ shift, int slots, int pieces, int groups, int c, float eps) {
  __shared__ float4 part_s[4];
  __shared__ float2 st_s;
  pdl_prologue_done();
  const int g = blockIdx.x, b = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float* counts = reinterpret_cast<const float*>(partial + (long long)gridDim.y * slots * pieces);
  const int m = pieces / groups;
  const float2* pp = partial + (long long)b * slots * pieces + g * m;
  const float m0 = __ldg(pp).x;
  float A = 0.f, Q = 0.f, M = 0.f, N = 0.f;
#pragma unroll 4
  for (int k = warp * 32 + lane; k < slots * m; k += 128) {
    const int slot = k / m, j = k - slot * m;
    const float2 v = __ldg(pp + (long long)slot * pieces + j);
    const float n = __ldg(counts + slot);
    const float d = v.x - m0, nd = n * d;
    A += nd; Q = fmaf(nd, d, Q); M += v.y; N += n;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    A += __shfl_xor_sync(0xffffffffu, A, o);
    Q += __shfl_xor_sync(0xffffffffu, Q, o);
    M += __shfl_xor_sync(0xffffffffu, M, o);
    N += __shfl_xor_sync(0xffffffffu, N, o);
  }
  if (lane == 0) part_s[warp] = make_float4(A, Q, M, N);
  __syncthreads();
  if (threadIdx.x == 0) {
    float4 t = part_s[0];
#pragma unroll
    for (int w = 1; w < 4; ++w) { t.x += part_s[w].x; t.y += part_s[w].y; t.z += part_s[w].z; t.w += part_s[w].w; }
    const float inv_n = 1.0f / t.w, dm = t.x * inv_n;
    const float var = fmaxf((t.z + fmaf(-t.x, dm, t.y)) * inv_n, 0.f);
    st_s = make_float2(m0 + dm, 1.0f / sqrtf(var + eps));
  }
  __syncthreads();
  const float2 st = st_s;
  const int cpg = c / groups;
  for (int k = threadIdx.x; k < cpg; k += 128) {
    const int ch = g * cpg + k;
    const float sc = st.y * __ldg(gamma + ch);
    scale[(long long)b * c + ch] = sc;
    shift[(long long)b * c + ch] = fmaf(-st.x, sc, __ldg(beta + ch));
  }
}

int launch_gn_affine(const float2* partial, const float* gamma, const float* beta, float* scale, float* shift, int batch,
                     int slots, int pieces, int groups, int c, float eps, cudaStream_t stream) {
  CLPK_REQUIRE(groups > 0 && c % groups == 0 && pieces % groups == 0, "GroupNorm affine: bad group layout");
  CLPK_CHECK_CUDA(launch_kernel_pdl(gn_affine_kernel, dim3(groups, batch), dim3(128), 0, stream, partial, gamma, beta, scale,
                                    shift, slots, pieces, groups, c, eps));
  CLPK_CHECK_LAUNCH();
  return CLPK_OK;
}

int launch_gn_finalize(const float2* partial, float2* stats, int batch, int slots, int groups, float eps,
                       cudaStream_t stream) {
  CLPK_REQUIRE(groups <= 32, "GroupNorm finalize supports <= 32 groups");
  gn_finalize_kernel<<<batch, 32 * groups, 0, stream>>>(partial, stats, slots, groups, eps);
  CLPK_CHECK_LAUNCH();
  return CLPK_OK;
}

// x: fp32 NHWC (x_is_16 == 0) or 16-bit NHWC in the operand format.  Exactly one of stats / partial is used.
int launch_gn_apply_ex(const void* x, int x_is_16, const float* gamma, const float* beta, const float2* stats,
                       const float2* partial, int slots, int pieces, double n_per_group, float eps, void* y_op,
                       const GnShape& s,
                       int silu, int op_dtype, cudaStream_t stream) {
  const long long octs = (long long)s.hw * (s.c / 8);
  CLPK_REQUIRE(s.c % 8 == 0 && s.c % s.groups == 0 && (s.c / s.groups) % 4 == 0,
               "GroupNorm needs C %% 8 == 0 and (C/groups) %% 4 == 0 (C=%d groups=%d)", s.c, s.groups);
  CLPK_REQUIRE(octs < (1ll << 31) && s.groups <= 32, "GroupNorm: image too large or too many groups");
  const int per_img_blocks = (int)std::min<long long>((octs + kGnThreads - 1) / kGnThreads,
                                                      std::max(1, num_sms() * 4 / s.batch));  // one wave at 4 blocks / SM
  dim3 agrid(std::max(per_img_blocks, 1), s.batch);
  uint16_t* y = reinterpret_cast<uint16_t*>(y_op);
  const bool hoist = ((long long)agrid.x * kGnThreads) % (s.c / 8) == 0;  // every thread stays on one channel octet
  const bool f16 = op_dtype == CLPK_OP_F16;
#define CLPK_GN_APPLY(IN16, HOIST)                                                                                    \
  launch_err = launch_kernel_pdl(gn_apply_kernel<IN16, HOIST>, agrid, dim3(kGnThreads), 0, stream, x, gamma, beta, stats,  \
                                 partial, slots, pieces, n_per_group, eps, y, s.hw, s.c, s.groups, silu, (int)f16)
  cudaError_t launch_err = cudaSuccess;
  if (x_is_16) { if (hoist) CLPK_GN_APPLY(true, true); else CLPK_GN_APPLY(true, false); }
  else { if (hoist) CLPK_GN_APPLY(false, true); else CLPK_GN_APPLY(false, false); }
#undef CLPK_GN_APPLY
  CLPK_CHECK_CUDA(launch_err);
  CLPK_CHECK_LAUNCH();
  return CLPK_OK;
}

int launch_gn_apply(const float* x, const float* gamma, const float* beta, const float2* stats, void* y_op,
                    const GnShape& s, int silu, int op_dtype, cudaStream_t stream) {
  return launch_gn_apply_ex(x, 0, gamma, beta, stats, nullptr, 0, s.groups, 1.0, 0.f, y_op, s, silu, op_dtype, stream);
}

// statistics pass only: leaves (mean, rstd) in the stats area of ws (see layout above) and returns its address
int launch_gn_stats(const float* x, void* ws, const GnShape& s, float eps, const float2** stats_out, cudaStream_t stream);

int launch_groupnorm(const float* x, const float* gamma, const float* beta, void* y_bf16, void* ws, const GnShape& s,
                     float eps, int silu, int op_dtype, cudaStream_t stream) {
  const float2* stats = nullptr;
  int rc = launch_gn_stats(x, ws, s, eps, &stats, stream);
  if (rc) return rc;
  return launch_gn_apply(x, gamma, beta, stats, y_bf16, s, silu, op_dtype, stream);
}

int launch_gn_stats(const float* x, void* ws, const GnShape& s, float eps, const float2** stats_out,
                    cudaStream_t stream) {
  CLPK_REQUIRE(s.c % 8 == 0 && s.c % s.groups == 0 && (s.c / s.groups) % 4 == 0,
               "GroupNorm needs C %% 8 == 0 and (C/groups) %% 4 == 0 (C=%d groups=%d)", s.c, s.groups);
  CLPK_REQUIRE(s.c <= 4 * kGnMaxJ * kGnThreads, "GroupNorm supports C <= %d", 4 * kGnMaxJ * kGnThreads);
  int* counter = reinterpret_cast<int*>(ws);
  float2* stats = reinterpret_cast<float2*>(reinterpret_cast<char*>(ws) + ws_counter_bytes(s));
  float2* partial = reinterpret_cast<float2*>(reinterpret_cast<char*>(stats) + ws_stats_bytes(s));
  const int qc = s.c / 4;
  const int pix_per_chunk = (s.hw + s.chunks - 1) / s.chunks;
  dim3 grid(s.chunks, s.batch);
  if (qc <= kGnThreads) {
    const int threads = ((kGnThreads / qc) * qc + 31) / 32 * 32;
    CLPK_CHECK_CUDA(launch_kernel_pdl(gn_stats_kernel<1>, grid, dim3(threads), 0, stream, x, partial, stats, counter, s.hw,
                                      s.c, s.groups, s.chunks, pix_per_chunk, eps));
  } else {
    const int j = (qc + kGnThreads - 1) / kGnThreads;
    if (j == 2)
      CLPK_CHECK_CUDA(launch_kernel_pdl(gn_stats_kernel<2>, grid, dim3(kGnThreads), 0, stream, x, partial, stats, counter, s.hw,
                                      s.c, s.groups, s.chunks, pix_per_chunk, eps));
    else if (j == 3)
      CLPK_CHECK_CUDA(launch_kernel_pdl(gn_stats_kernel<3>, grid, dim3(kGnThreads), 0, stream, x, partial, stats, counter, s.hw,
                                      s.c, s.groups, s.chunks, pix_per_chunk, eps));
    else
      CLPK_CHECK_CUDA(launch_kernel_pdl(gn_stats_kernel<4>, grid, dim3(kGnThreads), 0, stream, x, partial, stats, counter, s.hw,
                                      s.c, s.groups, s.chunks, pix_per_chunk, eps));
  }
  CLPK_CHECK_LAUNCH();
  *stats_out = stats;
  return CLPK_OK;
}

}  // namespace clpk

using namespace clpk;

extern "C" int64_t clpk_groupnorm_ws_bytes(int batch, int hw, int c, int groups) {
  if (batch <= 0 || hw <= 0 || c <= 0 || groups <= 0) return -1;
  return gn_ws_bytes(gn_shape(batch, hw, c, groups));
}

extern "C" int clpk_groupnorm_silu(const float* x, const float* gamma, const float* beta, void* y, void* ws, int batch,
                                   int hw, int c, int groups, float eps, int silu, int op_dtype, void* stream) {
  CLPK_REQUIRE(op_dtype == CLPK_OP_BF16 || op_dtype == CLPK_OP_F16, "clpk_groupnorm_silu: operand dtype %d unknown", op_dtype);
  CLPK_REQUIRE(x && gamma && beta && y && ws && batch > 0 && hw > 0 && c > 0 && groups > 0,
               "clpk_groupnorm_silu: bad arguments");
  const GnShape s = gn_shape(batch, hw, c, groups);
  // the leaf entry point cannot assume the counters (start of ws) are zero
  CLPK_CHECK_CUDA(cudaMemsetAsync(ws, 0, (size_t)batch * 4, (cudaStream_t)stream));
  return launch_groupnorm(x, gamma, beta, y, ws, s, eps, silu, op_dtype, (cudaStream_t)stream);
}

extern "C" int clpk_groupnorm_finalize(const void* partial, void* stats, int batch, int slots, int groups,
                                       double n_per_group, float eps, void* stream) {
  CLPK_REQUIRE(partial && stats && batch > 0 && slots > 0 && groups > 0 && n_per_group > 0,
               "clpk_groupnorm_finalize: bad arguments");
  return launch_gn_finalize(reinterpret_cast<const float2*>(partial), reinterpret_cast<float2*>(stats), batch, slots,
                            groups, eps, (cudaStream_t)stream);
}

extern "C" int clpk_groupnorm_affine(const void* partial, const float* gamma, const float* beta, float* scale, float* shift,
                                     int batch, int slots, int pieces, int groups, int c, float eps, void* stream) {
  CLPK_REQUIRE(partial && gamma && beta && scale && shift && batch > 0 && slots > 0 && pieces > 0 && groups > 0 && c > 0,
               "clpk_groupnorm_affine: bad arguments");
  return launch_gn_affine(reinterpret_cast<const float2*>(partial), gamma, beta, scale, shift, batch, slots, pieces, groups,
                          c, eps, (cudaStream_t)stream);
}

extern "C" int clpk_groupnorm_apply(const float* x, const float* gamma, const float* beta, const void* stats, void* y,
                                    int batch, int hw, int c, int groups, int silu, int op_dtype, void* stream) {
  CLPK_REQUIRE(x && gamma && beta && stats && y && batch > 0 && hw > 0 && c > 0 && groups > 0,
               "clpk_groupnorm_apply: bad arguments");
  CLPK_REQUIRE(op_dtype == CLPK_OP_BF16 || op_dtype == CLPK_OP_F16, "clpk_groupnorm_apply: operand dtype %d unknown", op_dtype);
  return launch_gn_apply(x, gamma, beta, reinterpret_cast<const float2*>(stats), y, gn_shape(batch, hw, c, groups), silu,
                         op_dtype, (cudaStream_t)stream);
}
