// head_conv.cu — the UNet head  out(out_norm(x))  (PKG/models/unet.py:78-79,105) as ONE bandwidth-bound kernel.
//
//   out_norm = GroupNorm(8, C) WITHOUT activation, out = Conv2d(C -> img_ch (3), 3x3, pad 1).
//
// As an implicit GEMM this conv is hostile to the tensor cores: N = 3 pads to 16, and an M = 128 MMA costs its A-operand
// read whatever N is, so the 9 taps cost 9 A reads per pixel (85 us at batch 8, 256 px, plus 46 us for the stand-alone
// out_norm pass).  Here the conv is evaluated as a POINTWISE GEMM followed by a 9-point shift-add:
//
//   D[pixel p, n = (r*3+s)*3 + co] = sum_ci  norm(x)[p, ci] * W[co, ci, r, s]          (one K = C GEMM, N = 27 -> 32)
//   out[h, w, co] = bias[co] + sum_{r,s} D[(h+r-1, w+s-1), (r*3+s)*3 + co]             (zero outside the image)
//
// so every input pixel passes the tensor core once (8 MMAs of 128x32x16 per 128 pixels at C = 128 instead of 72 of
// 128x16x16), and the kernel reads the 16-bit activation exactly once from HBM.  out_norm is applied to the A tile in
// shared memory before the MMA reads it (per-(image, channel) scale / shift from clpk_groupnorm_affine; this kernel
// issues ~8 MMAs per 32 KB tile, so — unlike in the 3x3 slab mainloop — shared-memory bandwidth is plentiful).
//
// One CTA owns a band of consecutive image rows; the D rows live in a 4-deep shared-memory ring (tap-major, one halo
// column left and right kept at zero; tap rows stored shifted so that the shift-add reads aligned float4), and output
// row g is emitted as soon as D rows g-1, g, g+1 are there.
// Warp roles (448 threads): warp 0 TMA producer, warp 1 MMA issuer (+ TMEM alloc), warps 2..9 normalise the landed tile,
// warps 10..13 move accumulators TMEM -> D ring and do the shift-add + NCHW fp32 store.
#include "conv_igemm.cuh"
#include "kernels.cuh"
#include "ddim_math.cuh"

#include <algorithm>
#include <mutex>
#include <stdlib.h>

namespace clpk {

constexpr int kHeadXfWarps = 8;                       // warps normalising the landed tile (the kernel's critical resource)
constexpr int kHeadThreads = 64 + 32 * kHeadXfWarps + 128;
constexpr int kHeadTileM = 128;
constexpr int kHeadN = 32;          // 27 used: n = (r*3 + s)*3 + co
constexpr int kHeadTaps = 27;
constexpr int kHeadRing = 4;
constexpr int kHeadMaxStages = 6;
constexpr int kHeadMaxKpt = 4;      // C <= 256: the per-thread (scale, shift) of all channel blocks stay in registers
constexpr int kHeadSmemBudget = 232448;

struct HeadParams {
  int batch, h, w, c;
  int tiles_w, kpt, stages;
  int rows_total, band;
  int op_f16;
  int dpitch;                 // floats per tap row of the D ring: w + 8 (shifted storage, see emit_row)
  const float* scale;         // [batch][c]
  const float* shift;
  const float* bias;          // [3]
  float* out;                 // NCHW fp32 [batch][3][h][w]
  // optional DDIM update fused into the epilogue (ddim.py:36-45): x <- update(x, eps) for the run's current step, in place
  // on ddim_x (same NCHW shape as `out`, which still receives eps).  All three NULL = plain head.
  float* ddim_x;
  const float* ddim_coef;     // [steps][5]
  const DdimRun* ddim_run;
};

struct __align__(8) HeadBarriers {
  uint64_t full[kHeadMaxStages];
  uint64_t ready[kHeadMaxStages];
  uint64_t empty[kHeadMaxStages];
  uint64_t tmem_full[2];
  uint64_t tmem_empty[2];
  uint64_t wbar;
  uint32_t tmem_base;
};

__global__ void __launch_bounds__(kHeadThreads, 1)
head_conv_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_w,
                 const __grid_constant__ HeadParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int stage_bytes = p.kpt * kHeadTileM * 128;           // kpt boxes of 128 pixels x 64 channels (16-bit)
  uint8_t* smem_a = smem;
  uint8_t* smem_w = smem_a + (size_t)p.stages * stage_bytes;  // kpt atoms of 32 rows x 128 B
  float* dring = reinterpret_cast<float*>(smem_w + (size_t)p.kpt * kHeadN * 128);
  HeadBarriers* bars = reinterpret_cast<HeadBarriers*>(dring + (size_t)kHeadRing * kHeadTaps * p.dpitch);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if ((smem_u32(smem) & 1023u) != 0) {
    if (threadIdx.x == 0) printf("clpk: head kernel shared memory is not 1024-byte aligned\n");
    __trap();
  }
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_a);
    tma_prefetch_desc(&map_w);
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(&bars->full[s], 1);
      mbar_init(&bars->ready[s], kHeadXfWarps);
      mbar_init(&bars->empty[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&bars->tmem_full[s], 1);
      mbar_init(&bars->tmem_empty[s], 4);
    }
    mbar_init(&bars->wbar, 1);
    fence_mbar_init();
  }
  if (warp == 1) { tmem_alloc(&bars->tmem_base, 64u); tmem_relinquish(); }
  // zero the D ring once: the halo columns (index 0 and w + 1 of every tap row) are never written again
  for (int i = threadIdx.x; i < kHeadRing * kHeadTaps * p.dpitch; i += kHeadThreads) dring[i] = 0.f;
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = bars->tmem_base;
  pdl_prologue_done();

  // band of output rows of this CTA (global row g = b * h + y) and the D rows it needs
  const int g0 = blockIdx.x * p.band, g1 = min(g0 + p.band, p.rows_total);
  const int j0 = max(g0 - 1, 0), j1 = min(g1, p.rows_total - 1);  // D rows j0 .. j1 inclusive
  const int n_items = (j1 - j0 + 1) * p.tiles_w;

  if (g0 >= g1) {
    // (grid rounded up: nothing to do)
  } else if (warp == 0) {
    // ===================================================== TMA producer
    if (elect_one()) {
      mbar_arrive_expect_tx(&bars->wbar, (uint32_t)(p.kpt * kHeadN * 128));
      for (int kc = 0; kc < p.kpt; ++kc) tma_load_2d(smem_w + (size_t)kc * kHeadN * 128, &map_w, &bars->wbar, kc * 64, 0);
    }
    __syncwarp();
    int stage = 0;
    uint32_t phase = 0;
    for (int it = 0; it < n_items; ++it) {
      const int j = j0 + it / p.tiles_w, tw = it - (it / p.tiles_w) * p.tiles_w;
      const int b = j / p.h, y = j - b * p.h;
      mbar_wait(&bars->empty[stage], phase ^ 1u);
      if (elect_one()) {
        mbar_arrive_expect_tx(&bars->full[stage], (uint32_t)stage_bytes);
        for (int kc = 0; kc < p.kpt; ++kc)
          tma_load_5d(smem_a + (size_t)stage * stage_bytes + (size_t)kc * kHeadTileM * 128, &map_a, &bars->full[stage], kc * 64,
                      tw * kHeadTileM, 0, y, b);
      }
      __syncwarp();
      if (++stage == p.stages) { stage = 0; phase ^= 1u; }
    }
  } else if (warp == 1) {
    // ===================================================== MMA issuer: D[128 x 32] = A[128 x C] * W[32 x C]^T per tile
    const uint32_t idesc = make_idesc_16bit(kHeadTileM, kHeadN, p.op_f16 != 0);
    const uint64_t adesc0 = make_kmajor_desc<128>(smem_u32(smem_a));
    const uint64_t bdesc0 = make_kmajor_desc<128>(smem_u32(smem_w));
    mbar_wait(&bars->wbar, 0);
    tc_fence_after();
    int stage = 0;
    uint32_t phase = 0;
    for (int it = 0; it < n_items; ++it) {
      const int as = it & 1;
      mbar_wait(&bars->tmem_empty[as], ((uint32_t)(it >> 1) & 1u) ^ 1u);
      mbar_wait(&bars->ready[stage], phase);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t tmem_d = tmem_base + (uint32_t)(as * kHeadN);
        const uint64_t adesc = adesc0 + (uint64_t)((stage * stage_bytes) >> 4);
        for (int kc = 0; kc < p.kpt; ++kc) {
#pragma unroll
          for (int kk = 0; kk < 4; ++kk)
            umma_f16kind(tmem_d, adesc + (uint64_t)((kc * kHeadTileM * 128) >> 4) + 2u * kk,
                         bdesc0 + (uint64_t)((kc * kHeadN * 128) >> 4) + 2u * kk, idesc, (kc | kk) ? 1u : 0u);
        }
        umma_commit(&bars->empty[stage]);
        umma_commit(&bars->tmem_full[as]);
      }
      __syncwarp();
      if (++stage == p.stages) { stage = 0; phase ^= 1u; }
    }
  } else if (warp < 2 + kHeadXfWarps) {
    // ===================================================== transform warps: out_norm in place on the landed tile
    constexpr int kXfThreads = 32 * kHeadXfWarps, kRowStep = kXfThreads / 8, kRows = kHeadTileM / kRowStep;
    const int xt = threadIdx.x - 64;
    const int jp = xt & 7, i0 = xt >> 3;     // 16-byte piece (8 channels) and first row; rows i0 + kRowStep * k
    const bool f16 = p.op_f16 != 0;
    int cur_b = -1, stage = 0;
    uint32_t phase = 0;
    // (scale, shift) of this thread's 8 channels of every channel block live in REGISTERS and are reloaded only when the
    // band crosses into the next image: reading them from a shared table per tile cost as many smem wavefronts as the
    // tile itself (profiles/ncu_full_r2.txt)
    float sc[kHeadMaxKpt][8], sh[kHeadMaxKpt][8];
    for (int it = 0; it < n_items; ++it) {
      const int j = j0 + it / p.tiles_w, tw = it - (it / p.tiles_w) * p.tiles_w;
      const int b = j / p.h;
      if (b != cur_b) {
#pragma unroll
        for (int kc = 0; kc < kHeadMaxKpt; ++kc) {
          if (kc < p.kpt) {
            const float4* ps = reinterpret_cast<const float4*>(p.scale + (long long)b * p.c + kc * 64 + 8 * jp);
            const float4* ph = reinterpret_cast<const float4*>(p.shift + (long long)b * p.c + kc * 64 + 8 * jp);
            const float4 s0 = __ldg(ps), s1 = __ldg(ps + 1), h0 = __ldg(ph), h1 = __ldg(ph + 1);
            sc[kc][0] = s0.x; sc[kc][1] = s0.y; sc[kc][2] = s0.z; sc[kc][3] = s0.w;
            sc[kc][4] = s1.x; sc[kc][5] = s1.y; sc[kc][6] = s1.z; sc[kc][7] = s1.w;
            sh[kc][0] = h0.x; sh[kc][1] = h0.y; sh[kc][2] = h0.z; sh[kc][3] = h0.w;
            sh[kc][4] = h1.x; sh[kc][5] = h1.y; sh[kc][6] = h1.z; sh[kc][7] = h1.w;
          }
        }
        cur_b = b;
      }
      mbar_wait(&bars->full[stage], phase);
      const int px_valid = min(kHeadTileM, p.w - tw * kHeadTileM);   // pixels beyond the row are TMA zero fill: keep 0
#pragma unroll
      for (int kc = 0; kc < kHeadMaxKpt; ++kc) {
        if (kc >= p.kpt) break;
        uint8_t* a = smem_a + (size_t)stage * stage_bytes + (size_t)kc * kHeadTileM * 128;
        // (rows beyond px_valid hold TMA zero fill; transforming them too would turn the padding into `shift`, so the
        //  whole tile takes the unpredicated fast path only when it is full)
        uint4 q[kRows];
        if (px_valid == kHeadTileM) {
#pragma unroll
          for (int r = 0; r < kRows; ++r) {
            const int i = i0 + kRowStep * r;
            q[r] = *reinterpret_cast<const uint4*>(a + i * 128 + ((jp ^ (i & 7)) << 4));
          }
#pragma unroll
          for (int r = 0; r < kRows; ++r) {
            const int i = i0 + kRowStep * r;
            *reinterpret_cast<uint4*>(a + i * 128 + ((jp ^ (i & 7)) << 4)) = affine_act8(q[r], sc[kc], sh[kc], false, f16);
          }
        } else {
          for (int r = 0; r < kRows; ++r) {
            const int i = i0 + kRowStep * r;
            if (i < px_valid) {
              uint4* ptr = reinterpret_cast<uint4*>(a + i * 128 + ((jp ^ (i & 7)) << 4));
              *ptr = affine_act8(*ptr, sc[kc], sh[kc], false, f16);
            }
          }
        }
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars->ready[stage]);
      if (++stage == p.stages) { stage = 0; phase ^= 1u; }
    }
  } else {
    // ===================================================== epilogue warps: TMEM -> D ring, then shift-add of row j - 1
    const int quarter = warp & 3;                 // TMEM lane quarter this warp may read
    const int et = threadIdx.x - (64 + 32 * kHeadXfWarps);   // 0 .. 127
    const int row = quarter * 32 + lane;          // pixel of the tile
    const float b0 = __ldg(p.bias), b1 = __ldg(p.bias + 1), b2 = __ldg(p.bias + 2);
    const long long plane = (long long)p.h * p.w;
    // fused DDIM update: coefficients / noise source of the run's current step (the same values ddim_step_kernel reads)
    const bool ddim = p.ddim_x != nullptr;
    DdimCoef dk{};
    int dstep = 0;
    unsigned long long dseed = 0;
    const float* dnz = nullptr;
    if (ddim) {
      dstep = p.ddim_run->step;
      dseed = p.ddim_run->seed;
      dk = ddim_load_coef(p.ddim_coef, dstep);
      if (p.ddim_run->noise && dk.sigma > 0.f) dnz = p.ddim_run->noise + (long long)dstep * p.ddim_run->noise_step_stride;
    }
    const bool dstoch = dk.sigma > 0.f;
    // Tap row (r, s) of a D row is stored SHIFTED by 5 - s columns (source pixel q at index q + 5 - s), so the values output
    // pixel x needs from all 27 tap rows sit at the same 16-byte aligned index x + 4: four outputs per LDS.128.
    auto emit_row = [&](int g) {                  // output row g from D rows g-1, g, g+1 (all 128 epilogue threads)
      const int b = g / p.h, y = g - b * p.h;
      float* o = p.out + ((long long)b * 3) * plane + (long long)y * p.w;
      for (int x = 4 * et; x < p.w; x += 4 * 128) {
        float4 a0 = make_float4(b0, b0, b0, b0), a1 = make_float4(b1, b1, b1, b1), a2 = make_float4(b2, b2, b2, b2);
        // fused DDIM update: the three x loads are issued before the 27-tap shift-add so that their latency hides under it
        const long long off = (o - p.out) + x;
        float* xo = p.ddim_x + off;
        float4 x0 = a0, x1 = a0, x2 = a0;
        if (ddim) {
          x0 = *reinterpret_cast<const float4*>(xo);
          x1 = *reinterpret_cast<const float4*>(xo + plane);
          x2 = *reinterpret_cast<const float4*>(xo + 2 * plane);
        }
#pragma unroll
        for (int r = 0; r < 3; ++r) {
          const int yy = y + r - 1;
          if (yy < 0 || yy >= p.h) continue;      // zero padding of the normalised tensor
          const float* d = dring + (size_t)((g + r - 1) & (kHeadRing - 1)) * kHeadTaps * p.dpitch + x + 4;
#pragma unroll
          for (int s = 0; s < 3; ++s) {
            const float* t = d + (size_t)((r * 3 + s) * 3) * p.dpitch;
            const float4 v0 = *reinterpret_cast<const float4*>(t);
            const float4 v1 = *reinterpret_cast<const float4*>(t + p.dpitch);
            const float4 v2 = *reinterpret_cast<const float4*>(t + 2 * p.dpitch);
            a0.x += v0.x; a0.y += v0.y; a0.z += v0.z; a0.w += v0.w;
            a1.x += v1.x; a1.y += v1.y; a1.z += v1.z; a1.w += v1.w;
            a2.x += v2.x; a2.y += v2.y; a2.z += v2.z; a2.w += v2.w;
          }
        }
        *reinterpret_cast<float4*>(o + x) = a0;
        *reinterpret_cast<float4*>(o + plane + x) = a1;
        *reinterpret_cast<float4*>(o + 2 * plane + x) = a2;
        if (ddim) {   // x <- update(x, eps): element offsets / Philox counters exactly those of ddim_step_kernel
          *reinterpret_cast<float4*>(xo) = ddim_update4(x0, a0, dk, dstoch, dnz, off >> 2, dstep, dseed);
          *reinterpret_cast<float4*>(xo + plane) = ddim_update4(x1, a1, dk, dstoch, dnz, (off + plane) >> 2, dstep, dseed);
          *reinterpret_cast<float4*>(xo + 2 * plane) = ddim_update4(x2, a2, dk, dstoch, dnz, (off + 2 * plane) >> 2, dstep, dseed);
        }
      }
    };
    int it = 0;
    for (int j = j0; j <= j1; ++j) {
      float* drow = dring + (size_t)(j & (kHeadRing - 1)) * kHeadTaps * p.dpitch;
      for (int tw = 0; tw < p.tiles_w; ++tw, ++it) {
        const int as = it & 1;
        mbar_wait(&bars->tmem_full[as], (uint32_t)(it >> 1) & 1u);
        tc_fence_after();
        uint32_t r[32];
        const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(as * kHeadN);
        tmem_ld16(taddr, *reinterpret_cast<uint32_t(*)[16]>(&r[0]));
        tmem_ld16(taddr + 16u, *reinterpret_cast<uint32_t(*)[16]>(&r[16]));
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bars->tmem_empty[as]);
        const int px = tw * kHeadTileM + row;
        if (px < p.w) {
#pragma unroll
          for (int n = 0; n < kHeadTaps; ++n) drow[(size_t)n * p.dpitch + px + 5 - ((n / 3) % 3)] = __uint_as_float(r[n]);
        }
      }
      named_bar_sync(2, 128);                      // D row j complete (every warp wrote its pixel quarter of every tile)
      if (j - 1 >= g0 && j - 1 < g1) emit_row(j - 1);
    }
    if (j1 == g1 - 1) emit_row(g1 - 1);            // the band ends with the last row of the last image
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 1) tmem_dealloc(tmem_base, 64u);
}

// [img_ch=3][C][3][3] fp32 -> [32][C] 16-bit, row n = (r*3+s)*3 + co (rows 27..31 zero)
__global__ void pack_head_weight_kernel(const float* __restrict__ w, uint16_t* __restrict__ out, int c, int op_f16) {
  const int total = kHeadN * c;
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
    const int ci = idx % c, n = idx / c;
    float v = 0.f;
    if (n < kHeadTaps) {
      const int co = n % 3, tap = n / 3;
      v = w[((long long)co * c + ci) * 9 + tap];
    }
    out[idx] = to_op(v, op_f16 != 0);
  }
}

static int head_stage_count(int w, int c, int* smem_bytes) {
  const int kpt = c / 64;
  const int stage_bytes = kpt * kHeadTileM * 128;
  const int dpitch = w + 8;
  const int fixed = kpt * kHeadN * 128 + kHeadRing * kHeadTaps * dpitch * 4 + (int)sizeof(HeadBarriers) + 64;
  const int stages = std::min(kHeadMaxStages, (kHeadSmemBudget - fixed) / stage_bytes);
  if (smem_bytes) *smem_bytes = fixed + stages * stage_bytes;
  return stages;
}

bool head_conv_supported(int h, int w, int c, int cout) {
  (void)h;
  if (cout != 3 || c % 64 != 0 || c < 64 || c > 64 * kHeadMaxKpt || w < 8 || w % 4 != 0) return false;
  return head_stage_count(w, c, nullptr) >= 2;
}

int launch_head_conv(const void* x_op, const float* scale, const float* shift, const void* w_packed, const float* bias,
                     float* out_nchw, int batch, int h, int w, int c, int op_dtype, cudaStream_t stream, int max_stages,
                     float* ddim_x, const float* ddim_coef, const DdimRun* ddim_run) {
  CLPK_REQUIRE(head_conv_supported(h, w, c, 3), "fused head kernel unsupported for W=%d C=%d", w, c);
  static std::mutex mu;
  static bool attr_done[64] = {};
  int dev = 0;
  CLPK_CHECK_CUDA(cudaGetDevice(&dev));
  {
    std::lock_guard<std::mutex> lock(mu);
    if (dev >= 0 && dev < 64 && !attr_done[dev]) {
      CLPK_CHECK_CUDA(cudaFuncSetAttribute(head_conv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kHeadSmemBudget));
      attr_done[dev] = true;
    }
  }
  HeadParams p{};
  p.batch = batch; p.h = h; p.w = w; p.c = c;
  p.tiles_w = (w + kHeadTileM - 1) / kHeadTileM;
  p.kpt = c / 64;
  int smem_bytes = 0;
  p.stages = head_stage_count(w, c, &smem_bytes);
  if (max_stages >= 2 && max_stages < p.stages) p.stages = max_stages;  // (experiments: CLPK_HEAD_STAGES at plan creation)
  p.rows_total = batch * h;
  p.band = (p.rows_total + num_sms() - 1) / num_sms();
  p.op_f16 = (op_dtype == CLPK_OP_F16) ? 1 : 0;
  p.dpitch = w + 8;
  p.ddim_x = ddim_x; p.ddim_coef = ddim_coef; p.ddim_run = ddim_run;
  p.scale = scale; p.shift = shift; p.bias = bias; p.out = out_nchw;
  const CUtensorMapDataType dt = p.op_f16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
  CUtensorMap map_a, map_w;
  const long long C = c, W = w, H = h;
  cuuint64_t dims[5] = {(cuuint64_t)C, (cuuint64_t)W, 1, (cuuint64_t)H, (cuuint64_t)batch};
  cuuint64_t strides[4] = {(cuuint64_t)(C * 2), (cuuint64_t)(W * C * 2), (cuuint64_t)(W * C * 2), (cuuint64_t)(H * W * C * 2)};
  cuuint32_t box_a[5] = {64, (cuuint32_t)kHeadTileM, 1, 1, 1};
  CLPK_TRY_RC(encode_tensor_map(&map_a, x_op, 5, dims, strides, box_a, 128, dt));
  cuuint64_t wdims[2] = {(cuuint64_t)C, (cuuint64_t)kHeadN};
  cuuint64_t wstr[1] = {(cuuint64_t)(C * 2)};
  cuuint32_t box_w[2] = {64, (cuuint32_t)kHeadN};
  CLPK_TRY_RC(encode_tensor_map(&map_w, w_packed, 2, wdims, wstr, box_w, 128, dt));
  const int grid = (p.rows_total + p.band - 1) / p.band;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3(kHeadThreads);
  cfg.dynamicSmemBytes = (size_t)smem_bytes;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  CLPK_CHECK_CUDA(cudaLaunchKernelEx(&cfg, head_conv_kernel, map_a, map_w, p));
  CLPK_CHECK_LAUNCH();
  return CLPK_OK;
}

int launch_pack_head_weight(const float* w, void* out_op, int c, int op_dtype, cudaStream_t stream) {
  pack_head_weight_kernel<<<(kHeadN * c + 255) / 256, 256, 0, stream>>>(w, reinterpret_cast<uint16_t*>(out_op), c,
                                                                        op_dtype == CLPK_OP_F16);
  CLPK_CHECK_LAUNCH();
  return CLPK_OK;
}

}  // namespace clpk

using namespace clpk;

extern "C" int clpk_head_conv_supported(int h, int w, int c, int cout) {
  return (h > 0 && w > 0 && c > 0 && head_conv_supported(h, w, c, cout)) ? 1 : 0;
}

extern "C" int clpk_pack_head_weight(const float* w_dev, void* out_op_dev, int c, int op_dtype, void* stream) {
  CLPK_REQUIRE(w_dev && out_op_dev && c > 0 && (op_dtype == CLPK_OP_BF16 || op_dtype == CLPK_OP_F16),
               "clpk_pack_head_weight: bad arguments");
  return launch_pack_head_weight(w_dev, out_op_dev, c, op_dtype, (cudaStream_t)stream);
}

extern "C" int clpk_head_conv(const void* x_op, const float* in_scale, const float* in_shift, const void* w_packed,
                              const float* bias, float* out_nchw, int batch, int h, int w, int c, int op_dtype, void* stream) {
  CLPK_REQUIRE(x_op && in_scale && in_shift && w_packed && bias && out_nchw && batch > 0, "clpk_head_conv: bad arguments");
  CLPK_REQUIRE(op_dtype == CLPK_OP_BF16 || op_dtype == CLPK_OP_F16, "clpk_head_conv: operand dtype %d unknown", op_dtype);
  return launch_head_conv(x_op, in_scale, in_shift, w_packed, bias, out_nchw, batch, h, w, c, op_dtype, (cudaStream_t)stream, 0);
}
