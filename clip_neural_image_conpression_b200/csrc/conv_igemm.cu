// conv_igemm.cu — implicit-GEMM convolution on the Blackwell tensor cores (tcgen05.mma + TMEM + TMA), sm_100a.
//
// Replaces the cuDNN calls behind nn.Conv2d / nn.ConvTranspose2d of the reference
// (PKG/models/blocks.py:34,36 ResBlock convs; PKG/models/unet.py:63 stride-2 convs, :75 transposed convs, :79 out).
//
// GEMM view:  D[M = pixels, N = Cout] = A[M, K = taps*Cin] * W[N, K]^T, 16-bit operands (fp16 default, bf16
//             selectable — same kind::f16 rate), fp32 accumulation in TMEM.
//   * A is never materialised: an M tile is a (wbox x hbox) patch of ONE image, and the k-block of tap (r,s) is the
//     same patch shifted by (r-1, s-1).  One tiled TMA load over a 5-D view of the NHWC activation fetches it, with
//     the conv zero padding produced by TMA out-of-bounds fill.  Views (innermost first):
//        3x3 s1 / convT : [C, W, 1, H, B]
//        3x3 s2         : [2C, W/2, 2, H/2, B]   (x = wpar*C + c, p = hpar)  -> stride-2 taps become unit-stride boxes
//   * ConvTranspose2d(4,2,1) = 4 output phases, each a 2x2-tap stride-1 conv of the input (K = 4*Cin) whose result is
//     scattered to (2h+ph, 2w+pw).
//   * Row-slab mainloop (kSlab; 3x3 s1 convs on rows of >= 128 pixels): a stage = one (wbox+2)-pixel slab of input row
//     h+r-1 (64 channels) + the weight blocks of taps (r, 0..2); the three taps read the slab through smem descriptors
//     shifted by 0 / 128 / 256 bytes, so each A byte is fetched 3x instead of 9x.
//   * Warp roles (320 threads, 1 CTA/SM or one CTA pair per two SMs, persistent over tiles): warp0 = TMA producer,
//     warp1 = MMA issuer (+TMEM alloc), warps 2..9 = epilogue (two groups of four; every warp moves its own 32-row
//     sub-box: TMEM -> registers -> bias/FiLM/residual/statistics -> swizzled smem -> TMA store).  smem ring of `stages`
//     stages; TMEM accumulator double buffered so the epilogue of tile i overlaps the MMAs of tile i+1.
//   * Env switches read at setup (experiments; defaults are the measured best): CLPK_IGEMM_SLAB=0, CLPK_IGEMM_NCTA,
//     CLPK_IGEMM_BK, CLPK_IGEMM_SLOTS, CLPK_IGEMM_PREFETCH; debug builds (-DCLPK_IGEMM_DEBUG) add CLPK_IGEMM_DBG.
#include "conv_igemm.cuh"

#include <algorithm>
#include <mutex>
#include <stdlib.h>

namespace clpk {

#ifndef CLPK_EPI_GROUPS
#define CLPK_EPI_GROUPS 2
#endif
constexpr int kEpiGroups = CLPK_EPI_GROUPS;         // epilogue warp groups (4 warps each) alternating 32-column chunks
constexpr int kNumThreads = 64 + 128 * kEpiGroups;   // warp 0 TMA, warp 1 MMA, then the epilogue groups
constexpr int kXformThreads = 128;                    // + 4 transform warps in the kXform variants (input GroupNorm fused)
constexpr int kTileM = 128;
constexpr int kMaxStages = 8;
constexpr int kSmemBudget = 232448;  // 227 KB

struct __align__(8) PipeBarriers {
  uint64_t full[kMaxStages];
  uint64_t empty[kMaxStages];
  uint64_t ready[kMaxStages];   // kXform: the stage's A slab has been normalised in place (leader CTA's copy is waited on)
  uint64_t tmem_full[4];    // 2 accumulator buffers; 4 in the two-rows-per-item slab variant (2 rows x 2 buffers)
  uint64_t tmem_empty[4];
  uint64_t full_b[4];       // kRows2: weight-block ring (the A slabs use full / empty)
  uint64_t empty_b[4];
  uint64_t res_full[4 * kEpiGroups][6];  // per epilogue warp: residual sub-box landed in the warp's staging slot s
  uint32_t tmem_base;
};

struct TileCoord {
  int phase, b, h0, w0, n0;
  int th, tw;  // tile indices (h0 / hbox, w0 / wbox)
  bool ok;  // false: the odd CTA of a pair past the last spatial tile (loads hit TMA out-of-bounds zeros, nothing stored)
};

// Work item -> tile.  Order (fastest first): channel tile, transposed-conv phase, spatial tile (pair).  Keeping the 4
// phases of one spatial tile adjacent in time lets their interleaved output rows meet in L2 before they reach HBM.
__device__ __forceinline__ uint32_t fd_div(uint32_t n, const FastDiv& f) {
  return (uint32_t)(((unsigned long long)n * f.mul) >> f.shift);
}
// rows2 (two-rows-per-item slab variant): work items are pairs of vertically adjacent row tiles; `sub` selects the row
__device__ __forceinline__ TileCoord decode_tile(const IgemmParams& p, int item, int rank, int sub = 0) {
  TileCoord c;
  if (p.reverse) item = p.num_tiles - 1 - item;  // walk the tensor back to front (see IgemmParams::reverse)
  uint32_t r = fd_div((uint32_t)item, p.fd_tiles_n);
  const int nt = item - (int)r * p.n_tiles_n;
  const uint32_t q = fd_div(r, p.fd_phases);
  c.phase = (int)(r - q * (uint32_t)p.phases);
  const uint32_t msp = q * (uint32_t)p.ncta + (uint32_t)rank;
  c.ok = (int)msp < p.spatial_tiles;
  r = fd_div(msp, p.fd_tiles_w);
  const int tw = (int)(msp - r * (uint32_t)p.tiles_w);
  const uint32_t b = fd_div(r, p.fd_tiles_h);
  int th = (int)(r - b * p.fd_tiles_h.d);   // (divisor = row tiles, or row-tile PAIRS in rows2 mode)
  if (p.rows2) {  // fd_tiles_h divides by the number of row PAIRS
    th = 2 * th + sub;
    c.ok = c.ok && th < p.tiles_h;
  }
  c.b = (int)b;
  c.th = th;
  c.tw = tw;
  c.h0 = th * p.hbox;
  c.w0 = tw * p.wbox;
  c.n0 = nt * p.block_n;
  return c;
}

// per-tile epilogue vectors: y = acc * mul[n] + add[n]   (mul = 1 + FiLM scale or 1, add = bias * mul + FiLM shift)
constexpr int kMaxGroupChunks = 4;  // 32-column chunks one epilogue group handles per tile (block_n <= 256, 2 groups)
// Fused GroupNorm statistics are SHIFTED sums (robust against |mean| >> std, like torch's Welford pass): every epilogue
// warp takes K = the first value it sees of a group (lane 0's first channel, broadcast) and accumulates sum(x - K),
// sum((x - K)^2) over its 32 rows; the per-tile fold re-bases the 4 warps' sums on warp 0's K and emits the tile's
// (mean, M2 = sum((x - mean)^2), n) triple, which the consumer combines in fp64 (Chan et al.).
struct EpiVectors {       // laid out in smem as: float mul[block_n] | float add[block_n] | red | xf_tab
  float* mul;
  float* add;
  // [tile parity][epilogue group][chunk of the group][warp][group pair]: the warp's shifted (sum, sumsq), its shift K and
  // its number of valid rows — one 16-byte entry, so the per-tile fold reads one LDS.128 per warp
  float4* red;
  float* xf_tab;  // kXform: (scale[cin] | shift[cin]) of the current image's input transform
  int nchg;       // chunks per epilogue group = ceil(block_n / 32 / kEpiGroups)
  __device__ __forceinline__ int idx(int parity, int eg, int ci, int warp) const {
    return (((parity * kEpiGroups + eg) * nchg + ci) * 4 + warp) * 8;
  }
};
static __host__ __device__ inline int red_chunks_per_group(int block_n) { return (block_n / 32 + kEpiGroups - 1) / kEpiGroups; }
static __host__ __device__ inline int red_bytes(int block_n) {
  const int n = 2 * kEpiGroups * (red_chunks_per_group(block_n) > 0 ? red_chunks_per_group(block_n) : 1) * 4 * 8;
  return n * (int)sizeof(float4);
}
constexpr int kWarpSlotBytes = 32 * 128;  // per-warp staging slot: 32 tile rows x 128 B (fp32) or x 64 B (16-bit)

// perf-debug switches (env CLPK_IGEMM_DBG) are compiled in only with -DCLPK_IGEMM_DEBUG: their uniform branches sit in
// the hottest loops
#ifdef CLPK_IGEMM_DEBUG
#define CLPK_DBG(bits) (p.dbg & (bits))
// phase trace of ONE epilogue warp (CTA 0, group 0, quarter 0) and of the MMA issuer: clock64 stamps, read back with
// clpk_debug_trace (debug builds only)
__device__ long long g_trace[8192];
__device__ int g_trace_n;
#define CLPK_GTRACE(cond, tag)                                                       \
  do {                                                                               \
    if ((p.dbg & 128) && (cond)) {                                                   \
      long long _t;                                                                  \
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(_t));                         \
      const int _i = atomicAdd(&g_trace_n, 1);                                       \
      if (_i < 4096) { g_trace[2 * _i] = (tag); g_trace[2 * _i + 1] = _t; }          \
    }                                                                                \
  } while (0)
#define CLPK_TRACE(cond, tag)                                                        \
  do {                                                                               \
    if ((p.dbg & 64) && (cond)) {                                                    \
      const int _i = atomicAdd(&g_trace_n, 1);                                       \
      if (_i < 4096) { g_trace[2 * _i] = (tag); g_trace[2 * _i + 1] = clock64(); }   \
    }                                                                                \
  } while (0)
#else
#define CLPK_DBG(bits) 0
#define CLPK_TRACE(cond, tag) do { } while (0)
#define CLPK_GTRACE(cond, tag) do { } while (0)
#endif
static inline int epi_vector_bytes(int block_n, int xf_cin) {
  return 2 * 4 * block_n + (red_bytes(block_n) + 63) / 64 * 64 + (2 * 4 * xf_cin + 63) / 64 * 64;
}


// Fused GroupNorm statistics helpers.  A thread holds 32 consecutive output channels of one pixel; P = number of
// consumer-GroupNorm groups those 32 channels span (1, 2, 4 or 8).  row_sums: per-thread (sum, sumsq) per group —
// 32 FADD + 32 FFMA, all indices static.  warp_reduce_scatter: folds NV values over the 32 lanes with a halving butterfly
// (NV/2 + NV/4 + ... shuffles instead of 5*NV); afterwards lane L holds the total of value index
// bitreverse-ordered by the lane bits consumed, see `owner` below.  Fixed tree -> deterministic.
// redw: the warp's red entries of this chunk (one float4 per group pair); lane 0 parks the shift K and the warp's valid-row
// count right away (nothing but the row sums is carried across the chunk's store)
template <int P>
__device__ __forceinline__ void row_sums(const float (&v)[32], bool valid, float nrows, int lane, float (&out)[2 * P],
                                         float4* redw) {
  constexpr int per = 32 / P;  // >= 4: consecutive value pairs always belong to one group -> packed fp32x2 accumulators
#pragma unroll
  for (int i = 0; i < P; ++i) {
#ifdef CLPK_STATS_NOSHIFT   // experiment: cost of the shift itself
    const float k = 0.f;
#else
    const float k = __shfl_sync(0xffffffffu, v[i * per], 0);  // the warp's shift for this group (any finite value works)
#endif
    if (lane == 0) *reinterpret_cast<float2*>(reinterpret_cast<float*>(redw + i) + 2) = make_float2(k, nrows);
    const f32x2 kk = pack2(k, k);
    f32x2 s1 = pack2(0.f, 0.f), s2 = s1;
#pragma unroll
    for (int j = 0; j < per; j += 2) {
      const f32x2 d = sub2(pack2(v[i * per + j], v[i * per + j + 1]), kk);
      s1 = add2(s1, d);
      s2 = fma2(d, d, s2);
    }
    float a0, a1, b0, b1;
    unpack2(s1, a0, a1);
    unpack2(s2, b0, b1);
    out[2 * i] = valid ? a0 + a1 : 0.f;
    out[2 * i + 1] = valid ? b0 + b1 : 0.f;
  }
}

template <int NV>
__device__ __forceinline__ float warp_reduce_scatter(float (&vals)[NV], int lane) {
  // halving steps: after the step with offset `off`, each lane keeps NV/2 values (those of its half)
  int off = 16;
#pragma unroll
  for (int n = NV; n > 1; n >>= 1) {
    const int h = n >> 1;
    const bool upper = (lane & off) != 0;
#pragma unroll
    for (int i = 0; i < h; ++i) {
      const float keep = upper ? vals[h + i] : vals[i];
      const float send = upper ? vals[i] : vals[h + i];
      vals[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
    }
    off >>= 1;
  }
  float r = vals[0];
  for (; off > 0; off >>= 1) r += __shfl_xor_sync(0xffffffffu, r, off);
  return r;
}

// value index owned by `lane` after warp_reduce_scatter<NV>: the halving steps consumed lane bits 4,3,... (MSB first)
template <int NV>
__device__ __forceinline__ int scatter_owner_index(int lane) {
  int idx = 0, off = 16;
#pragma unroll
  for (int n = NV; n > 1; n >>= 1) {
    idx = idx * 2 + ((lane & off) ? 1 : 0);  // NOTE: upper half at each step = +h in the ORIGINAL numbering
    off >>= 1;
  }
  return idx;
}

// phase 1 (while the chunk's values are live): per-thread shifted row sums into sums[0 .. 2P)
template <int P>
__device__ __forceinline__ void gn_row_sums(const float (&v)[32], bool valid, float nrows, int lane, float (&sums)[16],
                                            float4* redw) {
  float vals[2 * P];
  row_sums<P>(v, valid, nrows, lane, vals, redw);
#pragma unroll
  for (int i = 0; i < 2 * P; ++i) sums[i] = vals[i];
}
// phase 2 (after the chunk's store has been issued): fold over the warp's 32 rows and park the warp's partials
template <int P>
__device__ __forceinline__ void gn_chunk_reduce(const float (&sums)[16], int lane, float4* redw) {
  float vals[2 * P];
#pragma unroll
  for (int i = 0; i < 2 * P; ++i) vals[i] = sums[i];
  const float r = warp_reduce_scatter<2 * P>(vals, lane);
  // Which original value does this lane hold?  At the first step the kept half is [h, 2h) for `upper` lanes, i.e. the
  // top bit of the original index = lane bit 4; the next step decides the next bit, and so on.
  constexpr int kSteps = (P == 1) ? 1 : (P == 2) ? 2 : (P == 4) ? 3 : 4;
  const int idx = scatter_owner_index<2 * P>(lane);
  const int low_mask = (32 >> kSteps) - 1;  // lanes differing only in the untouched low bits hold duplicates
  if ((lane & low_mask) == 0) reinterpret_cast<float*>(redw)[4 * (idx >> 1) + (idx & 1)] = r;  // .x = sum, .y = sumsq of pair idx/2
}

constexpr int kSlabABytes = 17 * 1024;  // (128 + 2) slab rows x 128 B = 16640, padded to the 1024-byte swizzle-atom pitch

template <int BLOCK_K, int NCTA, bool kSlab, bool kXform, bool kRows2 = false>
__global__ void __launch_bounds__(kNumThreads + (kXform ? kXformThreads : 0), 1)
conv_igemm_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_w,
                  const __grid_constant__ OutMaps maps_out, const __grid_constant__ OutMaps maps_res,
                  const __grid_constant__ IgemmParams p) {
  // BLOCK_K elements of K per pipeline stage.  A smem row holds one swizzle atom of kAtomK elements (<= 128 bytes);
  // BLOCK_K = 128 stages two atoms side by side (two TMA boxes per operand, 8 MMAs per barrier handshake).
  constexpr int kAtomK = BLOCK_K > 64 ? 64 : BLOCK_K;
  constexpr int kAtoms = BLOCK_K / kAtomK;
  constexpr int kSwizzle = kAtomK * 2;            // bytes per smem row
  constexpr int kAAtomBytes = kTileM * kAtomK * 2;
  constexpr int kABytes = kSlab ? kSlabABytes : kAAtomBytes * kAtoms;   // one A stage
  static_assert(!kSlab || BLOCK_K == 64, "slab mode stages 64-channel slabs");
  static_assert(!kXform || kSlab, "the input transform works on row slabs");
  static_assert(!kRows2 || (kSlab && !kXform && NCTA == 2), "two rows per item: CTA-pair slab mainloop only");
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  CLPK_GTRACE(threadIdx.x == 0 && (blockIdx.x == 0 || blockIdx.x == gridDim.x - 2), 300);  // kernel entry
  const uint32_t raw_addr = smem_u32(smem_raw);
  const uint32_t pad = ((raw_addr + 1023u) & ~1023u) - raw_addr;
  uint8_t* smem = smem_raw + pad;  // 1024-byte aligned (swizzle atoms)

  const int b_rows = p.block_n / NCTA;            // each CTA of a pair stages half of the B (weight) tile
  const int b_atom_bytes = b_rows * kAtomK * 2;
  const int b_bytes = kSlab ? 3 * b_atom_bytes : b_atom_bytes * kAtoms;  // slab mode: the 3 taps of one kernel row
  const uint32_t rank = (NCTA == 2) ? cluster_ctarank() : 0u;
  const int cluster_id = blockIdx.x / NCTA, num_clusters = gridDim.x / NCTA;
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + (size_t)p.stages * kABytes;
  uint8_t* smem_s = smem_b + (size_t)(kRows2 ? p.stages_b : p.stages) * b_bytes;  // staging buffers (1024-aligned sizes)
  EpiVectors vec_s;
  vec_s.mul = reinterpret_cast<float*>(smem_s + (size_t)p.n_staging * 4 * p.slot_bytes);
  vec_s.add = vec_s.mul + p.block_n;
  vec_s.nchg = red_chunks_per_group(p.block_n) > 0 ? red_chunks_per_group(p.block_n) : 1;
  vec_s.red = reinterpret_cast<float4*>(vec_s.add + p.block_n);
  vec_s.xf_tab = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(vec_s.red) + (red_bytes(p.block_n) + 63) / 64 * 64);
  const EpiVectors* vec = &vec_s;
  PipeBarriers* bars = reinterpret_cast<PipeBarriers*>(reinterpret_cast<uint8_t*>(vec_s.xf_tab) +
                                                       (kXform ? (2 * 4 * p.cin + 63) / 64 * 64 : 0));
  if (p.smem_slack == 0 && pad != 0) {  // the launch reserved no alignment slack: the 1024-byte alignment must hold
    if (threadIdx.x == 0) printf("clpk: dynamic shared memory is not 1024-byte aligned (0x%x)\n", raw_addr);
    __trap();
  }

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int k_blocks = kSlab ? 3 * p.kpt : p.taps * p.kpt;  // pipeline stages per tile

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_a);
    tma_prefetch_desc(&map_w);
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(&bars->full[s], 1);
      mbar_init(&bars->empty[s], 1);
      mbar_init(&bars->ready[s], 4 * NCTA);  // one arrive per transform warp (of both CTAs of a pair)
    }
    for (int s = 0; s < 4; ++s) {
      mbar_init(&bars->tmem_full[s], 1);
      mbar_init(&bars->tmem_empty[s], 4 * kEpiGroups * NCTA);  // one arrive per epilogue warp (of both CTAs of a pair)
      mbar_init(&bars->full_b[s], 1);
      mbar_init(&bars->empty_b[s], 1);
    }
    for (int g = 0; g < 4 * kEpiGroups; ++g)
      for (int s = 0; s < 6; ++s) mbar_init(&bars->res_full[g][s], 1);
    fence_mbar_init();
  }
  if (warp == 1) {
    if (NCTA == 2) { tmem_alloc_pair(&bars->tmem_base, (uint32_t)p.tmem_cols); tmem_relinquish_pair(); }
    else { tmem_alloc(&bars->tmem_base, (uint32_t)p.tmem_cols); tmem_relinquish(); }
  }
  tc_fence_before();
  if (NCTA == 2) cluster_sync_all(); else __syncthreads();  // barriers of BOTH CTAs initialised before any remote arrive
  tc_fence_after();
  const uint32_t tmem_base = bars->tmem_base;
  pdl_prologue_done();  // everything above overlapped the previous kernel's tail; its outputs are visible from here on
  CLPK_GTRACE(threadIdx.x == 0 && (blockIdx.x == 0 || blockIdx.x == gridDim.x - 2), 301);  // setup done

  if (kRows2 && warp == 0) {
    // ===================================================== TMA producer, two output rows per item (kRows2)
    // Item = row tiles (h0, h0 + 1) of one 128-pixel column segment, two accumulators.  Per 64-channel block the four
    // slabs of input rows h0-1 .. h0+2 go through the A ring (slab i feeds row h0 with kernel row i and row h0+1 with
    // kernel row i-1) and the three weight blocks B_r (taps (r, 0..2)) through their own ring: every slab is staged once
    // for two output rows (4 instead of 6 slab fills) and every weight block once for both rows (3 instead of 6) —
    // 140 KB instead of 249 KB of operand fill per 128x128 output tile.  (Built to test whether the mainloop is bound by
    // operand-fill bytes: it is not — see igemm_setup — so this variant is opt-in.)
    // Issue order = consumption order: slab0 slab1 B0 slab2 B1 slab3 B2.
    const uint32_t slab_tx = (uint32_t)(p.wbox + 2) * 128u, b_tx = (uint32_t)b_bytes;
    int sa = 0, sb = 0;
    uint32_t pa = 0, pb = 0;
    for (int item = cluster_id; item < p.num_tiles; item += num_clusters) {
      const TileCoord tc = decode_tile(p, item, (int)rank, 0);
      const int w_row = tc.n0 + (int)rank * b_rows;
      for (int kc = 0; kc < p.kpt; ++kc) {
#pragma unroll
        for (int step = 0; step < 7; ++step) {
          const bool is_b = (step == 2 || step == 4 || step == 6);
          const int idx = is_b ? (step - 2) / 2 : (step < 2 ? step : (step + 1) / 2);   // B_r index or slab index
          if (!is_b) {
            mbar_wait(&bars->empty[sa], pa ^ 1u);
            if (elect_one()) {
              const uint32_t lead_full = mapa_u32(smem_u32(&bars->full[sa]), 0u);
              if (rank == 0) mbar_arrive_expect_tx(&bars->full[sa], 2u * slab_tx);
              tma_load_5d_pair(smem_a + (size_t)sa * kABytes, &map_a, lead_full, kc * 64, tc.w0 - 1, 0, tc.h0 + idx - 1, tc.b);
            }
            __syncwarp();
            if (++sa == p.stages) { sa = 0; pa ^= 1u; }
          } else {
            mbar_wait(&bars->empty_b[sb], pb ^ 1u);
            if (elect_one()) {
              const uint32_t lead_full = mapa_u32(smem_u32(&bars->full_b[sb]), 0u);
              if (rank == 0) mbar_arrive_expect_tx(&bars->full_b[sb], 2u * b_tx);
#pragma unroll
              for (int s3 = 0; s3 < 3; ++s3)
                tma_load_2d_pair(smem_b + (size_t)sb * b_bytes + s3 * b_atom_bytes, &map_w, lead_full,
                                 (idx * 3 + s3) * p.cin + kc * 64, w_row);
            }
            __syncwarp();
            if (++sb == p.stages_b) { sb = 0; pb ^= 1u; }
          }
        }
      }
    }
  } else if (kRows2 && warp == 1) {
    // ===================================================== MMA issuer, two output rows per item (leader CTA issues)
    if (rank == 0) {
      const uint32_t idesc = make_idesc_16bit(kTileM * NCTA, p.block_n, p.op_f16 != 0);
      const uint64_t adesc0 = make_kmajor_desc<kSwizzle>(smem_u32(smem_a));
      const uint64_t bdesc0 = make_kmajor_desc<kSwizzle>(smem_u32(smem_b));
      const uint64_t a_step = (uint64_t)(kABytes >> 4), b_step = (uint64_t)(b_bytes >> 4);
      int sa = 0, sb = 0;      // ring slots of the NEXT slab / weight block to be waited for
      uint32_t pa = 0, pb = 0;
      int it = 0;
      for (int item = cluster_id; item < p.num_tiles; item += num_clusters, ++it) {
        const int as0 = (2 * it) & 3;                       // accumulators of rows h0 and h0 + 1
        const uint32_t aphase = (uint32_t)(it >> 1) & 1u;   // buffer index (2 it + sub) & 3 is reused every 2 items
        mbar_wait(&bars->tmem_empty[as0], aphase ^ 1u);
        mbar_wait(&bars->tmem_empty[as0 + 1], aphase ^ 1u);
        tc_fence_after();
        const uint32_t tmem_d0 = tmem_base + (uint32_t)(as0 * p.block_n), tmem_d1 = tmem_d0 + (uint32_t)p.block_n;
        for (int kc = 0; kc < p.kpt; ++kc) {
          int slot[4];         // ring slots of this channel block's four slabs
#pragma unroll
          for (int r = 0; r < 3; ++r) {
            // slabs first needed at this step: 0 and 1 at r = 0, then r + 1
#pragma unroll
            for (int i = (r == 0 ? 0 : r + 1); i <= r + 1; ++i) {
              mbar_wait(&bars->full[sa], pa);
              slot[i] = sa;
              if (++sa == p.stages) { sa = 0; pa ^= 1u; }
            }
            mbar_wait(&bars->full_b[sb], pb);
            tc_fence_after();
            if (elect_one()) {
              const uint64_t bdesc = bdesc0 + b_step * (uint64_t)sb;
              const uint64_t ad0 = adesc0 + a_step * (uint64_t)slot[r], ad1 = adesc0 + a_step * (uint64_t)slot[r + 1];
#pragma unroll
              for (int s3 = 0; s3 < 3; ++s3) {
#pragma unroll
                for (int kk = 0; kk < 4; ++kk) {
                  const uint64_t ao = (uint64_t)(s3 * 8 + 2 * kk);
                  const uint64_t bo = (uint64_t)(s3 * (b_atom_bytes >> 4) + 2 * kk);
                  const uint32_t acc = (s3 == 0 && kk == 0 && r == 0) ? (uint32_t)(kc != 0) : 1u;
                  umma_f16kind_pair(tmem_d0, ad0 + ao, bdesc + bo, idesc, acc);
                  umma_f16kind_pair(tmem_d1, ad1 + ao, bdesc + bo, idesc, acc);
                }
              }
              // the weight block and slab r are done (slab r + 1 is used once more, by row h0 at the next step)
              umma_commit_pair(&bars->empty_b[sb]);
              umma_commit_pair(&bars->empty[slot[r]]);
              if (r == 2) {
                umma_commit_pair(&bars->empty[slot[3]]);
                if (kc == p.kpt - 1) {
                  umma_commit_pair(&bars->tmem_full[as0]);
                  umma_commit_pair(&bars->tmem_full[as0 + 1]);
                }
              }
            }
            __syncwarp();
            if (++sb == p.stages_b) { sb = 0; pb ^= 1u; }
          }
        }
      }
    }
  } else if (warp == 0) {
    // ===================================================== TMA producer
    // The whole warp walks the loop converged (every lane waits on the barrier); only the issue of the uniform-datapath
    // TMA instructions sits under elect.sync.  Issuing them from a divergent `if (lane == 0)` region makes the compiler
    // wrap each one in an ELECT/BRA.U.ANY serialisation loop, which throttles the pipeline.
    {
      const uint32_t rows = kSlab ? (uint32_t)(p.wbox + 2) : (uint32_t)(p.wbox * p.hbox);
      const uint32_t tx_bytes = rows * BLOCK_K * 2 + (uint32_t)b_bytes;
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = cluster_id; tile < p.num_tiles; tile += num_clusters) {
        const TileCoord tc = decode_tile(p, tile, (int)rank);
        const int w_row = tc.phase * p.cout_pad + tc.n0 + (int)rank * b_rows;
        if ((p.ep.resid || p.ep.resid_op) && p.chunked && tc.ok && elect_one()) {
          // the epilogue will add this residual tile: pull it into L2 while the MMAs run (the residual map's box is one
          // epilogue warp's 32-row sub-box)
          for (int q = 0; q < p.res_ahead; ++q) {
            const int qh = (q * 32) / p.wbox, qw = q * 32 - qh * p.wbox;
            if (qh >= p.hbox) break;
            for (int c = 0; c < p.block_n; c += 32)
              tma_prefetch_5d(&maps_res.m[tc.phase], tc.n0 + c, tc.w0 + qw, 0, tc.h0 + qh, tc.b);
          }
        }
        __syncwarp();
        int tap = 0, kc = 0;
        for (int kb = 0; kb < k_blocks; ++kb) {
          const int ti = tc.phase * p.taps + tap;
          mbar_wait(&bars->empty[stage], phase ^ 1u);
          if (elect_one()) {
            uint32_t tx = tx_bytes;
            if (CLPK_DBG(4)) tx -= rows * BLOCK_K * 2;
            if (CLPK_DBG(8)) tx -= (uint32_t)b_bytes;
            uint8_t* a_dst = smem_a + (size_t)stage * kABytes;
            uint8_t* b_dst = smem_b + (size_t)stage * b_bytes;
            if (kSlab && kXform) {
              // input transform on: every CTA signals its OWN full barrier (its transform warps wait there, locally); the
              // MMA issuer waits on `ready`, which those warps arrive on after normalising the slab
              mbar_arrive_expect_tx(&bars->full[stage], tx);
              tma_load_5d(a_dst, &map_a, &bars->full[stage], kc * 64, tc.w0 - 1, 0, tc.h0 + tap - 1, tc.b);
#pragma unroll
              for (int s3 = 0; s3 < 3; ++s3) {
                const int kcol = (tap * 3 + s3) * p.cin + kc * 64;
                tma_load_2d(b_dst + s3 * b_atom_bytes, &map_w, &bars->full[stage], kcol, w_row);
              }
            } else if (kSlab) {
              // stage kb = (kernel row r = tap, channel block kc): slab = input row h0 + r - 1, pixels w0 - 1 .. w0 + wbox
              // (TMA zero-fills what lies outside the image), weights = taps (r, 0..2) of that channel block
              const uint32_t lead_full = (NCTA == 2) ? mapa_u32(smem_u32(&bars->full[stage]), 0u) : smem_u32(&bars->full[stage]);
              if (NCTA == 1 || rank == 0) mbar_arrive_expect_tx(&bars->full[stage], (uint32_t)NCTA * tx);
              if (!(CLPK_DBG(4))) {
                if (NCTA == 2) tma_load_5d_pair(a_dst, &map_a, lead_full, kc * 64, tc.w0 - 1, 0, tc.h0 + tap - 1, tc.b);
                else tma_load_5d(a_dst, &map_a, &bars->full[stage], kc * 64, tc.w0 - 1, 0, tc.h0 + tap - 1, tc.b);
              }
              if (!(CLPK_DBG(8))) {
#pragma unroll
                for (int s3 = 0; s3 < 3; ++s3) {
                  const int kcol = (tap * 3 + s3) * p.cin + kc * 64;
                  if (NCTA == 2) tma_load_2d_pair(b_dst + s3 * b_atom_bytes, &map_w, lead_full, kcol, w_row);
                  else tma_load_2d(b_dst + s3 * b_atom_bytes, &map_w, &bars->full[stage], kcol, w_row);
                }
              }
            } else if (NCTA == 2) {
              // both CTAs' bytes are accounted on the LEADER's full barrier (the MMA issuer lives there)
              if (rank == 0) mbar_arrive_expect_tx(&bars->full[stage], 2u * tx);
              const uint32_t lead_full = mapa_u32(smem_u32(&bars->full[stage]), 0u);
#pragma unroll
              for (int at = 0; at < kAtoms; ++at) {
                if (!(CLPK_DBG(4)))
                  tma_load_5d_pair(a_dst + at * kAAtomBytes, &map_a, lead_full, p.tap_x[ti] + kc * BLOCK_K + at * kAtomK,
                                   tc.w0 + p.tap_dw[ti], p.tap_p[ti], tc.h0 + p.tap_dh[ti], tc.b);
                if (!(CLPK_DBG(8)))
                  tma_load_2d_pair(b_dst + at * b_atom_bytes, &map_w, lead_full, kb * BLOCK_K + at * kAtomK, w_row);
              }
            } else {
              mbar_arrive_expect_tx(&bars->full[stage], tx);
#pragma unroll
              for (int at = 0; at < kAtoms; ++at) {
                if (!(CLPK_DBG(4)))
                  tma_load_5d(a_dst + at * kAAtomBytes, &map_a, &bars->full[stage], p.tap_x[ti] + kc * BLOCK_K + at * kAtomK,
                              tc.w0 + p.tap_dw[ti], p.tap_p[ti], tc.h0 + p.tap_dh[ti], tc.b);
                if (!(CLPK_DBG(8)))
                  tma_load_2d(b_dst + at * b_atom_bytes, &map_w, &bars->full[stage], kb * BLOCK_K + at * kAtomK, w_row);
              }
            }
          }
          __syncwarp();
          if (++kc == p.kpt) { kc = 0; ++tap; }
          if (++stage == p.stages) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================================================== MMA issuer (warp converged, one elected lane issues)
    if (rank == 0) {  // in a CTA pair only the leader issues (for both CTAs)
      const uint32_t idesc = make_idesc_16bit(kTileM * NCTA, p.block_n, p.op_f16 != 0);
      const uint64_t adesc0 = make_kmajor_desc<kSwizzle>(smem_u32(smem_a));
      const uint64_t bdesc0 = make_kmajor_desc<kSwizzle>(smem_u32(smem_b));
      const uint64_t a_step = (uint64_t)(kABytes >> 4), b_step = (uint64_t)(b_bytes >> 4);  // per stage, in 16-byte units
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int tile = cluster_id; tile < p.num_tiles; tile += num_clusters, ++it) {
        const int as = it & (p.nacc - 1);                   // accumulator ring of p.nacc (2 or 4) TMEM buffers
        const uint32_t aphase = (uint32_t)(it >> p.nacc_shift) & 1u;
        CLPK_TRACE(blockIdx.x == 0 && lane == 0, 200);
        mbar_wait(&bars->tmem_empty[as], aphase ^ 1u);  // epilogue(s) have drained this accumulator
        tc_fence_after();
        CLPK_TRACE(blockIdx.x == 0 && lane == 0, 201);
        const uint32_t tmem_d = tmem_base + (uint32_t)(as * p.block_n);
        for (int kb = 0; kb < k_blocks; ++kb) {
          mbar_wait(kXform ? &bars->ready[stage] : &bars->full[stage], phase);  // operands (of both CTAs) are in place
          tc_fence_after();
          CLPK_GTRACE(lane == 0 && it == 0 && kb == 0 && (blockIdx.x == 0 || blockIdx.x == gridDim.x - 2), 302);  // first operands landed
          if (elect_one()) {
            const uint64_t adesc = adesc0 + a_step * (uint64_t)stage;
            const uint64_t bdesc = bdesc0 + b_step * (uint64_t)stage;
            if (kSlab) {
              if (!(CLPK_DBG(2))) {
#pragma unroll
                for (int s3 = 0; s3 < 3; ++s3) {
#pragma unroll
                  for (int kk = 0; kk < 4; ++kk) {
                    // tap (r, s3): the tile's 128 pixels are slab rows s3 .. s3 + 127 -> start address + s3 * 128 B (the
                    // 128B swizzle is a function of the absolute smem address, so a row-shifted start stays consistent
                    // with what the TMA wrote); K advance as usual (+32 B = +2 in the addr >> 4 field)
                    const uint64_t ao = (uint64_t)(s3 * 8 + 2 * kk);
                    const uint64_t bo = (uint64_t)(s3 * (b_atom_bytes >> 4) + 2 * kk);
                    const uint32_t acc = (s3 == 0 && kk == 0) ? (uint32_t)(kb != 0) : 1u;
                    if (NCTA == 2) umma_f16kind_pair(tmem_d, adesc + ao, bdesc + bo, idesc, acc);
                    else umma_f16kind(tmem_d, adesc + ao, bdesc + bo, idesc, acc);
                  }
                }
              }
            } else if (!(CLPK_DBG(2))) {
#pragma unroll
              for (int k = 0; k < BLOCK_K / 16; ++k) {
                // K advance: 16 elements = 32 bytes inside the swizzled row (+2 in the addr>>4 field); next atom = next
                // 128-row block of the stage
                const int at = k / (kAtomK / 16), kk = k % (kAtomK / 16);
                const uint64_t ao = (uint64_t)(at * (kAAtomBytes >> 4) + 2 * kk);
                const uint64_t bo = (uint64_t)(at * (b_atom_bytes >> 4) + 2 * kk);
                const uint32_t acc = (k == 0) ? (uint32_t)(kb != 0) : 1u;
                if (NCTA == 2) umma_f16kind_pair(tmem_d, adesc + ao, bdesc + bo, idesc, acc);
                else umma_f16kind(tmem_d, adesc + ao, bdesc + bo, idesc, acc);
              }
            }
            // smem slot reusable (in both CTAs) once these MMAs retire; last k-block: accumulator complete
            if (NCTA == 2) umma_commit_pair(&bars->empty[stage]); else umma_commit(&bars->empty[stage]);
            if (kb == k_blocks - 1) {
              if (NCTA == 2) umma_commit_pair(&bars->tmem_full[as]); else umma_commit(&bars->tmem_full[as]);
              CLPK_GTRACE((blockIdx.x == 0 || blockIdx.x == gridDim.x - 2), 303);  // a tile's MMAs issued (last one = mainloop end)
            }
          }
          __syncwarp();
          if (++stage == p.stages) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (kXform && warp >= 2 + 4 * kEpiGroups) {
    // ===================================================== transform warps (kXform): GroupNorm apply [+ SiLU] of the
    // consumer side, in place on the freshly landed A slab: a <- act(a * scale[b, c] + shift[b, c]) (fp32 math, 16-bit
    // in/out), i.e. blocks.py:41,43 / unet.py:105 without the stand-alone normalisation pass over HBM.  Pixels outside
    // the image (TMA zero fill = the conv's padding of the NORMALISED tensor) are left untouched.
    // Thread (row i0 + 16 k, 16-byte piece j): the 128B TMA swizzle puts logical piece j of slab row i at physical piece
    // j ^ (i & 7) (stages are 1024-byte aligned) -> a quarter warp covers one 128-byte row: conflict-free.
    const int xt = threadIdx.x - kNumThreads;
    const int j = xt & 7, i0 = xt >> 3;
    const bool f16 = p.op_f16 != 0, act = p.ep.in_silu != 0;
    const bool h2 = f16 && act && p.xform_h2 != 0;  // SiLU on packed halves (one MUFU op per pair)
    float* tab = vec->xf_tab;
    const uint32_t ready0 = (NCTA == 2) ? mapa_u32(smem_u32(&bars->ready[0]), 0u) : 0u;
    int cur_b = -1, stage = 0;
    uint32_t phase = 0;
    for (int tile = cluster_id; tile < p.num_tiles; tile += num_clusters) {
      const TileCoord tc = decode_tile(p, tile, (int)rank);
      const int vb = tc.ok ? tc.b : 0;
      if (vb != cur_b) {  // new image: refresh the (scale | shift) table — every ~(tiles per image / CTAs) tiles
        named_bar_sync(kEpiGroups + 2, kXformThreads);
        for (int c = xt; c < p.cin; c += kXformThreads) {
          tab[c] = __ldg(p.ep.in_scale + (long long)vb * p.cin + c);
          tab[p.cin + c] = __ldg(p.ep.in_shift + (long long)vb * p.cin + c);
        }
        named_bar_sync(kEpiGroups + 2, kXformThreads);
        cur_b = vb;
      }
      int tap = 0, kc = 0;
      for (int kb = 0; kb < k_blocks; ++kb) {
        mbar_wait(&bars->full[stage], phase);  // this CTA's slab (and weight blocks) have landed
        const int hin = tc.h0 + tap - 1;
        if (tc.ok && hin >= 0 && hin < p.grid_h) {
          float sc[8], sh[8];
          const float4 s0 = *reinterpret_cast<const float4*>(tab + kc * 64 + 8 * j);
          const float4 s1 = *reinterpret_cast<const float4*>(tab + kc * 64 + 8 * j + 4);
          const float4 h0 = *reinterpret_cast<const float4*>(tab + p.cin + kc * 64 + 8 * j);
          const float4 h1 = *reinterpret_cast<const float4*>(tab + p.cin + kc * 64 + 8 * j + 4);
          sc[0] = s0.x; sc[1] = s0.y; sc[2] = s0.z; sc[3] = s0.w; sc[4] = s1.x; sc[5] = s1.y; sc[6] = s1.z; sc[7] = s1.w;
          sh[0] = h0.x; sh[1] = h0.y; sh[2] = h0.z; sh[3] = h0.w; sh[4] = h1.x; sh[5] = h1.y; sh[6] = h1.z; sh[7] = h1.w;
          uint8_t* a = smem_a + (size_t)stage * kABytes;
          // all of the thread's rows are loaded before the first one is transformed: 9 independent LDS -> math -> STS
          // chains in flight (a serial row loop is latency-bound at ~2x the stage's MMA time)
          constexpr int kRowsPerThread = (kTileM + 2 + kXformThreads / 8 - 1) / (kXformThreads / 8);
          uint4 q[kRowsPerThread];
#pragma unroll
          for (int r = 0; r < kRowsPerThread; ++r) {
            const int i = i0 + r * (kXformThreads / 8);
            const int w = tc.w0 - 1 + i;
            if (i < kTileM + 2 && w >= 0 && w < p.grid_w)
              q[r] = *reinterpret_cast<const uint4*>(a + i * 128 + ((j ^ (i & 7)) << 4));
          }
#pragma unroll
          for (int r = 0; r < kRowsPerThread; ++r) {
            const int i = i0 + r * (kXformThreads / 8);
            const int w = tc.w0 - 1 + i;
            if (i < kTileM + 2 && w >= 0 && w < p.grid_w)
              *reinterpret_cast<uint4*>(a + i * 128 + ((j ^ (i & 7)) << 4)) =
                  h2 ? affine_silu8_h2(q[r], sc, sh) : affine_act8(q[r], sc, sh, act, f16);
          }
        }
        fence_proxy_async_smem();  // generic-proxy writes -> visible to the tensor core's async-proxy reads
        __syncwarp();
        if (lane == 0) {
          if (NCTA == 2) mbar_arrive_release_cluster(ready0 + (uint32_t)stage * 8u);
          else mbar_arrive(&bars->ready[stage]);
        }
        if (++kc == p.kpt) { kc = 0; ++tap; }
        if (++stage == p.stages) { stage = 0; phase ^= 1u; }
      }
    }
  } else {
    // ===================================================== epilogue warps (TMEM lane quarter = warp % 4)
    // kEpiGroups groups of 4 warps; group eg owns the 32-column chunks with (chunk index % kEpiGroups) == eg.  Every warp
    // is autonomous inside a tile: it owns the 32 tile rows of its TMEM lane quarter, its own staging slots (32 rows x
    // 128 B), its own residual-load mbarriers and issues its own TMA loads / stores for that 32-row sub-box — no
    // cross-warp barrier per chunk.  The only group-wide synchronisation is ONE named barrier per tile for the 4-warp
    // fold of the fused GroupNorm partial sums.
    const int eg = (warp - 2) >> 2;          // epilogue group
    const int quarter = warp & 3;
    const int ew = eg * 4 + quarter;         // epilogue warp index
    const int row = quarter * 32 + lane;
    const int hl = row / p.wbox;
    const int wl = row - hl * p.wbox;
    const int sub_h = (quarter * 32) / p.wbox, sub_w = (quarter * 32) - sub_h * p.wbox;  // this warp's sub-box origin
    const clpk_conv_epilogue& ep = p.ep;
    const int ldc = ep.cout_valid;
    const int etid = threadIdx.x - 64;       // 0 .. 128*kEpiGroups-1 among all epilogue threads
    const int gtid = etid & 127;             // within the group
    const bool f16 = p.op_f16 != 0;
    const int S = p.n_staging / kEpiGroups;  // staging slots of this warp (a 16 KB staging buffer = 4 warp slots)
    const int slot_bytes = p.slot_bytes;     // per-warp slot: 32 rows x 128 B (fp32 sub-box) or x 64 B (all-16-bit epilogue)
    uint8_t* wslots = smem_s + (size_t)ew * S * slot_bytes;
    uint64_t* res_full = bars->res_full[ew];
    const int bar_id = 1 + eg;
    int it = 0;
    uint32_t slot = 0, sphase = 0;           // staging slot of the current chunk and how often the ring has wrapped (parity)
    bool slot_pending = false;               // a store was committed whose slot-reuse bookkeeping is still to be done
    int cached_key = -1;
    // Residual sub-boxes are TMA-loaded straight into the staging slot that is later stored from (read-modify-write in
    // place).  Lane 0 keeps `A` of them in flight; (ld_tile, ld_c, ld_slot) is its load cursor.
    const bool res16 = p.chunked && (ep.resid_op != nullptr);  // 16-bit residual (operand format): 64-byte rows, 64B swizzle
    const bool res_tma = p.chunked && (ep.resid != nullptr || res16);
    const bool st_f32 = p.chunked && (ep.out_f32 != nullptr);  // fp32 output goes through the staging slots + TMA store
    // a 16-bit-only output (ResBlock conv1) is staged too (64-byte rows, 64B swizzle): 4x fewer L1/smem wavefronts than
    // 32 lanes x 16 B scattered over 32 lines per store instruction
    const bool st_16 = p.chunked && (ep.out_f32 == nullptr) && (ep.out_op != nullptr) && p.n_staging > 0;
    const bool st_tma = st_f32 || st_16;
    const int A = (S >= 2) ? S - 1 : 1;      // residual look-ahead (chunks of this warp)
    const int cstep = 32 * kEpiGroups;
    const int cpg = ep.gn_cpg;
    const int npairs = cpg >= 32 ? 1 : (cpg > 0 ? 32 / cpg : 0);
    const int npairs_shift = npairs >= 8 ? 3 : npairs >= 4 ? 2 : npairs >= 2 ? 1 : 0;
    int ld_tile = cluster_id, ld_c = 32 * eg, ld_sub = 0;
    uint32_t ld_slot = 0;
    int lc_tile = -1, lc_sub = -1;
    TileCoord lc{};
    auto issue_res_load = [&]() {  // lane 0 only
      while (ld_tile < p.num_tiles && ld_c >= p.block_n) {
        ld_c = 32 * eg;
        if (kRows2 && ld_sub == 0) ld_sub = 1; else { ld_sub = 0; ld_tile += num_clusters; }
      }
      if (ld_tile >= p.num_tiles) return;
      if (ld_tile != lc_tile || ld_sub != lc_sub) { lc = decode_tile(p, ld_tile, (int)rank, ld_sub); lc_tile = ld_tile; lc_sub = ld_sub; }
      mbar_arrive_expect_tx(&res_full[ld_slot], (uint32_t)(res16 ? 32 * 64 : kWarpSlotBytes));
      tma_load_5d(wslots + (size_t)ld_slot * slot_bytes, &maps_res.m[lc.phase], &res_full[ld_slot], lc.n0 + ld_c,
                  lc.w0 + sub_w, 0, lc.h0 + sub_h, lc.b);
      if (++ld_slot == (uint32_t)S) ld_slot = 0;
      ld_c += cstep;
    };
    if (res_tma && lane == 0 && !(CLPK_DBG(1)))
      for (int i = 0; i < A; ++i) issue_res_load();
    const uint32_t lead_tmem_empty0 = (NCTA == 2) ? mapa_u32(smem_u32(&bars->tmem_empty[0]), 0u) : 0u;
    constexpr int kSub = kRows2 ? 2 : 1;     // row tiles per work item; TMEM holds 2 * kSub accumulators
    for (int tile = cluster_id; tile < p.num_tiles; tile += num_clusters)
    for (int sub = 0; sub < kSub; ++sub, ++it) {
      const int as = kRows2 ? (it & 3) : (it & (p.nacc - 1));
      const uint32_t aphase = (uint32_t)(it >> (kRows2 ? 2 : p.nacc_shift)) & 1u;
      const TileCoord tc = decode_tile(p, tile, (int)rank, sub);
      const int h = tc.h0 + hl, w = tc.w0 + wl;
      const bool valid = tc.ok && (hl < p.hbox) && (h < p.grid_h) && (w < p.grid_w);
      const int oh = h * p.out_scale + (tc.phase >> 1), ow = w * p.out_scale + (tc.phase & 1);
      const long long opix = ((long long)tc.b * p.out_h + oh) * p.out_w + ow;
      const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(as * p.block_n);

      if (p.chunked) {
        // ------------------------------------------------ chunked path: TMEM -> regs -> swizzled smem -> TMA store
        const int vb = tc.ok ? tc.b : 0;  // a masked tile still runs the (uniform) protocol, on image 0's vectors
        const int key = vb * 4096 + tc.n0;
        if (key != cached_key) {  // (image, channel-tile) changed: refresh the folded bias / FiLM vectors (all groups)
          named_bar_sync(kEpiGroups + 1, 128 * kEpiGroups);
          for (int n = etid; n < p.block_n; n += 128 * kEpiGroups) {
            float mul = 1.0f, add = __ldg(ep.bias + tc.n0 + n);
            if (ep.film_scale1p) {
              mul = __ldg(ep.film_scale1p + (long long)vb * ep.film_stride + tc.n0 + n);
              add = fmaf(add, mul, __ldg(ep.film_shift + (long long)vb * ep.film_stride + tc.n0 + n));
            }
            vec->mul[n] = mul;
            vec->add[n] = add;
          }
          named_bar_sync(kEpiGroups + 1, 128 * kEpiGroups);
          cached_key = key;
        }
        const int red_par = it & 1;
        const float nrows = ep.gn_partial ? (float)__popc(__ballot_sync(0xffffffffu, valid)) : 0.f;  // valid rows of this warp
        [[maybe_unused]] const bool tr = blockIdx.x == 0 && ew == 0 && lane == 0;  // traced warp (debug builds)
        CLPK_TRACE(tr, 100);
        mbar_wait(&bars->tmem_full[as], aphase);
        tc_fence_after();
        CLPK_TRACE(tr, 101);
        int ci = 0;
        for (int c = 32 * eg; c < p.block_n; c += cstep, ++ci) {
          uint8_t* sbuf = wslots + (size_t)slot * slot_bytes;
          uint32_t r[32];
          float gsums[16];
          float4* redw = vec->red + vec->idx(red_par, eg, ci, quarter);
          __syncwarp();
          tmem_ld16(taddr + (uint32_t)c, *reinterpret_cast<uint32_t(*)[16]>(&r[0]));
          tmem_ld16(taddr + (uint32_t)c + 16u, *reinterpret_cast<uint32_t(*)[16]>(&r[16]));
          if (slot_pending) {
            // Deferred slot bookkeeping of the previous chunk (its store was committed there), done by lane 0 while the
            // TMEM loads are in flight instead of right after the store, when its wait would stall the whole warp:
            //  residual: the load of chunk g+A targets the slot last stored from by chunk g+A-S -> at most S-A store groups
            //            may still be reading smem (A = S-1 -> 1; S = 1, A = 1 -> 0);
            //  plain   : this chunk writes the slot last stored from S chunks ago -> at most S-1 pending.
            if (lane == 0) {
              if (res_tma) {
                if (S - A >= 1) bulk_wait_group_read<1>(); else bulk_wait_group_read<0>();
                issue_res_load();
              } else {
                if (S - 1 >= 2) bulk_wait_group_read<2>(); else if (S - 1 == 1) bulk_wait_group_read<1>(); else bulk_wait_group_read<0>();
              }
            }
            slot_pending = false;
            __syncwarp();  // publishes the slot state to the other lanes before this chunk's smem accesses
          }
          tmem_ld_wait();
          CLPK_TRACE(tr, 102);
          if (!(CLPK_DBG(1))) {
            float v[32];  // (all epilogue arithmetic as packed fp32x2: half the issue slots of the scalar form)
#pragma unroll
            for (int j4 = 0; j4 < 8; ++j4) {
              const float4 m4 = *reinterpret_cast<const float4*>(&vec->mul[c + 4 * j4]);  // smem broadcast
              const float4 a4 = *reinterpret_cast<const float4*>(&vec->add[c + 4 * j4]);
              unpack2(fma2(pack2(__uint_as_float(r[4 * j4 + 0]), __uint_as_float(r[4 * j4 + 1])), pack2(m4.x, m4.y),
                           pack2(a4.x, a4.y)), v[4 * j4 + 0], v[4 * j4 + 1]);
              unpack2(fma2(pack2(__uint_as_float(r[4 * j4 + 2]), __uint_as_float(r[4 * j4 + 3])), pack2(m4.z, m4.w),
                           pack2(a4.z, a4.w)), v[4 * j4 + 2], v[4 * j4 + 3]);
            }
            // staging row = lane (row of the warp's sub-box); 16-byte piece j4 lives at piece (j4 ^ (lane & 7)) — the
            // 128B TMA swizzle (slots are 1024-byte aligned)
            uint8_t* srow = sbuf + lane * 128;
            if (res_tma) {
              CLPK_TRACE(tr, 103);
              mbar_wait(&res_full[slot], sphase);  // residual sub-box has landed (implies the slot was free)
              CLPK_TRACE(tr, 104);
              if (res16) {
                // 16-bit residual: row `lane` is 64 B, piece j8 (channels 8 j8 .. 8 j8 + 7) at (j8 ^ ((lane >> 1) & 3)) —
                // the layout the 16-bit result is stored back in, so every lane only ever touches its own row
                const uint8_t* rrow = sbuf + lane * 64;
#pragma unroll
                for (int j8 = 0; j8 < 4; ++j8) {
                  const uint4 q = *reinterpret_cast<const uint4*>(rrow + ((j8 ^ ((lane >> 1) & 3)) << 4));
                  const uint32_t qq[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
                  for (int k = 0; k < 4; ++k) {
                    const float2 f = unpack_op2(qq[k], f16);
                    unpack2(add2(pack2(v[8 * j8 + 2 * k], v[8 * j8 + 2 * k + 1]), pack2(f.x, f.y)), v[8 * j8 + 2 * k],
                            v[8 * j8 + 2 * k + 1]);
                  }
                }
              } else
#pragma unroll
              for (int j4 = 0; j4 < 8; ++j4) {
                const float4 q = *reinterpret_cast<const float4*>(srow + ((j4 ^ (lane & 7)) << 4));
                unpack2(add2(pack2(v[4 * j4 + 0], v[4 * j4 + 1]), pack2(q.x, q.y)), v[4 * j4 + 0], v[4 * j4 + 1]);
                unpack2(add2(pack2(v[4 * j4 + 2], v[4 * j4 + 3]), pack2(q.z, q.w)), v[4 * j4 + 2], v[4 * j4 + 3]);
              }
            }
            if (ep.gn_partial && !(CLPK_DBG(32))) {
              // per-thread (sum, sumsq) of this row's 32 values split by consumer-GroupNorm group; the cross-lane fold
              // happens after the chunk's store has been issued (it is off the store's critical path)
              if (cpg >= 32) gn_row_sums<1>(v, valid, nrows, lane, gsums, redw);
              else if (cpg == 16) gn_row_sums<2>(v, valid, nrows, lane, gsums, redw);
              else if (cpg == 8) gn_row_sums<4>(v, valid, nrows, lane, gsums, redw);
              else gn_row_sums<8>(v, valid, nrows, lane, gsums, redw);
            }
            if (st_f32) {
#pragma unroll
              for (int j4 = 0; j4 < 8; ++j4)
                *reinterpret_cast<float4*>(srow + ((j4 ^ (lane & 7)) << 4)) =
                    make_float4(v[4 * j4 + 0], v[4 * j4 + 1], v[4 * j4 + 2], v[4 * j4 + 3]);
            }
            if (st_16 || (ep.out_op && !(CLPK_DBG(16)))) {
              uint4 pk[4];
              if (f16) {
#pragma unroll
                for (int j8 = 0; j8 < 4; ++j8)
                  pk[j8] = make_uint4(pack_f16x2(v[8 * j8 + 0], v[8 * j8 + 1]), pack_f16x2(v[8 * j8 + 2], v[8 * j8 + 3]),
                                      pack_f16x2(v[8 * j8 + 4], v[8 * j8 + 5]), pack_f16x2(v[8 * j8 + 6], v[8 * j8 + 7]));
              } else {
#pragma unroll
                for (int j8 = 0; j8 < 4; ++j8)
                  pk[j8] = make_uint4(pack_bf16x2(v[8 * j8 + 0], v[8 * j8 + 1]), pack_bf16x2(v[8 * j8 + 2], v[8 * j8 + 3]),
                                      pack_bf16x2(v[8 * j8 + 4], v[8 * j8 + 5]), pack_bf16x2(v[8 * j8 + 6], v[8 * j8 + 7]));
              }
              if (st_16) {
                // 16-byte piece j (channels 8j .. 8j+7) of row `lane` lives at piece (j ^ ((lane >> 1) & 3)): 64B swizzle
                // (with a residual the slot still holds the fp32 sub-box other lanes may be reading: reconverge first)
                if (res_tma && !res16) __syncwarp();
                uint8_t* srow16 = sbuf + lane * 64;
#pragma unroll
                for (int j8 = 0; j8 < 4; ++j8) *reinterpret_cast<uint4*>(srow16 + ((j8 ^ ((lane >> 1) & 3)) << 4)) = pk[j8];
              } else if (valid) {
                uint4* op = reinterpret_cast<uint4*>(reinterpret_cast<uint16_t*>(ep.out_op) + opix * ldc + tc.n0 + c);
#pragma unroll
                for (int j8 = 0; j8 < 4; ++j8) op[j8] = pk[j8];
              }
            }
          }
          CLPK_TRACE(tr, 105);
          if (st_tma) {
            fence_proxy_async_smem();  // generic-proxy smem writes -> visible to the TMA (async proxy)
            __syncwarp();
            CLPK_TRACE(tr, 106);
            if (lane == 0 && !(CLPK_DBG(1))) {
              if (!(CLPK_DBG(1024))) tma_store_5d(&maps_out.m[tc.phase], sbuf, tc.n0 + c, tc.w0 + sub_w, 0, tc.h0 + sub_h, tc.b);
              bulk_commit_group();
            }
            slot_pending = !(CLPK_DBG(1));  // slot reuse is settled at the top of the next chunk
            if (++slot == (uint32_t)S) { slot = 0; sphase ^= 1u; }
          }
          CLPK_TRACE(tr, 107);
          if (ep.gn_partial && !(CLPK_DBG(33 | 512))) {
            // fold the row sums over the warp's 32 rows; the owning lanes park the warp's partials for the fixed-order
            // 4-warp fold at the end of the tile
            __syncwarp();
            if (cpg >= 32) gn_chunk_reduce<1>(gsums, lane, redw);
            else if (cpg == 16) gn_chunk_reduce<2>(gsums, lane, redw);
            else if (cpg == 8) gn_chunk_reduce<4>(gsums, lane, redw);
            else gn_chunk_reduce<8>(gsums, lane, redw);
          }
          CLPK_TRACE(tr, 108);
        }
        // accumulator drained: hand it back to the MMA issuer before the statistics fold
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if (NCTA == 2) mbar_arrive_cluster(lead_tmem_empty0 + (uint32_t)as * 8u);  // the issuer waits in the leader CTA
          else mbar_arrive(&bars->tmem_empty[as]);
        }
        if (ep.gn_partial && !(CLPK_DBG(33 | 512 | 256))) {
          // one barrier per tile: the 4 warps' partials of every chunk of this group are parked in red_t; thread
          // (chunk ci, pair pr) folds them in fixed order.  red is double-buffered by tile parity: a thread can only write
          // red[it & 1] again (tile it + 2) after passing the barrier of tile it + 1, which the folding threads of tile it
          // reach after their fold.
          named_bar_sync(bar_id, 128);
          const int nch = (p.block_n - 32 * eg + cstep - 1) / cstep;
          // The fold duty ROTATES over the group's 4 warps (tile it -> warp it % 4): its ~200 clk of dependent latency would
          // otherwise always delay the same warp, and the accumulator is only released when the slowest warp is done.
          if (quarter == (it & 3) && lane < nch * npairs && tc.ok) {
            const int fci = lane >> npairs_shift, pr = lane & (npairs - 1);   // npairs is 1, 2, 4 or 8
            const float4* rr = vec->red + vec->idx(red_par, eg, fci, 0) + pr;
            const float cnt = (float)(cpg >= 32 ? 32 : cpg);  // channels behind one partial of this chunk
            // re-base the 4 warps' shifted sums on the first valid warp's K (exact algebra, fixed order).  A warp without
            // valid rows — tile rows beyond the TMA box — is skipped: its accumulator rows, hence its K, are whatever the
            // uninitialised part of the A stage produced.
            float k0 = 0.f, s1 = 0.f, s2 = 0.f, n = 0.f;
            bool have = false;
#pragma unroll
            for (int wq = 0; wq < 4; ++wq) {
              const float4 a = rr[wq * 8];   // (sum, sumsq, K, valid rows) of warp wq
              if (a.w > 0.f) {
                if (!have) { k0 = a.z; have = true; }
                const float d = a.z - k0;
                const float nw = a.w * cnt;
                s1 += fmaf(nw, d, a.x);
                s2 += fmaf(nw * d, d, fmaf(2.f * d, a.x, a.y));
                n += nw;
              }
            }
            const float inv_n = __fdividef(1.0f, n);  // n is a small integer: the approximate reciprocal is within 1 ulp
            const float mean = fmaf(s1, inv_n, k0);
            const float m2 = fmaxf(fmaf(-s1 * inv_n, s1, s2), 0.f);
            const int ch = tc.n0 + 32 * eg + cstep * fci;
            int g, sub;
            if (cpg >= 32) {
              if (p.gn_cpg_shift >= 0) { g = ch >> p.gn_cpg_shift; sub = (ch & (cpg - 1)) >> 5; }
              else { g = ch / cpg; sub = (ch - g * cpg) >> 5; }
            } else {
              g = (ch >> p.gn_cpg_shift) + pr;   // cpg in {4, 8, 16}
              sub = 0;
            }
            const int mtile = (tc.phase * p.tiles_h + tc.th) * p.tiles_w + tc.tw;
            const int slotg = mtile * p.gn_sub + sub;
            float2* part = reinterpret_cast<float2*>(ep.gn_partial);
            part[((long long)tc.b * p.gn_slots + slotg) * p.gn_groups + g] = make_float2(mean, m2);
            // element count behind every triple of this slot: geometry only, so image 0's tiles publish it for all images
            if (tc.b == 0)
              reinterpret_cast<float*>(part + (long long)p.batch * p.gn_slots * p.gn_groups)[slotg] = n;
          }
        }
        CLPK_TRACE(tr, 109);
        continue;
      } else {
        // ------------------------------------------------ direct path (narrow N: the 3-channel `out` conv, NCHW store)
        mbar_wait(&bars->tmem_full[as], aphase);
        tc_fence_after();
        if (eg == 0) {
          for (int c = 0; c < p.block_n; c += 16) {
            uint32_t r[16];
            __syncwarp();  // tcgen05.ld is .sync.aligned: reconverge after the predicated stores of the last chunk
            tmem_ld16(taddr + (uint32_t)c, r);
            tmem_ld_wait();
            const int n = tc.n0 + c;
            if (valid && n < ldc && !(CLPK_DBG(1))) {
#pragma unroll
              for (int j = 0; j < 16; ++j) {
                if (n + j < ldc) {
                  float y = __uint_as_float(r[j]) + __ldg(ep.bias + n + j);
                  if (ep.film_scale1p)
                    y = fmaf(y, __ldg(ep.film_scale1p + (long long)tc.b * ep.film_stride + n + j),
                             __ldg(ep.film_shift + (long long)tc.b * ep.film_stride + n + j));
                  if (ep.resid) y += ep.resid[opix * ldc + n + j];
                  if (ep.resid_op) y += from_op(reinterpret_cast<const uint16_t*>(ep.resid_op)[opix * ldc + n + j], f16);
                  if (ep.out_f32) ep.out_f32[opix * ldc + n + j] = y;
                  if (ep.out_op) reinterpret_cast<uint16_t*>(ep.out_op)[opix * ldc + n + j] = to_op(y, f16);
                  if (ep.out_nchw) ep.out_nchw[(((long long)tc.b * ldc + n + j) * p.out_h + oh) * p.out_w + ow] = y;
                }
              }
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (NCTA == 2) mbar_arrive_cluster(lead_tmem_empty0 + (uint32_t)as * 8u);  // the issuer waits in the leader CTA
        else mbar_arrive(&bars->tmem_empty[as]);
      }
    }
    if (lane == 0 && st_tma) bulk_wait_group_all();  // all of this warp's stores retired before smem goes away
    CLPK_GTRACE(lane == 0 && ew == 0 && (blockIdx.x == 0 || blockIdx.x == gridDim.x - 2), 304);  // epilogue done
  }

  tc_fence_before();
  if (NCTA == 2) cluster_sync_all(); else __syncthreads();  // no remote arrive / multicast commit may still target a peer
  tc_fence_after();
  if (warp == 1) {
    if (NCTA == 2) tmem_dealloc_pair(tmem_base, (uint32_t)p.tmem_cols); else tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
  }
  CLPK_GTRACE(threadIdx.x == 0 && (blockIdx.x == 0 || blockIdx.x == gridDim.x - 2), 305);  // kernel exit
}

// ------------------------------------------------------------------------------------------------ cross-check kernel
// One thread per output element, same operands / tap tables / epilogue, plain fp32 FMAs.
__global__ void conv_direct_kernel(const uint16_t* __restrict__ x, const uint16_t* __restrict__ wpk,
                                   const IgemmParams p) {
  const bool f16 = p.op_f16 != 0;
  const long long total = (long long)p.phases * p.batch * p.grid_h * p.grid_w * p.cout_pad;
  const clpk_conv_epilogue& ep = p.ep;
  const int ldc = ep.cout_valid;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int n = (int)(idx % p.cout_pad);
    long long r = idx / p.cout_pad;
    const int w = (int)(r % p.grid_w); r /= p.grid_w;
    const int h = (int)(r % p.grid_h); r /= p.grid_h;
    const int b = (int)(r % p.batch);
    const int phase = (int)(r / p.batch);
    if (n >= ldc) continue;
    const long long ktot = (long long)p.taps * p.cin;
    const uint16_t* wrow = wpk + ((long long)phase * p.cout_pad + n) * ktot;
    float acc = 0.f;
    for (int t = 0; t < p.taps; ++t) {
      const int ti = phase * p.taps + t;
      const int cw = w + p.tap_dw[ti], ch = h + p.tap_dh[ti], cp = p.tap_p[ti];
      if (cw < 0 || cw >= p.a_dim_w || ch < 0 || ch >= p.a_dim_h) continue;  // zero padding
      const uint16_t* xp =
          x + (long long)b * p.a_stride_b + (long long)ch * p.a_stride_h + (long long)cp * p.a_stride_p +
          (long long)cw * p.a_stride_w + p.tap_x[ti];
      const uint16_t* wp = wrow + (long long)t * p.cin;
      for (int c = 0; c < p.cin; ++c) acc = fmaf(from_op(xp[c], f16), from_op(wp[c], f16), acc);
    }
    float y = acc + ep.bias[n];
    if (ep.film_scale1p)
      y = fmaf(y, ep.film_scale1p[(long long)b * ep.film_stride + n], ep.film_shift[(long long)b * ep.film_stride + n]);
    const int oh = h * p.out_scale + (phase >> 1), ow = w * p.out_scale + (phase & 1);
    const long long opix = ((long long)b * p.out_h + oh) * p.out_w + ow;
    if (ep.resid) y += ep.resid[opix * ldc + n];
    if (ep.resid_op) y += from_op(reinterpret_cast<const uint16_t*>(ep.resid_op)[opix * ldc + n], f16);
    if (ep.out_f32) ep.out_f32[opix * ldc + n] = y;
    if (ep.out_op) reinterpret_cast<uint16_t*>(ep.out_op)[opix * ldc + n] = to_op(y, f16);
    if (ep.out_nchw) ep.out_nchw[(((long long)b * ldc + n) * p.out_h + oh) * p.out_w + ow] = y;
  }
}

// ------------------------------------------------------------------------------------------------ weight packing
// Conv2d [Cout,Cin,3,3] -> [cout_pad][tap][Cin];  ConvTranspose2d [Cin,Cout,4,4] -> [phase][cout_pad][tap][Cin]
__global__ void pack_weight_kernel(const float* __restrict__ w, uint16_t* __restrict__ out, int kind, int cin,
                                   int cout, int cout_pad, int op_f16) {
  const int taps = (kind == CLPK_CONVT_4X4_S2) ? 4 : (kind == CLPK_CONV_1X1) ? 1 : 9;
  const int phases = (kind == CLPK_CONVT_4X4_S2) ? 4 : 1;
  const long long total = (long long)phases * cout_pad * taps * cin;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(idx % cin);
    long long r = idx / cin;
    const int t = (int)(r % taps); r /= taps;
    const int n = (int)(r % cout_pad);
    const int phase = (int)(r / cout_pad);
    float v = 0.f;
    if (n < cout) {
      if (kind == CLPK_CONVT_4X4_S2) {
        // phase (ph,pw), tap (a,b): kernel index kh = ph + 1 - 2*dh with dh = {0,-1} (ph=0) / {+1,0} (ph=1)
        const int ph = phase >> 1, pw = phase & 1, a = t >> 1, bb = t & 1;
        const int dh = (ph == 0) ? (a == 0 ? 0 : -1) : (a == 0 ? 1 : 0);
        const int dw = (pw == 0) ? (bb == 0 ? 0 : -1) : (bb == 0 ? 1 : 0);
        const int kh = ph + 1 - 2 * dh, kw = pw + 1 - 2 * dw;
        v = w[(((long long)c * cout + n) * 4 + kh) * 4 + kw];
      } else {
        v = w[((long long)n * cin + c) * taps + t];  // Conv2d [Cout,Cin,kh,kw] (3x3: taps 9; 1x1: taps 1)
      }
    }
    out[idx] = to_op(v, op_f16 != 0);
  }
}

int igemm_cout_pad(int cout) { return (cout + 15) / 16 * 16; }

// geometry of the M tiling (shared by igemm_setup and the GN-slot computation)
static void tile_geometry(int kind, int h_in, int w_in, int* grid_h, int* grid_w, int* wbox, int* hbox, int* phases) {
  *grid_h = (kind == CLPK_CONV_3X3_S2) ? h_in / 2 : h_in;
  *grid_w = (kind == CLPK_CONV_3X3_S2) ? w_in / 2 : w_in;
  *phases = (kind == CLPK_CONVT_4X4_S2) ? 4 : 1;
  // tile width: the whole row when it is <= 128 pixels and 32 tile rows form a rectangle of it (the epilogue moves
  // per-warp 32-row sub-boxes by TMA), else the largest power of two below it (edge tiles are masked / clipped)
  int wb = std::min(*grid_w, kTileM);
  if (!(wb % 32 == 0 || 32 % wb == 0)) {
    int p2 = 1;
    while (p2 * 2 <= wb) p2 *= 2;
    wb = p2;
  }
  *wbox = wb;
  *hbox = std::max(1, std::min(*grid_h, kTileM / *wbox));
}

// the chunked epilogue moves per-warp sub-boxes (32 tile rows) by TMA: 32 rows must form a rectangle of the tile
static bool warp_subbox_ok(int wbox) { return wbox % 32 == 0 || 32 % wbox == 0; }

int igemm_gn_slots(int kind, int h_in, int w_in, int cout, int gn_cpg) {
  if (gn_cpg <= 0 || cout % gn_cpg != 0 || cout % 32 != 0) return 0;
  if (!(gn_cpg == 4 || gn_cpg == 8 || gn_cpg == 16 || gn_cpg % 32 == 0)) return 0;
  if (cout / gn_cpg > 512) return 0;
  int gh, gw, wbox, hbox, phases;
  tile_geometry(kind, h_in, w_in, &gh, &gw, &wbox, &hbox, &phases);
  if (!warp_subbox_ok(wbox)) return 0;
  const int sub = gn_cpg >= 32 ? gn_cpg / 32 : 1;
  return phases * ((gh + hbox - 1) / hbox) * ((gw + wbox - 1) / wbox) * sub;
}
static bool slab_geometry_ok(int kind, int w_in, int cin, int block_n) {
  const bool slab_pair = block_n % 32 == 0 && block_n <= 128;
  return kind == CLPK_CONV_3X3_S1 && cin % 64 == 0 && w_in >= kTileM && (slab_pair || block_n <= 64);
}
bool igemm_xform_ok(int kind, int h_in, int w_in, int cin, int cout) {
  (void)h_in;
  { const char* e = getenv("CLPK_IGEMM_SLAB"); if (e && atoi(e) == 0) return false; }
  return slab_geometry_ok(kind, w_in, cin, igemm_block_n(igemm_cout_pad(cout))) && cin <= 1024;
}
int igemm_block_n(int cout_pad) {
  // widest N tile (<= 256) that divides the padded Cout; env CLPK_IGEMM_MAXBN caps it (experiments: 128-wide tiles for the
  // 256- / 512-channel layers give 2x the work items per launch at 2/3 of the MACs per operand byte)
  int cap = 256;
  { const char* e = getenv("CLPK_IGEMM_MAXBN"); if (e && atoi(e) >= 16 && atoi(e) <= 256) cap = atoi(e); }
  int best = 16;
  for (int n = 16; n <= cap && n <= cout_pad; n += 16)
    if (cout_pad % n == 0) best = n;
  return best;
}

// ------------------------------------------------------------------------------------------------ host setup
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode_fn() {
  static PFN_encodeTiled fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_encodeTiled>(p);
  });
  return fn;
}

int encode_tensor_map(CUtensorMap* m, const void* base, int rank, const cuuint64_t* dims, const cuuint64_t* strides_b,
                      const cuuint32_t* box, int swizzle_bytes, CUtensorMapDataType dtype) {
  PFN_encodeTiled fn = get_encode_fn();
  if (!fn) {
    set_error("cuTensorMapEncodeTiled entry point unavailable (no CUDA driver?)");
    return CLPK_ERR_CUDA;
  }
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUresult r = fn(m, dtype, (cuuint32_t)rank, const_cast<void*>(base), dims, strides_b, box,
                  estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  swizzle_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed with CUresult %d (rank %d dims %llu %llu box %u %u)", (int)r, rank,
              (unsigned long long)dims[0], (unsigned long long)dims[1], box[0], box[1]);
    return CLPK_ERR_CUDA;
  }
  return CLPK_OK;
}

int igemm_setup(const void* x_bf16, const void* w_packed, int kind, int batch, int h_in, int w_in, int cin, int cout,
                int op_dtype, const clpk_conv_epilogue* ep, IgemmLaunch* out) {
  CLPK_REQUIRE(kind >= 0 && kind <= 3, "conv kind %d unknown", kind);
  CLPK_REQUIRE(op_dtype == CLPK_OP_BF16 || op_dtype == CLPK_OP_F16, "operand dtype %d unknown", op_dtype);
  CLPK_REQUIRE(batch > 0 && h_in > 0 && w_in > 0, "bad conv geometry");
  CLPK_REQUIRE(cin % 32 == 0, "implicit-GEMM conv needs Cin %% 32 == 0 (got %d)", cin);
  CLPK_REQUIRE(ep && ep->bias, "conv epilogue needs a bias");
  CLPK_REQUIRE((reinterpret_cast<uintptr_t>(x_bf16) & 15) == 0 && (reinterpret_cast<uintptr_t>(w_packed) & 15) == 0,
               "conv operands must be 16-byte aligned");
  if (kind == CLPK_CONV_3X3_S2) CLPK_REQUIRE(h_in % 2 == 0 && w_in % 2 == 0, "stride-2 conv needs even H, W");
  IgemmParams& p = out->p;
  memset(&p, 0, sizeof(p));
  { const char* e = getenv("CLPK_IGEMM_DBG"); p.dbg = e ? atoi(e) : 0; }
  // default on: +0.5 % images/s measured in the step graph (the tail of the operand is what the preceding kernel wrote last)
  { const char* e = getenv("CLPK_IGEMM_REVERSE"); p.reverse = (e && atoi(e) == 0) ? 0 : 1; }
  p.batch = batch;
  p.op_f16 = (op_dtype == CLPK_OP_F16) ? 1 : 0;
  p.cin = cin;
  p.cout_pad = igemm_cout_pad(cout);
  p.block_n = igemm_block_n(p.cout_pad);
  // Tile configuration (measured on B200, tools/bench_conv.py): N >= 128 tiles run as CTA pairs (cta_group::2, M = 256,
  // the B tile split over the pair -> a third less operand fill per MMA) with 64-wide k-blocks; narrower tiles and the
  // pointwise stem GEMM run single-CTA (128-wide k-blocks when Cin allows: 8 MMAs per barrier handshake).
  p.ncta = (p.block_n >= 256 || (p.block_n >= 128 && kind != CLPK_CONV_1X1)) ? 2 : 1;
  { const char* e = getenv("CLPK_IGEMM_NCTA"); if (e && (atoi(e) == 1 || atoi(e) == 2)) p.ncta = atoi(e); }
  if (p.block_n % 32 != 0) p.ncta = 1;  // a CTA pair splits N in two halves that must stay multiples of 16
  p.block_k = (cin % 128 == 0 && p.ncta == 1) ? 128 : (cin % 64 == 0) ? 64 : 32;
  // a short K loop with a residual epilogue (the transposed conv: 4 taps) is epilogue-bound: spend the shared memory on
  // residual look-ahead slots rather than on 128-wide k-blocks
  if (p.block_k == 128 && (ep->resid || ep->resid_op)) p.block_k = 64;
  { const char* e = getenv("CLPK_IGEMM_BK");
    if (e && atoi(e) == 64 && p.block_k == 128) p.block_k = 64;
    if (e && atoi(e) == 128 && cin % 128 == 0) p.block_k = 128; }
  // Row-slab mainloop for 3x3 s1 convs whose M tile is one image row segment (W >= 128): operand fill traffic drops
  // from 9 A tiles + 9 B tiles to 3 slabs + 9 B tiles per channel block, and CTA pairs halve the B part again.
  p.slab = 0;
  // A stage holds the 3 weight blocks of a kernel row, so the per-CTA share of N must stay <= 64 rows: CTA pairs for
  // N <= 128, a single CTA for narrow N (the 3-channel `out` conv, N padded to 16).
  const bool slab_pair = p.block_n % 32 == 0 && p.block_n <= 128;
  if (slab_geometry_ok(kind, w_in, cin, p.block_n)) {
    const char* e = getenv("CLPK_IGEMM_SLAB");
    p.slab = (e && atoi(e) == 0) ? 0 : 1;
  }
  if (p.slab) { p.ncta = slab_pair ? 2 : 1; p.block_k = 64; }
  p.xform = 0;
  if (ep->in_scale || ep->in_shift) {
    CLPK_REQUIRE(ep->in_scale && ep->in_shift, "in_scale and in_shift must be given together");
    CLPK_REQUIRE(p.slab && igemm_xform_ok(kind, h_in, w_in, cin, cout),
                 "the fused input transform needs the row-slab mainloop (3x3 s1, W >= 128, cout <= 128, cin %% 64 == 0)");
    p.xform = 1;
    { const char* e = getenv("CLPK_XF_H2"); p.xform_h2 = (e && atoi(e) == 0) ? 0 : 1; }
  }
  p.n_tiles_n = p.cout_pad / p.block_n;
  p.kpt = cin / p.block_k;
  p.ep = *ep;
  if (p.ep.cout_valid <= 0) p.ep.cout_valid = cout;
  CLPK_REQUIRE(p.ep.cout_valid == cout, "cout_valid must equal cout");
  if (cout % 16 != 0)
    CLPK_REQUIRE(!p.ep.out_f32 && !p.ep.out_op && !p.ep.resid && !p.ep.resid_op, "Cout %% 16 != 0 supports the NCHW output only");
  if (p.ep.resid_op) {
    CLPK_REQUIRE(!p.ep.resid, "resid and resid_op are exclusive");
    CLPK_REQUIRE(p.ep.out_op && !p.ep.out_f32 && !p.ep.out_nchw, "a 16-bit residual needs a 16-bit-only NHWC output");
    CLPK_REQUIRE((reinterpret_cast<uintptr_t>(p.ep.resid_op) & 15) == 0, "residual tensor must be 16-byte aligned");
  }

  cuuint64_t dims[5], strides[4];
  const long long C = cin, W = w_in, H = h_in;
  if (kind == CLPK_CONV_3X3_S2) {
    p.grid_h = h_in / 2; p.grid_w = w_in / 2;
    p.phases = 1; p.taps = 9;
    p.out_h = p.grid_h; p.out_w = p.grid_w; p.out_scale = 1;
    for (int r = 0; r < 3; ++r)
      for (int s = 0; s < 3; ++s) {
        const int t = r * 3 + s;
        p.tap_p[t] = (r == 1) ? 0 : 1;
        p.tap_dh[t] = (r == 0) ? -1 : 0;
        p.tap_x[t] = (s == 1) ? 0 : cin;
        p.tap_dw[t] = (s == 0) ? -1 : 0;
      }
    dims[0] = 2 * C; dims[1] = W / 2; dims[2] = 2; dims[3] = H / 2; dims[4] = batch;
    strides[0] = 2 * C * 2; strides[1] = W * C * 2; strides[2] = 2 * W * C * 2; strides[3] = H * W * C * 2;
    p.a_stride_w = 2 * C; p.a_stride_p = W * C; p.a_stride_h = 2 * W * C; p.a_stride_b = H * W * C;
    p.a_dim_w = w_in / 2; p.a_dim_p = 2; p.a_dim_h = h_in / 2;
  } else {
    p.grid_h = h_in; p.grid_w = w_in;
    if (kind == CLPK_CONV_1X1) {
      p.phases = 1; p.taps = 1;
      p.out_h = h_in; p.out_w = w_in; p.out_scale = 1;
      p.tap_dh[0] = 0; p.tap_dw[0] = 0;
    } else if (kind == CLPK_CONV_3X3_S1) {
      p.phases = 1; p.taps = 9;
      p.out_h = h_in; p.out_w = w_in; p.out_scale = 1;
      for (int r = 0; r < 3; ++r)
        for (int s = 0; s < 3; ++s) {
          const int t = r * 3 + s;
          p.tap_dh[t] = r - 1; p.tap_dw[t] = s - 1;
        }
    } else {
      p.phases = 4; p.taps = 4;
      p.out_h = 2 * h_in; p.out_w = 2 * w_in; p.out_scale = 2;
      for (int phase = 0; phase < 4; ++phase)
        for (int t = 0; t < 4; ++t) {
          const int ph = phase >> 1, pw = phase & 1, a = t >> 1, bb = t & 1;
          p.tap_dh[phase * 4 + t] = (ph == 0) ? (a == 0 ? 0 : -1) : (a == 0 ? 1 : 0);
          p.tap_dw[phase * 4 + t] = (pw == 0) ? (bb == 0 ? 0 : -1) : (bb == 0 ? 1 : 0);
        }
    }
    dims[0] = C; dims[1] = W; dims[2] = 1; dims[3] = H; dims[4] = batch;
    strides[0] = C * 2; strides[1] = W * C * 2; strides[2] = W * C * 2; strides[3] = H * W * C * 2;
    p.a_stride_w = C; p.a_stride_p = 0; p.a_stride_h = W * C; p.a_stride_b = H * W * C;
    p.a_dim_w = w_in; p.a_dim_p = 1; p.a_dim_h = h_in;
  }
  // M tile = wbox x hbox patch, wbox*hbox <= 128
  { int gh, gw, ph; tile_geometry(kind, h_in, w_in, &gh, &gw, &p.wbox, &p.hbox, &ph); }
  p.tiles_w = (p.grid_w + p.wbox - 1) / p.wbox;
  p.tiles_h = (p.grid_h + p.hbox - 1) / p.hbox;
  // two rows per item (see the kRows2 producer): CTA-pair slab convs without the in-smem input transform; whether the two
  // operand rings fit shared memory is decided below, together with the staging slots
  p.rows2 = 0;
  if (p.slab && p.ncta == 2 && !p.xform && p.tiles_h >= 2 && 4 * p.block_n <= 512) {
    // only where pairing rows does not cost a wave: an item is twice as long, so ceil(items / pairs) must halve
    // (256 px at batch 8: 28 -> 14 items per CTA pair; 128 px: 7 -> 4 would be 14 % more work on the critical CTA)
    const int pairs = std::max(1, num_sms() / 2);
    const long long n1 = ((long long)batch * p.tiles_h * p.tiles_w + 1) / 2 * p.n_tiles_n;
    const long long n2 = ((long long)batch * ((p.tiles_h + 1) / 2) * p.tiles_w + 1) / 2 * p.n_tiles_n;
    const bool no_extra_wave = 2 * ((n2 + pairs - 1) / pairs) <= (n1 + pairs - 1) / pairs;
    // OFF by default (env CLPK_IGEMM_ROWS2=1: where wave-neutral, =2: always): measured on B200 it is neutral at best
    // (256 px conv1 123.9 -> 126.0 us with 44 % less operand fill; conv2 loses a residual staging slot to the second ring
    // unless the weight ring is cut to 2) — the slab mainloop is NOT bound by operand-fill bytes.
    const char* e = getenv("CLPK_IGEMM_ROWS2");
    p.rows2 = e ? (atoi(e) == 1 ? (no_extra_wave ? 1 : 0) : atoi(e) == 2 ? 1 : 0) : 0;
  }
  auto compute_tiles = [&]() -> int {
    const int item_rows = p.rows2 ? (p.tiles_h + 1) / 2 : p.tiles_h;   // row tiles (or row-tile pairs) per image column
    p.spatial_tiles = batch * item_rows * p.tiles_w;
    const long long nt = (long long)((p.spatial_tiles + p.ncta - 1) / p.ncta) * p.phases * p.n_tiles_n;
    CLPK_REQUIRE(nt < (1ll << 30), "too many tiles");
    p.num_tiles = (int)nt;
    p.fd_tiles_n = make_fastdiv(p.n_tiles_n);
    p.fd_phases = make_fastdiv(p.phases);
    p.fd_tiles_w = make_fastdiv(p.tiles_w);
    p.fd_tiles_h = make_fastdiv(item_rows);
    // accumulator ring: 2 TMEM buffers; env CLPK_IGEMM_NACC=4 makes it 4 when the tile is narrow enough (block_n <= 128 ->
    // 512 columns).  Measured neutral (stem 103 -> 105 us, transposed conv 138 -> 138 us, whole step +0.8 %): the short-K
    // kernels are not bound by the MMA -> epilogue -> MMA hand-over.  The two-rows variant always uses 4 (2 rows x 2).
    p.nacc = 2;
    { const char* e = getenv("CLPK_IGEMM_NACC"); if (e && atoi(e) == 4 && !p.rows2 && 4 * p.block_n <= 512) p.nacc = 4; }
    p.nacc_shift = p.nacc == 4 ? 2 : 1;
    p.tmem_cols = 32;
    while (p.tmem_cols < (p.rows2 ? 4 : p.nacc) * p.block_n) p.tmem_cols *= 2;
    return CLPK_OK;
  };
  CLPK_TRY_RC(compute_tiles());
  const int stage_bytes = p.slab ? kSlabABytes + 3 * (p.block_n / p.ncta) * 128
                                 : kTileM * p.block_k * 2 + (p.block_n / p.ncta) * p.block_k * 2;
  // chunked epilogue whenever the tile has >= 32 channels and an NHWC output; its fp32 output (if any) is staged
  // through smem and written by TMA, a 16-bit output is stored directly from registers
  const bool f32_ok = p.ep.out_f32 != nullptr && (reinterpret_cast<uintptr_t>(p.ep.out_f32) & 15) == 0;
  p.chunked = (p.block_n % 32 == 0 && cout % 32 == 0 && (f32_ok || (p.ep.out_op && !p.ep.out_f32)) && !p.ep.out_nchw &&
               warp_subbox_ok(p.wbox)) ? 1 : 0;
  // The dynamic shared memory is declared __align__(1024) and is the kernel's only shared memory, so its base is
  // 1024-aligned and no slack is reserved (the kernel traps if that ever fails to hold); CLPK_IGEMM_SLACK=1 restores it.
  { const char* e = getenv("CLPK_IGEMM_SLACK"); p.smem_slack = (e && atoi(e) != 0) ? 1024 : 0; }
  const int fixed = p.smem_slack + epi_vector_bytes(p.block_n, p.xform ? cin : 0) + (int)sizeof(PipeBarriers) + 16;
  p.n_staging = 0;
  const bool op16_ok = p.ep.out_op != nullptr && (reinterpret_cast<uintptr_t>(p.ep.out_op) & 15) == 0;
  const bool any_resid = p.ep.resid || p.ep.resid_op;
  // an epilogue that moves only 16-bit tensors through the slots (16-bit-only output, no residual or a 16-bit one) gets
  // half-size slots: twice the residual look-ahead, or more pipeline stages, out of the same shared memory
  p.slot_bytes = (!p.ep.out_f32 && !p.ep.resid && op16_ok) ? kWarpSlotBytes / 2 : kWarpSlotBytes;
  { const char* e = getenv("CLPK_IGEMM_SLOT4K"); if (e && atoi(e) != 0) p.slot_bytes = kWarpSlotBytes; }
  const int staging_bytes = 4 * p.slot_bytes;   // one staging buffer = the 4 warp slots of an epilogue group
  if (p.chunked && (p.ep.out_f32 || any_resid || op16_ok)) {
    CLPK_REQUIRE(p.ep.out_f32 != nullptr || !any_resid || op16_ok, "a residual input needs an NHWC output");
    // as many staging slots per epilogue group (<= 3) as the smem ring can spare without losing depth; a residual
    // epilogue keeps (slots - 1) chunk loads in flight per group, so its throughput hangs on this
    // (measured: the 256-wide CTA-pair tiles of the 64x64 / 32x32 layers prefer a 5-deep ring over a second staging slot)
    int want_stages = (p.block_k == 128 || p.slab) ? 3 : (p.ncta == 2 && p.block_n >= 256) ? 5 : 4;
    // the in-smem transform adds a third phase (fill -> normalise -> MMA) to every stage's life: one more stage in flight
    if (p.xform) { const char* e = getenv("CLPK_XF_STAGES"); want_stages = e ? atoi(e) : 3; }
    int per_group = p.ep.resid ? 3 : p.ep.resid_op ? 4 : 2;
    { const char* e = getenv("CLPK_IGEMM_SLOTS16"); if (e && p.ep.resid_op && atoi(e) >= 1 && atoi(e) <= 6) per_group = atoi(e); }
    while (per_group > 1 && (kSmemBudget - fixed - kEpiGroups * per_group * staging_bytes) / stage_bytes < want_stages) --per_group;
    { const char* e = getenv("CLPK_IGEMM_SLOTS"); if (e && atoi(e) >= 1 && atoi(e) <= 4) per_group = atoi(e); }
    if (kind == CLPK_CONVT_4X4_S2) {  // (a transposed conv is bound by its residual look-ahead: slots beat ring depth)
      const char* e = getenv("CLPK_CONVT_SLOTS");
      const int want = e ? atoi(e) : per_group;
      if (want >= 1 && want <= 6 && (kSmemBudget - fixed - kEpiGroups * want * staging_bytes) / stage_bytes >= 3) per_group = want;
    }
    p.n_staging = kEpiGroups * per_group;
    if ((kSmemBudget - fixed - p.n_staging * staging_bytes) / stage_bytes < 2) {
      p.n_staging = 0;
      if (p.ep.out_f32 || any_resid) p.chunked = 0;  // (a 16-bit-only output falls back to direct stores from registers)
    }
  }
  p.gn_groups = p.gn_slots = 0;
  p.gn_sub = 1;
  if (p.ep.gn_partial) {
    p.gn_slots = igemm_gn_slots(kind, h_in, w_in, cout, p.ep.gn_cpg);
    CLPK_REQUIRE(p.gn_slots > 0 && p.chunked,
                 "fused GroupNorm statistics unsupported for this conv (cout=%d cpg=%d)", cout, p.ep.gn_cpg);
    p.gn_groups = cout / p.ep.gn_cpg;
    p.gn_sub = p.ep.gn_cpg >= 32 ? p.ep.gn_cpg / 32 : 1;
    p.gn_cpg_shift = -1;
    for (int sft = 2; sft < 16; ++sft)
      if ((1 << sft) == p.ep.gn_cpg) p.gn_cpg_shift = sft;
  }
  // residual sub-boxes (of 4 per chunk) the producer prefetches into L2 per tile: measured neutral-to-harmful for the
  // 3x3 convs (their per-warp look-ahead loads cover the latency), a small win for the short-K transposed conv
  p.res_ahead = (kind == CLPK_CONVT_4X4_S2) ? 2 : 0;
  { const char* e = getenv("CLPK_IGEMM_PREFETCH"); if (e) p.res_ahead = std::max(0, std::min(4, atoi(e))); }
  p.stages = std::min(kMaxStages, (kSmemBudget - fixed - p.n_staging * staging_bytes) / stage_bytes);
  CLPK_REQUIRE(p.stages >= 2, "tile does not fit shared memory");
  out->smem_bytes = p.stages * stage_bytes + p.n_staging * staging_bytes + fixed;
  p.stages_b = 0;
  if (p.rows2) {
    // two rings: slabs (17 KB each; 2 in use + prefetch) and weight blocks (3 taps of a kernel row; 1 in use + prefetch).
    // Wanted: >= 4 slabs + 3 weight blocks; staging slots are given up (down to 1 per group) before ring depth.
    const int bb = 3 * (p.block_n / p.ncta) * 128;
    int nb = 2;
    { const char* e = getenv("CLPK_ROWS2_NB"); if (e && atoi(e) >= 2 && atoi(e) <= 4) nb = atoi(e); }
    int per_group = p.n_staging / kEpiGroups;
    auto na_for = [&](int pg) { return (kSmemBudget - fixed - kEpiGroups * pg * staging_bytes - nb * bb) / kSlabABytes; };
    while (per_group > 1 && na_for(per_group) < 4) --per_group;
    const int na = std::min(kMaxStages, na_for(per_group));
    if (p.n_staging > 0 && na >= 3) {
      p.n_staging = kEpiGroups * per_group;
      p.stages = na;
      p.stages_b = nb;
      out->smem_bytes = na * kSlabABytes + nb * bb + p.n_staging * staging_bytes + fixed;
    } else {
      p.rows2 = 0;   // does not fit: back to one row per item
      CLPK_TRY_RC(compute_tiles());
    }
  }
  out->grid = p.ncta * std::min(p.num_tiles, num_sms() / p.ncta);

  const int atom_k = std::min(p.block_k, 64);  // one TMA box = one swizzle atom (<= 128 bytes of K per row)
  const int swz = atom_k * 2;
  cuuint32_t box_a[5] = {(cuuint32_t)atom_k, (cuuint32_t)(p.slab ? p.wbox + 2 : p.wbox), 1, (cuuint32_t)p.hbox, 1};
  const CUtensorMapDataType op_dt = p.op_f16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
  int rc = encode_tensor_map(&out->map_a, x_bf16, 5, dims, strides, box_a, swz, op_dt);
  if (rc) return rc;
  cuuint64_t wdims[2] = {(cuuint64_t)p.taps * cin, (cuuint64_t)p.phases * p.cout_pad};
  cuuint64_t wstr[1] = {(cuuint64_t)p.taps * cin * 2};
  cuuint32_t box_w[2] = {(cuuint32_t)atom_k, (cuuint32_t)(p.block_n / p.ncta)};
  rc = encode_tensor_map(&out->map_w, w_packed, 2, wdims, wstr, box_w, swz, op_dt);
  if (rc) return rc;
  // fp32 NHWC output maps for the TMA-store epilogue: [C, Wgrid, 1, Hgrid, B] per phase (transposed conv: the phase
  // (ph,pw) owns output pixels (2h+ph, 2w+pw) -> base offset + doubled w/h strides)
  memset(&out->maps_out, 0, sizeof(out->maps_out));
  const int wsub = std::min(p.wbox, 32);
  if (p.chunked && !p.ep.out_f32 && p.ep.out_op && p.n_staging > 0) {
    // 16-bit NHWC output map for the staged store of a 16-bit-only output: 32 channels (64 B) per row, 64B swizzle
    const long long CO = cout, OW = p.out_w, OH = p.out_h, sc = p.out_scale;
    cuuint64_t odims[5] = {(cuuint64_t)CO, (cuuint64_t)p.grid_w, 1, (cuuint64_t)p.grid_h, (cuuint64_t)batch};
    cuuint64_t ostr[4] = {(cuuint64_t)(sc * CO * 2), (cuuint64_t)(sc * OW * CO * 2), (cuuint64_t)(sc * OW * CO * 2),
                          (cuuint64_t)(OH * OW * CO * 2)};
    cuuint32_t obox[5] = {32, (cuuint32_t)wsub, 1, (cuuint32_t)(32 / wsub), 1};  // one epilogue warp's 32 tile rows
    for (int phase = 0; phase < p.phases; ++phase) {
      const long long off = ((long long)(phase >> 1) * OW + (phase & 1)) * CO;
      rc = encode_tensor_map(&out->maps_out.m[phase], reinterpret_cast<uint16_t*>(p.ep.out_op) + off, 5, odims, ostr, obox, 64, op_dt);
      if (rc) return rc;
    }
  }
  memset(&out->maps_res, 0, sizeof(out->maps_res));
  if (p.chunked && p.ep.resid_op) {  // 16-bit residual: same geometry as the 16-bit output map
    const long long CO = cout, OW = p.out_w, OH = p.out_h, sc = p.out_scale;
    cuuint64_t odims[5] = {(cuuint64_t)CO, (cuuint64_t)p.grid_w, 1, (cuuint64_t)p.grid_h, (cuuint64_t)batch};
    cuuint64_t ostr[4] = {(cuuint64_t)(sc * CO * 2), (cuuint64_t)(sc * OW * CO * 2), (cuuint64_t)(sc * OW * CO * 2),
                          (cuuint64_t)(OH * OW * CO * 2)};
    cuuint32_t obox[5] = {32, (cuuint32_t)wsub, 1, (cuuint32_t)(32 / wsub), 1};
    for (int phase = 0; phase < p.phases; ++phase) {
      const long long off = ((long long)(phase >> 1) * OW + (phase & 1)) * CO;
      rc = encode_tensor_map(&out->maps_res.m[phase], const_cast<uint16_t*>(reinterpret_cast<const uint16_t*>(p.ep.resid_op)) + off, 5,
                             odims, ostr, obox, 64, op_dt);
      if (rc) return rc;
    }
  }
  if (p.chunked && (p.ep.out_f32 || p.ep.resid)) {
    const long long CO = cout, OW = p.out_w, OH = p.out_h, sc = p.out_scale;
    cuuint64_t odims[5] = {(cuuint64_t)CO, (cuuint64_t)p.grid_w, 1, (cuuint64_t)p.grid_h, (cuuint64_t)batch};
    cuuint64_t ostr[4] = {(cuuint64_t)(sc * CO * 4), (cuuint64_t)(sc * OW * CO * 4), (cuuint64_t)(sc * OW * CO * 4),
                          (cuuint64_t)(OH * OW * CO * 4)};
    cuuint32_t obox[5] = {32, (cuuint32_t)wsub, 1, (cuuint32_t)(32 / wsub), 1};  // one epilogue warp's 32 tile rows
    for (int phase = 0; phase < p.phases; ++phase) {
      const long long off = ((long long)(phase >> 1) * OW + (phase & 1)) * CO;
      if (p.ep.out_f32) {
        rc = encode_tensor_map(&out->maps_out.m[phase], p.ep.out_f32 + off, 5, odims, ostr, obox, 128,
                        CU_TENSOR_MAP_DATA_TYPE_FLOAT32);
        if (rc) return rc;
      }
      if (p.ep.resid) {
        CLPK_REQUIRE((reinterpret_cast<uintptr_t>(p.ep.resid) & 15) == 0, "residual tensor must be 16-byte aligned");
        rc = encode_tensor_map(&out->maps_res.m[phase], const_cast<float*>(p.ep.resid) + off, 5, odims, ostr, obox, 128,
                        CU_TENSOR_MAP_DATA_TYPE_FLOAT32);
        if (rc) return rc;
      }
    }
  }
  return rc;
}

template <int BK, int NC, bool SLAB = false, bool XF = false, bool R2 = false>
static cudaError_t set_smem_attr() {
  return cudaFuncSetAttribute(conv_igemm_kernel<BK, NC, SLAB, XF, R2>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBudget);
}

// function attributes are per device (context): set them once for every device this process launches on
int igemm_init() {
  static std::mutex mu;
  static bool done[64] = {};
  int dev = 0;
  CLPK_CHECK_CUDA(cudaGetDevice(&dev));
  CLPK_REQUIRE(dev >= 0 && dev < 64, "device index %d out of range", dev);
  std::lock_guard<std::mutex> lock(mu);
  if (done[dev]) return CLPK_OK;
  cudaError_t attr_err = set_smem_attr<64, 1>();
  if (attr_err == cudaSuccess) attr_err = set_smem_attr<32, 1>();
  if (attr_err == cudaSuccess) attr_err = set_smem_attr<128, 1>();
  if (attr_err == cudaSuccess) attr_err = set_smem_attr<64, 2>();
  if (attr_err == cudaSuccess) attr_err = set_smem_attr<32, 2>();
  if (attr_err == cudaSuccess) attr_err = set_smem_attr<128, 2>();
  if (attr_err == cudaSuccess) attr_err = set_smem_attr<64, 2, true>();
  if (attr_err == cudaSuccess) attr_err = set_smem_attr<64, 1, true>();
  if (attr_err == cudaSuccess) attr_err = set_smem_attr<64, 2, true, true>();
  if (attr_err == cudaSuccess) attr_err = set_smem_attr<64, 1, true, true>();
  if (attr_err == cudaSuccess) attr_err = set_smem_attr<64, 2, true, false, true>();
  CLPK_CHECK_CUDA(attr_err);
  done[dev] = true;
  return CLPK_OK;
}

template <int BK, int NC, bool SLAB = false, bool XF = false, bool R2 = false>
static cudaError_t launch_variant(const IgemmLaunch& L, cudaStream_t stream) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)L.grid);
  cfg.blockDim = dim3(kNumThreads + (XF ? kXformThreads : 0));
  cfg.dynamicSmemBytes = (size_t)L.smem_bytes;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  int na = 0;
  if (NC > 1) {
    attr[na].id = cudaLaunchAttributeClusterDimension;
    attr[na].val.clusterDim.x = NC;
    attr[na].val.clusterDim.y = 1;
    attr[na].val.clusterDim.z = 1;
    ++na;
  }
  if (pdl_enabled()) {
    attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  cfg.attrs = attr;
  cfg.numAttrs = na;
  return cudaLaunchKernelEx(&cfg, conv_igemm_kernel<BK, NC, SLAB, XF, R2>, L.map_a, L.map_w, L.maps_out, L.maps_res, L.p);
}

int igemm_launch(const IgemmLaunch& L, cudaStream_t stream) {
  int irc = igemm_init();
  if (irc) return irc;
  cudaError_t e;
  if (L.p.slab && L.p.rows2) e = launch_variant<64, 2, true, false, true>(L, stream);
  else if (L.p.slab && L.p.xform) e = (L.p.ncta == 2) ? launch_variant<64, 2, true, true>(L, stream) : launch_variant<64, 1, true, true>(L, stream);
  else if (L.p.slab) e = (L.p.ncta == 2) ? launch_variant<64, 2, true>(L, stream) : launch_variant<64, 1, true>(L, stream);
  else if (L.p.block_k == 128) e = (L.p.ncta == 2) ? launch_variant<128, 2>(L, stream) : launch_variant<128, 1>(L, stream);
  else if (L.p.block_k == 64) e = (L.p.ncta == 2) ? launch_variant<64, 2>(L, stream) : launch_variant<64, 1>(L, stream);
  else e = (L.p.ncta == 2) ? launch_variant<32, 2>(L, stream) : launch_variant<32, 1>(L, stream);
  CLPK_CHECK_CUDA(e);
  CLPK_CHECK_LAUNCH();
  return CLPK_OK;
}

int direct_launch(const IgemmLaunch& L, const void* x_bf16, const void* w_packed, cudaStream_t stream) {
  const long long total = (long long)L.p.phases * L.p.batch * L.p.grid_h * L.p.grid_w * L.p.cout_pad;
  const int blocks = (int)std::min<long long>((total + 255) / 256, 65535ll * 8);
  conv_direct_kernel<<<blocks, 256, 0, stream>>>(reinterpret_cast<const uint16_t*>(x_bf16),
                                                 reinterpret_cast<const uint16_t*>(w_packed), L.p);
  CLPK_CHECK_LAUNCH();
  return CLPK_OK;
}

}  // namespace clpk

using namespace clpk;

extern "C" int64_t clpk_pack_conv_weight(const float* w_dev, void* out_bf16_dev, int kind, int cin, int cout,
                                         int op_dtype, void* stream) {
  if (kind < 0 || kind > 3 || cin <= 0 || cout <= 0 || (op_dtype != CLPK_OP_BF16 && op_dtype != CLPK_OP_F16)) {
    set_error("clpk_pack_conv_weight: bad arguments");
    return -1;
  }
  const int cout_pad = igemm_cout_pad(cout);
  const int taps = (kind == CLPK_CONVT_4X4_S2) ? 4 : (kind == CLPK_CONV_1X1) ? 1 : 9;
  const int phases = (kind == CLPK_CONVT_4X4_S2) ? 4 : 1;
  const long long total = (long long)phases * cout_pad * taps * cin;
  if (!out_bf16_dev) return total;
  const int blocks = (int)std::min<long long>((total + 255) / 256, 65535);
  pack_weight_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(w_dev, reinterpret_cast<uint16_t*>(out_bf16_dev), kind,
                                                               cin, cout, cout_pad, op_dtype == CLPK_OP_F16);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("pack_weight_kernel launch failed: %s", cudaGetErrorString(e));
    return -2;
  }
  count_launch();
  return total;
}

extern "C" int clpk_conv_in_affine_supported(int kind, int h_in, int w_in, int cin, int cout) {
  if (kind < 0 || kind > 3 || h_in <= 0 || w_in <= 0 || cin <= 0 || cout <= 0) return 0;
  return igemm_xform_ok(kind, h_in, w_in, cin, cout) ? 1 : 0;
}

extern "C" int clpk_conv_gn_slots(int kind, int h_in, int w_in, int cout, int gn_cpg) {
  if (kind < 0 || kind > 3 || h_in <= 0 || w_in <= 0 || cout <= 0) return 0;
  return igemm_gn_slots(kind, h_in, w_in, cout, gn_cpg);
}

extern "C" int clpk_conv_igemm(const void* x, const void* w, int kind, int batch, int h_in, int w_in, int cin, int cout,
                               int op_dtype, const clpk_conv_epilogue* ep, void* stream) {
  IgemmLaunch L;
  int rc = igemm_setup(x, w, kind, batch, h_in, w_in, cin, cout, op_dtype, ep, &L);
  if (rc) return rc;
  return igemm_launch(L, (cudaStream_t)stream);
}

extern "C" int clpk_conv_direct(const void* x, const void* w, int kind, int batch, int h_in, int w_in, int cin,
                                int cout, int op_dtype, const clpk_conv_epilogue* ep, void* stream) {
  IgemmLaunch L;
  int rc = igemm_setup(x, w, kind, batch, h_in, w_in, cin, cout, op_dtype, ep, &L);
  if (rc) return rc;
  return direct_launch(L, x, w, (cudaStream_t)stream);
}

#ifdef CLPK_IGEMM_DEBUG
// debug builds only: copies the phase trace (pairs of tag, clock64) to the host and resets it; returns the pair count
extern "C" int clpk_debug_trace(long long* out_host, int max_pairs) {
  int n = 0;
  cudaDeviceSynchronize();
  cudaMemcpyFromSymbol(&n, g_trace_n, sizeof(int));
  n = std::min(std::min(n, 4096), max_pairs);
  cudaMemcpyFromSymbol(out_host, g_trace, sizeof(long long) * 2 * n);
  const int zero = 0;
  cudaMemcpyToSymbol(g_trace_n, &zero, sizeof(int));
  return n;
}
#endif
