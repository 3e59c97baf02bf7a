// conv_igemm.cuh — host-visible description of one implicit-GEMM convolution launch.
#pragma once
#include "common.cuh"

namespace clpk {

constexpr int kMaxTapEntries = 16;  // 9 taps (3x3) or 4 phases x 4 taps (ConvTranspose 4x4 s2)

// Division by a launch-time constant as multiply + shift (exact for 0 <= n < 2^31): q = (n * mul) >> shift with
// mul = ceil(2^(31 + s) / d), s = ceil(log2 d), shift = 31 + s.  The tile decoder runs in every warp for every tile; four
// hardware-less integer divisions there were ~150 of the ~240 instructions a warp spends per tile outside its chunks.
struct FastDiv {
  uint32_t mul, shift, d;
};
inline FastDiv make_fastdiv(int d) {
  FastDiv f;
  uint32_t s = 0;
  while ((1u << s) < (uint32_t)d) ++s;
  const unsigned long long num = 1ull << (31 + s);
  f.mul = (uint32_t)((num + (unsigned long long)d - 1) / (unsigned long long)d);
  f.shift = 31 + s;
  f.d = (uint32_t)d;
  return f;
}

// Everything the kernels need, passed by value.
struct IgemmParams {
  // tile grid: output pixels for the 3x3 convs, INPUT pixels for the transposed conv (each phase maps them 1:1)
  int batch, grid_h, grid_w;
  int cin, cout_pad;        // cout_pad = GEMM N per phase (multiple of block_n)
  int block_k;              // 64 (128B swizzle) or 32 (64B swizzle)
  int block_n, n_tiles_n;
  int wbox, hbox;           // an M tile is a wbox x hbox patch of one image (wbox*hbox <= 128 rows)
  int tiles_w, tiles_h;
  int phases, taps, kpt;    // kpt = k-blocks per tap = cin / block_k
  int num_tiles, stages, tmem_cols;   // num_tiles = work items: ceil(spatial tiles / ncta) * phases * n_tiles_n
  int rows2;                // 1: two-rows-per-item slab variant (CTA pairs): a work item = row tiles (h0, h0+1) of one column
                            // segment with two accumulators; slabs and weight blocks are staged once for both rows
  int stages_b;             // rows2: depth of the weight-block ring (`stages` = depth of the slab ring)
  int ncta;                 // 1, or 2 = CTA pairs (tcgen05 cta_group::2): two M tiles share one B tile split over the pair
  int spatial_tiles;        // batch * tiles_h * tiles_w
  int op_f16;               // operand format: 1 = fp16, 0 = bf16
  int slab;                 // 1: row-slab mainloop (3x3 s1, one image row per tile): a stage holds ONE (wbox+2)-pixel slab of
                            // input row h+r-1 and the three weight blocks of taps (r, 0..2); the taps read the slab at row
                            // offsets 0/1/2 through shifted smem descriptors -> each A byte is fetched 3x instead of 9x
  int chunked;              // 1: chunked epilogue (32 columns at a time: folded vectors, fused GN stats, staged TMA store of
                            // the fp32 output if there is one); 0: narrow direct path (the 3-channel `out` conv)
  int n_staging;            // epilogue staging buffers (4 warp slots each) for the TMA-store path, 0 = direct stores
  int nacc, nacc_shift;     // TMEM accumulator ring depth (2 or 4) and its log2
  int slot_bytes;           // one warp slot: 32 rows x 128 B (4096: fp32 output / residual) or x 64 B (2048: all-16-bit epilogue)
  int res_ahead;            // residual chunks the epilogue leader keeps in flight ahead of the one being processed
  int gn_groups, gn_slots, gn_sub;  // fused GroupNorm statistics: groups, partial slots per image, chunks per group slot
  int gn_cpg_shift;         // log2(channels per partial) when that is a power of two, else -1
  int reverse;              // 1 (default; env CLPK_IGEMM_REVERSE=0 turns it off): tiles are walked from the END of the tensor —
                            // the tail of the operand is what the preceding kernel wrote last and is the likeliest L2 resident
  FastDiv fd_tiles_n, fd_phases, fd_tiles_w, fd_tiles_h;  // decode_tile divisors
  int xform;                // 1: input transform fused into the A path (ep.in_scale / in_shift; slab mainloop only): 4 extra
                            // warps normalise every landed slab in place before the tensor cores read it
  int xform_h2;             // 1 (default; env CLPK_XF_H2=0 turns it off): the transform's SiLU runs on packed halves (fp16 operands)
  int smem_slack;           // bytes reserved for aligning the dynamic shared memory to 1024 (0: the base is trusted)
  int dbg;                  // perf-debug switches (env CLPK_IGEMM_DBG): 1 no epilogue memory ops, 2 no MMA, 4 no A loads, 8 no B loads
  // A-operand coordinates (5-D view of the NHWC input, see make_a_map): per (phase*taps + tap)
  int tap_x[kMaxTapEntries], tap_dw[kMaxTapEntries], tap_p[kMaxTapEntries], tap_dh[kMaxTapEntries];
  // element strides of that 5-D view (used by the CUDA-core cross-check kernel only)
  long long a_stride_w, a_stride_p, a_stride_h, a_stride_b;
  int a_dim_w, a_dim_p, a_dim_h;
  // output addressing: pixel (h, w) of the tile grid, phase (ph, pw) -> (h*out_scale + ph, w*out_scale + pw)
  int out_h, out_w, out_scale;
  clpk_conv_epilogue ep;
};

struct alignas(64) OutMaps {
  CUtensorMap m[4];  // fp32 NHWC output viewed per transposed-conv phase (entry 0 for ordinary convs)
};

struct IgemmLaunch {
  IgemmParams p;
  CUtensorMap map_a;
  CUtensorMap map_w;
  OutMaps maps_out;
  OutMaps maps_res;  // same geometry over the residual tensor (== maps_out when the conv updates in place)
  int grid;
  int smem_bytes;
};

// Fills `out` for a convolution of `kind` over x (16-bit NHWC [batch,h_in,w_in,cin]) with packed weights.
int igemm_setup(const void* x_op, const void* w_packed, int kind, int batch, int h_in, int w_in, int cin, int cout,
                int op_dtype, const clpk_conv_epilogue* ep, IgemmLaunch* out);
int igemm_init();  // one-time function attributes; call outside stream capture
int igemm_launch(const IgemmLaunch& L, cudaStream_t stream);
int direct_launch(const IgemmLaunch& L, const void* x_bf16, const void* w_packed, cudaStream_t stream);

// cuTensorMapEncodeTiled through the runtime's driver entry point (no link-time libcuda dependency)
int encode_tensor_map(CUtensorMap* m, const void* base, int rank, const cuuint64_t* dims, const cuuint64_t* strides_b,
                      const cuuint32_t* box, int swizzle_bytes, CUtensorMapDataType dtype);

// padded GEMM-N of a conv with `cout` output channels, and the UMMA N tile chosen for it
int igemm_gn_slots(int kind, int h_in, int w_in, int cout, int gn_cpg);  // <= 0: unsupported
// floats of a gn_partial buffer: batch * slots * pieces (mean, M2) pairs followed by `slots` element counts
inline long long igemm_gn_partial_floats(int batch, int slots, int pieces) { return 2ll * batch * slots * pieces + slots; }
int igemm_cout_pad(int cout);
// can a conv of this geometry apply ep.in_scale / in_shift to its A operand in shared memory (row-slab mainloop)?
bool igemm_xform_ok(int kind, int h_in, int w_in, int cin, int cout);
int igemm_block_n(int cout_pad);

}  // namespace clpk
