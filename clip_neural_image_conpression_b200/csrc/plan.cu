// plan.cu — plan-level entry points: the whole CLIPCondUNet forward (PKG/models/unet.py:81-106) and the DDIM loop
// (PKG/diffusion/ddim.py:21-46) as a fixed launch sequence over plan-owned device memory, captured into a CUDA graph.
//
// HBM layout (all plan-owned, sized for the plan's fixed batch B and image size):
//   X[l]  fp32 NHWC  residual stream of resolution level l (l = 0 .. n_levels); X[l] doubles as the skip tensor.  NOT
//         allocated for the levels whose stream is kept in 16 bits only (see res16_wanted / lv16 below: fp16 operands,
//         rows of >= 128 pixels) — there X16[l] IS the stream: read as a 16-bit residual tile, updated in place
//   Y     16-bit NHWC conv1 output of the current ResBlock (max size over levels); its GroupNorm statistics are taken
//         from the fp32 accumulators in the conv1 epilogue, only the stored copy is rounded
//   T     16-bit NHWC  GroupNorm+SiLU output = A operand of the ResBlock convs (max size)
//   X16[l] 16-bit NHWC: the stream itself (16-bit levels) or the copy of X[l] that is the A operand of the stride-2 /
//         transposed convs, written by the conv2 epilogue of the last ResBlock of a level.  (With env CLPK_X16=1 EVERY
//         producer of an fp32 X[l] also writes it and the GroupNorms on the residual stream read it instead of fp32 X:
//         same HBM bytes, but measured slower — the extra direct 16-bit stores cost the conv epilogues more than the
//         GroupNorm reads save.)
//   packed 16-bit weights [Cout][tap][Cin] per conv, fp32 bias / gamma / beta / Linear weights, FiLM weights of all
//   ResBlocks concatenated into one [2*sumC, time_dim] matrix so ONE small GEMV launch yields every (1+scale, shift).
#include "conv_igemm.cuh"
#include "kernels.cuh"

#include <map>
#include <stdlib.h>
#include <string>
#include <vector>

namespace clpk {

struct ConvPlan {
  int kind = 0, cin = 0, cout = 0, h_in = 0, w_in = 0;
  uint16_t* w = nullptr;  // packed 16-bit operand (fp16 or bf16)
  float* bias = nullptr;
  IgemmLaunch L;
  double flops = 0;
};

// One GroupNorm instance.  When its input is produced by a tcgen05 conv whose epilogue can emit per-tile partial sums
// (`fused`), the statistics pass over the tensor disappears: finalize(partial) -> stats -> apply.
struct GnPlan {
  int level = 0, silu = 1;
  float *gamma = nullptr, *beta = nullptr;
  float2* stats = nullptr;    // [B][groups] (mean, rstd)
  float2* partial = nullptr;  // [B][slots][pieces] (mean, M2) per tile + [slots] element counts, written by the producer conv
  int slots = 0;
  int piece = 0, pieces = 0;  // channels per partial sum the producer epilogue emits, and their number (C / piece)
  bool fused = false;
  // `in_consumer`: the normalisation (+ SiLU) is applied by the consuming conv to its own A operand in shared memory
  // (clpk_conv_epilogue.in_scale / in_shift); the stand-alone apply pass is replaced by a tiny partial -> (scale, shift)
  // fold (gn_affine_kernel)
  bool in_consumer = false;
  float *scale = nullptr, *shift = nullptr;  // [B][C]
};

struct ResBlockPlan {
  int c = 0, h = 0, w = 0, level = 0;
  GnPlan gn1, gn2;
  ConvPlan conv1, conv2;
  int film_off = 0;      // scale1p at [film_off, film_off+c), shift at [film_off+c, film_off+2c)
  bool emit16 = false;   // conv2 also writes the 16-bit copy X16[level] (feeds the next resampling conv)
  bool y16 = false;        // conv1 output kept in the 16-bit operand format (needs fused GroupNorm statistics)
};

// GroupNorm-in-consumer switches (env CLPK_FUSE_GN, bit mask, default 0): 1 = norm2 + SiLU of a ResBlock applied inside
// conv2 (row-slab levels), 2 = out_norm applied inside the `out` conv.  Correct (tests/test_gpu_unet.py,
// test_gpu_kernels.py) but OFF by default: measured on B200 the in-smem transform costs more than the stand-alone
// GroupNorm pass it removes — the slab mainloop already spends ~3/4 of the shared-memory bandwidth on operand reads
// (every A byte is read by 3 taps), and the extra read + write of each slab stretches a stage from ~770 to ~2400 clk
// (DESIGN.md section 9).
static int fuse_gn_mask() {
  const char* e = getenv("CLPK_FUSE_GN");
  return e ? atoi(e) : 0;
}
// env CLPK_RES16 (default 1): with fp16 operands the residual stream x <- x + conv2(...) (blocks.py:44, unet.py:104) of the
// WIDE resolution levels (rows of >= CLPK_RES16_MIN_W pixels, default 128: the row-slab levels, whose tensors are the
// HBM-bound ones) is kept in fp16 ONLY — conv2 / the transposed conv read the residual as a 16-bit tile and write the
// 16-bit sum (rounded once per block from the fp32 accumulator + residual), GroupNorm reads 2 B per element, and no fp32
// copy of that level's stream exists.  Narrower levels keep the fp32 stream: their tensors live in L2, the 16-bit form buys
// little there, and every rounding of the stream costs accuracy (on the 64 px config-1 net a 16-bit stream at every level
// costs 2 dB of final PSNR after 50 closed-loop steps).
static bool res16_wanted() {
  const char* e = getenv("CLPK_RES16");
  return !(e && atoi(e) == 0);
}
static int res16_min_width() {
  const char* e = getenv("CLPK_RES16_MIN_W");
  return e ? atoi(e) : 128;
}
// env CLPK_HEAD16 (default 1): the last transposed conv, whose result only out_norm reads, stores just the 16-bit copy
// (statistics still come from its fp32 accumulators) and out_norm reads 2 B instead of 4 B per element.
static bool head16_on() {
  const char* e = getenv("CLPK_HEAD16");
  return !(e && atoi(e) == 0);
}

}  // namespace clpk

using namespace clpk;

// makes `dev` current for the lifetime of the guard (plan calls run on the plan's device whatever the caller's is)
struct DevGuard {
  int prev = -1;
  explicit DevGuard(int dev) {
    if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
    if (prev != dev) cudaSetDevice(dev); else prev = -1;
  }
  ~DevGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

struct clpk_plan {
  clpk_unet_config cfg;
  int device = 0;            // CUDA device the plan's memory, tensor maps and graph belong to
  int B = 0, H = 0, W = 0;
  int n_levels = 0;
  std::vector<int> lv_c, lv_h, lv_w;  // size n_levels + 1
  std::vector<void*> allocs;
  long long bytes = 0;
  // parameters
  float *tp0_w = nullptr, *tp0_b = nullptr, *tp2_w = nullptr, *tp2_b = nullptr, *zp_w = nullptr, *zp_b = nullptr;
  ConvPlan stem;             // in_conv as im2col (K = 27 -> 32) + pointwise tcgen05 GEMM
  uint16_t* stem_cols = nullptr;  // [B,H,W,32] 16-bit im2col columns
  GnPlan out_gn;
  float *film_w = nullptr, *film_b = nullptr;
  int film_n = 0;  // 2 * sum of ResBlock channels
  std::vector<ResBlockPlan> rbs;     // execution order
  std::vector<ConvPlan> downs, ups;  // per level
  ConvPlan out_conv;
  // workspace
  std::vector<float*> X;
  uint16_t* Y = nullptr;     // conv1 output of the current ResBlock, stored in the 16-bit operand format
  uint16_t* T = nullptr;     // GroupNorm+SiLU output = conv A operand (fp16 or bf16, cfg.op_dtype)
  std::vector<uint16_t*> X16;  // per level: 16-bit copy of X[l]
  bool fuse_head = false;    // out_norm applied inside the `out` conv (see fuse_gn_mask)
  bool head16 = false;       // the last transposed conv writes only the 16-bit copy X16[0] (see head16_on)
  bool head_fused = false;   // out_norm + out conv as the single head_conv kernel (env CLPK_HEAD_FUSED, default 1)
  uint16_t* head_w = nullptr;  // [32][base] packed head weight (head_conv.cu)
  int head_stages = 0;       // experiments: cap of the head kernel's A ring (env CLPK_HEAD_STAGES at plan creation)
  bool head_ddim = false;    // DDIM steps: the head kernel also applies the update x <- f(x, eps) (env CLPK_HEAD_DDIM, default 1)
  bool x16_gn = false;       // env CLPK_X16=1: GroupNorms on the residual stream read X16 instead of fp32 X
  std::vector<char> lv16;    // per level: the residual stream lives in X16[l] ONLY (fp16), no fp32 X[l] (see res16_wanted)
  bool s16(int level) const { return lv16[level] != 0; }
  bool x16(int level) const { return x16_gn || lv16[level] != 0; }  // GroupNorms on the stream of this level read X16
  float* Yf = nullptr;       // fp32 conv1 output, only for ResBlocks whose GroupNorm statistics cannot be fused
  void* gn_ws = nullptr;
  float *temb = nullptr, *h1 = nullptr, *ht = nullptr, *hcond = nullptr, *film = nullptr, *zemb = nullptr;
  float* time_freqs = nullptr;  // optional host-evaluated timestep-embedding frequencies [time_dim / 2] (clpk_plan_set_time_freqs)
  int64_t* t_buf = nullptr;
  float* eps_buf = nullptr;  // NCHW eps of the current step
  float* x_buf = nullptr;    // NCHW DDIM state
  double flops_fwd = 0;
  int launches_fwd = 0;
  // DDIM state
  int steps = 0;
  int tab_cap = 0;           // steps ht_tab / coef_tab are sized for
  float *ht_tab = nullptr, *coef_tab = nullptr;
  DdimRun* run_dev = nullptr;
  bool any_sigma = false;
  cudaGraph_t graph = nullptr;
  cudaGraphExec_t graph_exec = nullptr;
  // The per-step conditioning (cond_combine + the FiLM GEMV of all ResBlocks, ~22 us) does not depend on the stem, so the
  // DDIM step forks it onto a side stream and joins before the first ResBlock (a parallel branch of the captured graph).
  cudaStream_t side_stream = nullptr;
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  cudaStream_t cap_stream = nullptr;  // private stream used only to capture the step graph (the caller's stream may be
                                      // the legacy default stream, which cannot be captured)
  // optional per-launch timing (clpk_plan_profile_forward): event pairs around every launch, tagged by class
  bool prof_on = false;
  std::vector<cudaEvent_t> prof_ev;
  std::vector<int> prof_cat;
  size_t prof_used = 0;
  void prof_mark(int cat, cudaStream_t s) {
    if (!prof_on) return;
    if (prof_used == prof_ev.size()) {
      cudaEvent_t e;
      cudaEventCreate(&e);
      prof_ev.push_back(e);
    }
    cudaEventRecord(prof_ev[prof_used++], s);
    prof_cat.push_back(cat);
  }

  template <typename Tp>
  int alloc(Tp** p, long long n_elems) {
    void* q = nullptr;
    const size_t nb = (size_t)std::max<long long>(n_elems, 1) * sizeof(Tp);
    cudaError_t e = cudaMalloc(&q, nb);
    if (e != cudaSuccess) {
      set_error("cudaMalloc(%zu bytes) failed: %s", nb, cudaGetErrorString(e));
      return CLPK_ERR_CUDA;
    }
    allocs.push_back(q);
    bytes += (long long)nb;
    *p = reinterpret_cast<Tp*>(q);
    return CLPK_OK;
  }
};

namespace {

struct ParamTable {
  std::map<std::string, std::pair<const float*, int64_t>> m;
  int get(const std::string& name, int64_t numel, const float** out) const {
    auto it = m.find(name);
    if (it == m.end()) {
      set_error("state dict is missing parameter '%s'", name.c_str());
      return CLPK_ERR_ARG;
    }
    if (it->second.second != numel) {
      set_error("parameter '%s' has %lld elements, expected %lld", name.c_str(), (long long)it->second.second,
                (long long)numel);
      return CLPK_ERR_ARG;
    }
    *out = it->second.first;
    return CLPK_OK;
  }
};

#define CLPK_TRY(expr)            \
  do {                            \
    int _rc = (expr);             \
    if (_rc != CLPK_OK) return _rc; \
  } while (0)

// launch `expr` on stream `s`, bracketed by profiling events of class `cat` when profiling is on
#define CLPK_TIMED(P, cat, s, expr) \
  do {                              \
    (P)->prof_mark((cat), (s));     \
    int _rc = (expr);               \
    (P)->prof_mark(-1, (s));        \
    if (_rc != CLPK_OK) return _rc; \
  } while (0)

enum ProfClass { kProfConvRes = 0, kProfConvOther = 1, kProfGroupNorm = 2, kProfConvIn = 3, kProfCond = 4, kProfDdim = 5 };

int copy_param(clpk_plan* P, const ParamTable& tab, const std::string& name, int64_t numel, float** dst) {
  const float* src = nullptr;
  CLPK_TRY(tab.get(name, numel, &src));
  CLPK_TRY(P->alloc(dst, numel));
  CLPK_CHECK_CUDA(cudaMemcpy(*dst, src, (size_t)numel * sizeof(float), cudaMemcpyDeviceToDevice));
  return CLPK_OK;
}

int make_conv(clpk_plan* P, const ParamTable& tab, const std::string& prefix, int kind, int cin, int cout, int h_in,
              int w_in, ConvPlan* cv) {
  cv->kind = kind; cv->cin = cin; cv->cout = cout; cv->h_in = h_in; cv->w_in = w_in;
  const int taps = (kind == CLPK_CONVT_4X4_S2) ? 16 : 9;
  const float* w = nullptr;
  CLPK_TRY(tab.get(prefix + ".weight", (int64_t)cin * cout * taps, &w));
  const int64_t n = clpk_pack_conv_weight(nullptr, nullptr, kind, cin, cout, P->cfg.op_dtype, nullptr);
  if (n < 0) return CLPK_ERR_ARG;
  CLPK_TRY(P->alloc(&cv->w, n));
  if (clpk_pack_conv_weight(w, cv->w, kind, cin, cout, P->cfg.op_dtype, nullptr) < 0) return CLPK_ERR_CUDA;
  // bias padded to the GEMM N (zeros beyond cout) so the vectorised epilogue may read whole 16-wide chunks
  const int cout_pad = igemm_cout_pad(cout);
  const float* b = nullptr;
  CLPK_TRY(tab.get(prefix + ".bias", cout, &b));
  CLPK_TRY(P->alloc(&cv->bias, cout_pad));
  CLPK_CHECK_CUDA(cudaMemset(cv->bias, 0, (size_t)cout_pad * sizeof(float)));
  CLPK_CHECK_CUDA(cudaMemcpy(cv->bias, b, (size_t)cout * sizeof(float), cudaMemcpyDeviceToDevice));
  const long long out_pix = (kind == CLPK_CONV_3X3_S2) ? (long long)(h_in / 2) * (w_in / 2) : (long long)h_in * w_in;
  cv->flops = 2.0 * P->B * (double)out_pix * cout * cin * (kind == CLPK_CONVT_4X4_S2 ? 16 : 9);
  return CLPK_OK;
}

int make_resblock(clpk_plan* P, const ParamTable& tab, const std::string& prefix, int level, ResBlockPlan* rb) {
  const int c = P->lv_c[level], h = P->lv_h[level], w = P->lv_w[level], td = P->cfg.time_dim;
  rb->c = c; rb->h = h; rb->w = w; rb->level = level;
  rb->gn1.level = rb->gn2.level = level;
  CLPK_TRY(copy_param(P, tab, prefix + ".norm1.weight", c, &rb->gn1.gamma));
  CLPK_TRY(copy_param(P, tab, prefix + ".norm1.bias", c, &rb->gn1.beta));
  CLPK_TRY(copy_param(P, tab, prefix + ".norm2.weight", c, &rb->gn2.gamma));
  CLPK_TRY(copy_param(P, tab, prefix + ".norm2.bias", c, &rb->gn2.beta));
  CLPK_TRY(make_conv(P, tab, prefix + ".conv1", CLPK_CONV_3X3_S1, c, c, h, w, &rb->conv1));
  CLPK_TRY(make_conv(P, tab, prefix + ".conv2", CLPK_CONV_3X3_S1, c, c, h, w, &rb->conv2));
  // FiLM rows into the concatenated matrix: [scale rows | shift rows]; "+1" folded into the scale bias
  const float *sw, *sb, *hw_, *hb;
  CLPK_TRY(tab.get(prefix + ".film.to_scale.weight", (int64_t)c * td, &sw));
  CLPK_TRY(tab.get(prefix + ".film.to_scale.bias", c, &sb));
  CLPK_TRY(tab.get(prefix + ".film.to_shift.weight", (int64_t)c * td, &hw_));
  CLPK_TRY(tab.get(prefix + ".film.to_shift.bias", c, &hb));
  const size_t row = (size_t)td * sizeof(float);
  CLPK_CHECK_CUDA(cudaMemcpy(P->film_w + (size_t)rb->film_off * td, sw, row * c, cudaMemcpyDeviceToDevice));
  CLPK_CHECK_CUDA(cudaMemcpy(P->film_w + (size_t)(rb->film_off + c) * td, hw_, row * c, cudaMemcpyDeviceToDevice));
  CLPK_CHECK_CUDA(cudaMemcpy(P->film_b + rb->film_off, sb, (size_t)c * sizeof(float), cudaMemcpyDeviceToDevice));
  CLPK_CHECK_CUDA(cudaMemcpy(P->film_b + rb->film_off + c, hb, (size_t)c * sizeof(float), cudaMemcpyDeviceToDevice));
  CLPK_TRY(launch_add_const(P->film_b + rb->film_off, 1.0f, c, nullptr));
  return CLPK_OK;
}

// bind the buffers of a conv (A operand, epilogue) and encode its tensor maps
int bind_conv(clpk_plan* P, ConvPlan* cv, const void* a, const clpk_conv_epilogue& ep) {
  return igemm_setup(a, cv->w, cv->kind, P->B, cv->h_in, cv->w_in, cv->cin, cv->cout, P->cfg.op_dtype, &ep, &cv->L);
}

int gn_groups_of(const clpk_plan* P, int level) { return std::min(P->cfg.groups, P->lv_c[level]); }

// allocate the statistics buffers of a GroupNorm whose input comes from a conv of (kind, h_in, w_in) -> cout channels
int setup_gn(clpk_plan* P, GnPlan* gn, int producer_kind, int prod_h_in, int prod_w_in) {
  const int c = P->lv_c[gn->level], groups = gn_groups_of(P, gn->level);
  CLPK_TRY(P->alloc(&gn->stats, (long long)P->B * groups));
  // The conv epilogue sums 4 / 8 / 16 channels or whole multiples of 32 at a time.  A group of another size (24, 48:
  // base = 192) is summed in pieces of the largest of 16 / 8 / 4 that divides it; the apply kernel folds the pieces.
  const int cpg = c / groups;
  int piece = cpg;
  if (!(cpg == 4 || cpg == 8 || cpg == 16 || cpg % 32 == 0)) piece = (cpg % 16 == 0) ? 16 : (cpg % 8 == 0) ? 8 : (cpg % 4 == 0) ? 4 : 0;
  gn->piece = piece;
  gn->pieces = piece > 0 ? c / piece : 0;
  gn->slots = (producer_kind >= 0 && piece > 0) ? igemm_gn_slots(producer_kind, prod_h_in, prod_w_in, c, piece) : 0;
  gn->fused = gn->slots > 0;
  if (gn->fused) {
    float* buf = nullptr;
    CLPK_TRY(P->alloc(&buf, igemm_gn_partial_floats(P->B, gn->slots, gn->pieces) + 2));
    gn->partial = reinterpret_cast<float2*>(buf);
  }
  return CLPK_OK;
}

int run_groupnorm(clpk_plan* P, const void* x, int x_is_16, const GnPlan& gn, cudaStream_t s) {
  const int level = gn.level, groups = gn_groups_of(P, level);
  const int hw = P->lv_h[level] * P->lv_w[level], c = P->lv_c[level];
  if (gn.in_consumer) {  // statistics -> (scale, shift) table; the consuming conv normalises its own operand
    P->prof_mark(kProfCond, s);
    const int rc = launch_gn_affine(gn.partial, gn.gamma, gn.beta, gn.scale, gn.shift, P->B, gn.slots, gn.pieces, groups, c,
                                    1e-5f, s);
    P->prof_mark(-1, s);
    return rc;
  }
  const GnShape shp = gn_shape(P->B, hw, c, groups);
  P->prof_mark(kProfGroupNorm, s);
  int rc;
  if (gn.fused) {
    // statistics come from the producing conv's epilogue partials and are folded inside the apply kernel
    rc = launch_gn_apply_ex(x, x_is_16, gn.gamma, gn.beta, nullptr, gn.partial, gn.slots, gn.pieces, (double)hw * (c / groups), 1e-5f,
                            P->T, shp, gn.silu, P->cfg.op_dtype, s);
  } else {
    const float2* stats = nullptr;
    rc = x_is_16 ? CLPK_ERR_STATE : launch_gn_stats(reinterpret_cast<const float*>(x), P->gn_ws, shp, 1e-5f, &stats, s);
    if (rc == CLPK_OK)
      rc = launch_gn_apply_ex(x, 0, gn.gamma, gn.beta, stats, nullptr, 0, groups, 1.0, 0.f, P->T, shp, gn.silu, P->cfg.op_dtype, s);
  }
  P->prof_mark(-1, s);
  return rc;
}

// GroupNorm over the residual stream of `level`: the 16-bit copy when its statistics are fused, else fp32 X
int run_groupnorm_x(clpk_plan* P, int level, const GnPlan& gn, cudaStream_t s) {
  if (gn.fused && P->x16(level)) return run_groupnorm(P, P->X16[level], 1, gn, s);
  return run_groupnorm(P, P->X[level], 0, gn, s);
}

int run_resblock(clpk_plan* P, ResBlockPlan& rb, cudaStream_t s) {
  CLPK_TRY(run_groupnorm_x(P, rb.level, rb.gn1, s));               // blocks.py:41 act(norm1(x))
  CLPK_TIMED(P, kProfConvRes, s, igemm_launch(rb.conv1.L, s));     // conv1 + FiLM -> Y   (blocks.py:41-42)
  CLPK_TRY(run_groupnorm(P, rb.y16 ? (const void*)P->Y : (const void*)P->Yf, rb.y16 ? 1 : 0, rb.gn2, s));  // blocks.py:43
  CLPK_TIMED(P, kProfConvRes, s, igemm_launch(rb.conv2.L, s));     // conv2 + x -> X      (blocks.py:43-44)
  return CLPK_OK;
}

// everything after the conditioning vector: in_conv ... out   (unet.py:88-105).  film = [B, film_n].
// ddim_x != nullptr (DDIM steps with the fused head kernel): the head also updates ddim_x in place (ddim.py:36-45)
int forward_body(clpk_plan* P, const float* x_nchw, cudaStream_t s, cudaEvent_t cond_ready = nullptr, float* ddim_x = nullptr) {
  const clpk_unet_config& c = P->cfg;
  P->prof_mark(kProfConvIn, s);  // stem (unet.py:88): im2col columns + pointwise GEMM on the tensor cores
  int src = launch_stem_im2col(x_nchw, P->stem_cols, P->B, c.img_ch, P->H, P->W, c.op_dtype, s);
  if (src == CLPK_OK) src = igemm_launch(P->stem.L, s);
  P->prof_mark(-1, s);
  if (src != CLPK_OK) return src;
  if (cond_ready) CLPK_CHECK_CUDA(cudaStreamWaitEvent(s, cond_ready, 0));  // FiLM table (side stream) before the first conv1
  size_t r = 0;
  for (int l = 0; l < P->n_levels; ++l) {
    CLPK_TRY(run_resblock(P, P->rbs[r++], s));
    CLPK_TRY(run_resblock(P, P->rbs[r++], s));
    CLPK_TIMED(P, kProfConvOther, s, igemm_launch(P->downs[l].L, s));  // X16[l] -> X[l+1], X16[l+1]
  }
  CLPK_TRY(run_resblock(P, P->rbs[r++], s));
  CLPK_TRY(run_resblock(P, P->rbs[r++], s));
  for (int l = P->n_levels - 1; l >= 0; --l) {
    CLPK_TRY(run_resblock(P, P->rbs[r++], s));
    CLPK_TRY(run_resblock(P, P->rbs[r++], s));
    CLPK_TIMED(P, kProfConvOther, s, igemm_launch(P->ups[l].L, s));  // X16[l+1] -> X[l] += convT (unet.py:102-104)
  }
  if (P->head16) CLPK_TRY(run_groupnorm(P, P->X16[0], 1, P->out_gn, s));  // out_norm on the 16-bit transposed-conv output
  else CLPK_TRY(run_groupnorm_x(P, 0, P->out_gn, s));              // out_norm, no activation (unet.py:105)
  if (P->head_fused) {
    CLPK_TIMED(P, kProfConvOther, s,
               launch_head_conv(P->X16[0], P->out_gn.scale, P->out_gn.shift, P->head_w, P->out_conv.bias, P->eps_buf, P->B,
                                P->H, P->W, c.base, c.op_dtype, s, P->head_stages, ddim_x, ddim_x ? P->coef_tab : nullptr,
                                ddim_x ? P->run_dev : nullptr));  // out(out_norm(x)) -> eps_buf (NCHW) [+ DDIM update]
  } else {
    CLPK_TIMED(P, kProfConvOther, s, igemm_launch(P->out_conv.L, s));  // -> eps_buf (NCHW)
  }
  return CLPK_OK;
}

// film[B, film_n] = Linear_film(h) ; h = time_proj(temb(t)) + z_proj(z)   (unet.py:83-86, blocks.py:22-24)
int film_from_h(clpk_plan* P, cudaStream_t s) {
  CLPK_TIMED(P, kProfCond, s,
             launch_linear(P->hcond, P->film_w, P->film_b, nullptr, 0, P->film, P->B, P->film_n, P->cfg.time_dim, 0, s));
  return CLPK_OK;
}

}  // namespace

extern "C" void clpk_plan_destroy(clpk_plan* P) {
  if (!P) return;
  DevGuard dg(P->device);
  if (P->graph_exec) cudaGraphExecDestroy(P->graph_exec);
  if (P->graph) cudaGraphDestroy(P->graph);
  if (P->cap_stream) cudaStreamDestroy(P->cap_stream);
  if (P->side_stream) cudaStreamDestroy(P->side_stream);
  if (P->ev_fork) cudaEventDestroy(P->ev_fork);
  if (P->ev_join) cudaEventDestroy(P->ev_join);
  for (cudaEvent_t e : P->prof_ev) cudaEventDestroy(e);
  for (void* q : P->allocs) cudaFree(q);
  if (P->ht_tab) cudaFree(P->ht_tab);
  if (P->coef_tab) cudaFree(P->coef_tab);
  delete P;
}

extern "C" int clpk_plan_create(const clpk_unet_config* cfg, int batch, int height, int width, int n_params,
                                const char* const* names, const float* const* ptrs, const int64_t* numels,
                                clpk_plan** out_plan) {
  CLPK_REQUIRE(cfg && out_plan && names && ptrs && numels, "clpk_plan_create: null argument");
  CLPK_REQUIRE(batch > 0 && height > 0 && width > 0, "clpk_plan_create: bad batch/size");
  CLPK_REQUIRE(cfg->n_levels >= 1 && cfg->n_levels <= CLPK_MAX_LEVELS, "clpk_plan_create: bad n_levels");
  CLPK_REQUIRE(cfg->base % 32 == 0, "base channels must be a multiple of 32 (got %d)", cfg->base);
  CLPK_REQUIRE(cfg->op_dtype == CLPK_OP_BF16 || cfg->op_dtype == CLPK_OP_F16, "bad op_dtype %d", cfg->op_dtype);
  CLPK_REQUIRE(height % (1 << cfg->n_levels) == 0 && width % (1 << cfg->n_levels) == 0,
               "H, W must be divisible by 2^len(ch_mult)");
  int dev_count = 0;
  CLPK_CHECK_CUDA(cudaGetDeviceCount(&dev_count));
  CLPK_REQUIRE(dev_count > 0, "no CUDA device");
  CLPK_TRY(igemm_init());

  ParamTable tab;
  for (int i = 0; i < n_params; ++i) tab.m[names[i]] = {ptrs[i], numels[i]};

  clpk_plan* P = new clpk_plan();
  struct Guard {
    clpk_plan* p;
    ~Guard() { if (p) clpk_plan_destroy(p); }
  } guard{P};
  P->cfg = *cfg;
  CLPK_CHECK_CUDA(cudaGetDevice(&P->device));
  P->B = batch; P->H = height; P->W = width;
  P->n_levels = cfg->n_levels;
  const int td = cfg->time_dim, L = cfg->n_levels;
  int ch = cfg->base, hh = height, ww = width;
  for (int l = 0; l <= L; ++l) {
    P->lv_c.push_back(ch); P->lv_h.push_back(hh); P->lv_w.push_back(ww);
    if (l < L) {
      CLPK_REQUIRE(cfg->ch_mult[l] >= 1, "bad ch_mult");
      ch *= cfg->ch_mult[l]; hh /= 2; ww /= 2;
    }
  }
  // ---- conditioning MLPs (unet.py:47-53)
  CLPK_TRY(copy_param(P, tab, "time_proj.0.weight", (int64_t)4 * td * td, &P->tp0_w));
  CLPK_TRY(copy_param(P, tab, "time_proj.0.bias", 4 * td, &P->tp0_b));
  CLPK_TRY(copy_param(P, tab, "time_proj.2.weight", (int64_t)4 * td * td, &P->tp2_w));
  CLPK_TRY(copy_param(P, tab, "time_proj.2.bias", td, &P->tp2_b));
  CLPK_TRY(copy_param(P, tab, "z_proj.0.weight", (int64_t)td * cfg->z_dim, &P->zp_w));
  CLPK_TRY(copy_param(P, tab, "z_proj.0.bias", td, &P->zp_b));
  CLPK_REQUIRE(cfg->img_ch * 9 <= 27, "img_ch > 3 is not supported by the stem im2col");
  P->out_gn.level = 0;
  P->out_gn.silu = 0;
  CLPK_TRY(copy_param(P, tab, "out_norm.weight", cfg->base, &P->out_gn.gamma));
  CLPK_TRY(copy_param(P, tab, "out_norm.bias", cfg->base, &P->out_gn.beta));

  // ---- ResBlock list in execution order + FiLM offsets
  struct RbSpec { std::string prefix; int level; bool emit; };
  std::vector<RbSpec> specs;
  for (int l = 0; l < L; ++l) {
    specs.push_back({"down." + std::to_string(3 * l), l, false});
    specs.push_back({"down." + std::to_string(3 * l + 1), l, true});
  }
  specs.push_back({"mid1", L, false});
  specs.push_back({"mid2", L, false});
  for (int i = 0; i < L; ++i) {
    const int l = L - i;  // ResBlocks of up stage i run at level l = L - i... (coarse to fine)
    specs.push_back({"up." + std::to_string(3 * i), l, false});
    specs.push_back({"up." + std::to_string(3 * i + 1), l, true});
  }
  int film_n = 0;
  for (auto& sp : specs) film_n += 2 * P->lv_c[sp.level];
  P->film_n = film_n;
  CLPK_TRY(P->alloc(&P->film_w, (long long)film_n * td));
  CLPK_TRY(P->alloc(&P->film_b, film_n));

  // ---- workspace
  long long max_act = 0;
  for (int l = 0; l <= L; ++l) {
    const long long n = (long long)batch * P->lv_h[l] * P->lv_w[l] * P->lv_c[l];
    max_act = std::max(max_act, n);
    P->X.push_back(nullptr);  // fp32 stream: allocated below unless the plan keeps the stream in 16 bits only
    uint16_t* x16 = nullptr;
    CLPK_TRY(P->alloc(&x16, n));
    P->X16.push_back(x16);
  }
  { const char* e = getenv("CLPK_X16"); P->x16_gn = e && atoi(e) != 0; }
  P->lv16.assign(L + 1, 0);
  CLPK_TRY(P->alloc(&P->Y, max_act));
  CLPK_TRY(P->alloc(&P->Yf, max_act));
  CLPK_TRY(P->alloc(&P->T, max_act));
  long long gn_bytes = 0;
  for (int l = 0; l <= L; ++l)
    gn_bytes = std::max(gn_bytes, gn_ws_bytes(gn_shape(batch, P->lv_h[l] * P->lv_w[l], P->lv_c[l],
                                                       std::min(cfg->groups, P->lv_c[l]))));
  {
    char* ws = nullptr;
    CLPK_TRY(P->alloc(&ws, gn_bytes));
    CLPK_CHECK_CUDA(cudaMemset(ws, 0, (size_t)gn_bytes));
    P->gn_ws = ws;
  }
  CLPK_TRY(P->alloc(&P->temb, (long long)batch * td));
  CLPK_TRY(P->alloc(&P->h1, (long long)batch * 4 * td));
  CLPK_TRY(P->alloc(&P->ht, (long long)batch * td));
  CLPK_TRY(P->alloc(&P->hcond, (long long)batch * td));
  CLPK_TRY(P->alloc(&P->zemb, (long long)batch * td));
  CLPK_TRY(P->alloc(&P->film, (long long)batch * film_n));
  CLPK_TRY(P->alloc(&P->t_buf, batch));
  const long long img_elems = (long long)batch * cfg->img_ch * height * width;
  CLPK_TRY(P->alloc(&P->eps_buf, img_elems));
  CLPK_TRY(P->alloc(&P->x_buf, img_elems));
  CLPK_TRY(P->alloc(&P->run_dev, 1));
  CLPK_CHECK_CUDA(cudaMemset(P->run_dev, 0, sizeof(DdimRun)));

  // ---- ResBlocks: parameters first, then the GroupNorm statistics wiring (who produces each GN's input)
  P->rbs.resize(specs.size());
  int off = 0;
  for (size_t i = 0; i < specs.size(); ++i) {
    ResBlockPlan& rb = P->rbs[i];
    rb.film_off = off;
    rb.emit16 = specs[i].emit || P->x16_gn;
    CLPK_TRY(make_resblock(P, tab, specs[i].prefix, specs[i].level, &rb));
    off += 2 * rb.c;
  }
  std::vector<GnPlan*> gn_after_conv2(specs.size(), nullptr), gn_after_down(L, nullptr), gn_after_up(L, nullptr);
  for (size_t i = 0; i < specs.size(); ++i) {
    ResBlockPlan& rb = P->rbs[i];
    const int l = rb.level;
    CLPK_TRY(setup_gn(P, &rb.gn2, CLPK_CONV_3X3_S1, rb.h, rb.w));  // input = this block's conv1 output
    if (i == 0) {
      CLPK_TRY(setup_gn(P, &rb.gn1, CLPK_CONV_1X1, rb.h, rb.w));   // input = stem conv (pointwise tcgen05 GEMM)
    } else if (specs[i - 1].level == l) {
      CLPK_TRY(setup_gn(P, &rb.gn1, CLPK_CONV_3X3_S1, rb.h, rb.w));
      gn_after_conv2[i - 1] = &rb.gn1;
    } else if (specs[i - 1].level < l) {                            // came down: stride-2 conv from level l-1
      CLPK_TRY(setup_gn(P, &rb.gn1, CLPK_CONV_3X3_S2, P->lv_h[l - 1], P->lv_w[l - 1]));
      gn_after_down[l - 1] = &rb.gn1;
    } else {                                                        // came up: transposed conv from level l+1
      CLPK_TRY(setup_gn(P, &rb.gn1, CLPK_CONVT_4X4_S2, P->lv_h[l + 1], P->lv_w[l + 1]));
      gn_after_up[l] = &rb.gn1;
    }
  }
  CLPK_TRY(setup_gn(P, &P->out_gn, CLPK_CONVT_4X4_S2, P->lv_h[1], P->lv_w[1]));
  gn_after_up[0] = &P->out_gn;
  {
    // 16-bit-only residual stream per level (res16_wanted): every GroupNorm on that level's stream must take its statistics
    // from the producing conv's fp32 accumulators (a statistics pass over the rounded copy would be slower and less exact).
    // fp16 operands only: bf16's 8-bit mantissa is too coarse for a running sum.
    std::vector<char> fused(L + 1, 1);
    fused[0] = P->out_gn.fused ? 1 : 0;
    for (const ResBlockPlan& rb : P->rbs)
      if (!rb.gn1.fused) fused[rb.level] = 0;
    for (int l = 0; l <= L; ++l) {
      P->lv16[l] = res16_wanted() && cfg->op_dtype == CLPK_OP_F16 && fused[l] && !P->x16_gn && P->lv_w[l] >= res16_min_width();
      if (!P->lv16[l]) CLPK_TRY(P->alloc(&P->X[l], (long long)batch * P->lv_h[l] * P->lv_w[l] * P->lv_c[l]));
    }
    for (ResBlockPlan& rb : P->rbs)
      if (P->s16(rb.level)) rb.emit16 = true;
  }
  P->fuse_head = (fuse_gn_mask() & 2) && P->out_gn.fused && !P->x16_gn &&
                 igemm_xform_ok(CLPK_CONV_3X3_S1, height, width, cfg->base, cfg->img_ch);
  P->head16 = P->fuse_head || P->s16(0) || (head16_on() && P->out_gn.fused && !P->x16_gn && cfg->base % 32 == 0);
  {
    const char* e = getenv("CLPK_HEAD_FUSED");
    P->head_fused = !(e && atoi(e) == 0) && !P->fuse_head && P->head16 &&
                    head_conv_supported(height, width, cfg->base, cfg->img_ch);
    const char* hs = getenv("CLPK_HEAD_STAGES");
    P->head_stages = hs ? atoi(hs) : 0;
    const char* hd = getenv("CLPK_HEAD_DDIM");
    P->head_ddim = !(hd && atoi(hd) == 0);
  }
  if (P->head_fused) {  // out_norm is applied inside head_conv: only the (scale, shift) table is computed
    P->out_gn.in_consumer = true;
    CLPK_TRY(P->alloc(&P->out_gn.scale, (long long)batch * cfg->base));
    CLPK_TRY(P->alloc(&P->out_gn.shift, (long long)batch * cfg->base));
  }
  if (P->fuse_head) {
    P->out_gn.in_consumer = true;
    CLPK_TRY(P->alloc(&P->out_gn.scale, (long long)batch * cfg->base));
    CLPK_TRY(P->alloc(&P->out_gn.shift, (long long)batch * cfg->base));
  }
  auto wire_gn = [&](clpk_conv_epilogue* e, const GnPlan* gn) {
    if (gn && gn->fused) {
      e->gn_partial = gn->partial;
      e->gn_cpg = gn->piece;  // channels per partial sum (== channels per group unless the group is summed in pieces)
    }
  };
  for (size_t i = 0; i < specs.size(); ++i) {
    ResBlockPlan& rb = P->rbs[i];
    clpk_conv_epilogue e1{};
    e1.bias = rb.conv1.bias;
    e1.film_scale1p = P->film + rb.film_off;
    e1.film_shift = P->film + rb.film_off + rb.c;
    e1.film_stride = film_n;
    { const char* e = getenv("CLPK_Y_FP32"); rb.y16 = rb.gn2.fused && !(e && atoi(e) != 0); }
    if (rb.y16) e1.out_op = P->Y; else e1.out_f32 = P->Yf;
    e1.cout_valid = rb.c;
    wire_gn(&e1, &rb.gn2);
    CLPK_TRY(bind_conv(P, &rb.conv1, P->T, e1));
    clpk_conv_epilogue e2{};
    e2.bias = rb.conv2.bias;
    if (P->s16(rb.level)) {   // blocks.py:44 on the 16-bit stream, in place
      e2.resid_op = P->X16[rb.level];
      e2.out_op = P->X16[rb.level];
    } else {
      e2.resid = P->X[rb.level];
      e2.out_f32 = P->X[rb.level];
      e2.out_op = rb.emit16 ? P->X16[rb.level] : nullptr;
    }
    e2.cout_valid = rb.c;
    wire_gn(&e2, gn_after_conv2[i]);
    const void* a2 = P->T;
    if ((fuse_gn_mask() & 1) && rb.y16 && rb.gn2.fused && igemm_xform_ok(CLPK_CONV_3X3_S1, rb.h, rb.w, rb.c, rb.c)) {
      // blocks.py:43 act(norm2(y)) happens inside conv2: A operand = the raw conv1 + FiLM output Y
      rb.gn2.in_consumer = true;
      CLPK_TRY(P->alloc(&rb.gn2.scale, (long long)batch * rb.c));
      CLPK_TRY(P->alloc(&rb.gn2.shift, (long long)batch * rb.c));
      e2.in_scale = rb.gn2.scale;
      e2.in_shift = rb.gn2.shift;
      e2.in_silu = 1;
      a2 = P->Y;
    }
    CLPK_TRY(bind_conv(P, &rb.conv2, a2, e2));
    P->flops_fwd += rb.conv1.flops + rb.conv2.flops;
  }
  // ---- resampling convs
  P->downs.resize(L);
  P->ups.resize(L);
  for (int l = 0; l < L; ++l) {
    ConvPlan& dn = P->downs[l];
    CLPK_TRY(make_conv(P, tab, "down." + std::to_string(3 * l + 2), CLPK_CONV_3X3_S2, P->lv_c[l], P->lv_c[l + 1],
                       P->lv_h[l], P->lv_w[l], &dn));
    clpk_conv_epilogue ed{};
    ed.bias = dn.bias;
    ed.out_f32 = P->s16(l + 1) ? nullptr : P->X[l + 1];
    ed.out_op = P->x16(l + 1) ? P->X16[l + 1] : nullptr;
    ed.cout_valid = P->lv_c[l + 1];
    wire_gn(&ed, gn_after_down[l]);
    CLPK_TRY(bind_conv(P, &dn, P->X16[l], ed));
    P->flops_fwd += dn.flops;
    // up stage i = L-1-l maps level l+1 -> l: module up.(3*i+2), ConvTranspose2d(C[l+1] -> C[l])
    const int i = L - 1 - l;
    ConvPlan& up = P->ups[l];
    CLPK_TRY(make_conv(P, tab, "up." + std::to_string(3 * i + 2), CLPK_CONVT_4X4_S2, P->lv_c[l + 1], P->lv_c[l],
                       P->lv_h[l + 1], P->lv_w[l + 1], &up));
    clpk_conv_epilogue eu{};
    eu.bias = up.bias;
    eu.resid = P->X[l];  // skip connection, added in place
    eu.out_f32 = P->X[l];
    eu.out_op = P->x16(l) ? P->X16[l] : nullptr;
    if (P->s16(l)) {
      eu.resid = nullptr;
      eu.resid_op = P->X16[l];
      eu.out_f32 = nullptr;
    } else if (l == 0 && P->head16) {
      // the last transposed conv's result is read by out_norm only: keep just the 16-bit copy (its GroupNorm statistics
      // still come from the fp32 accumulators), which the `out` conv normalises in shared memory
      eu.out_f32 = nullptr;
      eu.out_op = P->X16[0];
    }
    eu.cout_valid = P->lv_c[l];
    wire_gn(&eu, gn_after_up[l]);
    CLPK_TRY(bind_conv(P, &up, P->X16[l + 1], eu));
    P->flops_fwd += up.flops;
  }
  // ---- stem: in_conv.weight [base, img_ch, 3, 3] = [base][27] -> zero-padded [base][32] -> packed pointwise weight
  {
    ConvPlan& st = P->stem;
    st.kind = CLPK_CONV_1X1; st.cin = 32; st.cout = cfg->base; st.h_in = height; st.w_in = width;
    const int k_src = cfg->img_ch * 9;
    const float *w = nullptr, *b = nullptr;
    CLPK_TRY(tab.get("in_conv.weight", (int64_t)cfg->base * k_src, &w));
    CLPK_TRY(tab.get("in_conv.bias", cfg->base, &b));
    float* wpad = nullptr;
    CLPK_TRY(P->alloc(&wpad, (long long)cfg->base * 32));
    CLPK_TRY(launch_pad_rows(w, wpad, cfg->base, k_src, 32, nullptr));
    const int64_t n = clpk_pack_conv_weight(nullptr, nullptr, CLPK_CONV_1X1, 32, cfg->base, cfg->op_dtype, nullptr);
    CLPK_TRY(P->alloc(&st.w, n));
    if (clpk_pack_conv_weight(wpad, st.w, CLPK_CONV_1X1, 32, cfg->base, cfg->op_dtype, nullptr) < 0) return CLPK_ERR_CUDA;
    CLPK_TRY(P->alloc(&st.bias, igemm_cout_pad(cfg->base)));
    CLPK_CHECK_CUDA(cudaMemset(st.bias, 0, (size_t)igemm_cout_pad(cfg->base) * sizeof(float)));
    CLPK_CHECK_CUDA(cudaMemcpy(st.bias, b, (size_t)cfg->base * sizeof(float), cudaMemcpyDeviceToDevice));
    CLPK_TRY(P->alloc(&P->stem_cols, (long long)batch * height * width * 32));
    clpk_conv_epilogue es{};
    es.bias = st.bias;
    es.out_f32 = P->s16(0) ? nullptr : P->X[0];
    es.out_op = P->x16(0) ? P->X16[0] : nullptr;
    es.cout_valid = cfg->base;
    wire_gn(&es, &P->rbs[0].gn1);
    CLPK_TRY(bind_conv(P, &st, P->stem_cols, es));
  }
  // ---- head
  CLPK_TRY(make_conv(P, tab, "out", CLPK_CONV_3X3_S1, cfg->base, cfg->img_ch, height, width, &P->out_conv));
  {
    clpk_conv_epilogue eo{};
    eo.bias = P->out_conv.bias;
    eo.out_nchw = P->eps_buf;
    eo.cout_valid = cfg->img_ch;
    if (P->fuse_head) {  // unet.py:105 out(out_norm(x)): no activation
      eo.in_scale = P->out_gn.scale;
      eo.in_shift = P->out_gn.shift;
      eo.in_silu = 0;
    }
    CLPK_TRY(bind_conv(P, &P->out_conv, P->fuse_head ? (const void*)P->X16[0] : (const void*)P->T, eo));
    P->flops_fwd += P->out_conv.flops;
    if (P->head_fused) {
      const float* w = nullptr;
      CLPK_TRY(tab.get("out.weight", (int64_t)cfg->img_ch * cfg->base * 9, &w));
      CLPK_TRY(P->alloc(&P->head_w, 32ll * cfg->base));
      CLPK_TRY(launch_pack_head_weight(w, P->head_w, cfg->base, cfg->op_dtype, nullptr));
    }
  }
  P->flops_fwd += 2.0 * batch * (double)height * width * cfg->base * cfg->img_ch * 9;  // in_conv
  P->flops_fwd += 2.0 * batch * ((double)td * 4 * td * 2 + (double)cfg->z_dim * td + (double)film_n * td);
  // launches per forward: conditioning (5) + stem (2) + per ResBlock (2 GN x 2 + 2 conv) + resamplers + out_norm(2) + out
  P->launches_fwd = 5 + 2 + (int)P->rbs.size() * 2 + 2 * L + 1;  // + one or two launches per GroupNorm:
  for (const ResBlockPlan& rb : P->rbs) P->launches_fwd += (rb.gn1.fused ? 1 : 2) + (rb.gn2.fused ? 1 : 2);
  P->launches_fwd += P->out_gn.fused ? 1 : 2;
  CLPK_CHECK_CUDA(cudaStreamCreateWithFlags(&P->side_stream, cudaStreamNonBlocking));
  CLPK_CHECK_CUDA(cudaEventCreateWithFlags(&P->ev_fork, cudaEventDisableTiming));
  CLPK_CHECK_CUDA(cudaEventCreateWithFlags(&P->ev_join, cudaEventDisableTiming));
  CLPK_CHECK_CUDA(cudaDeviceSynchronize());
  guard.p = nullptr;
  *out_plan = P;
  return CLPK_OK;
}

extern "C" int clpk_plan_set_time_freqs(clpk_plan* P, const float* freqs_dev) {
  CLPK_REQUIRE(P && freqs_dev, "clpk_plan_set_time_freqs: null argument");
  DevGuard dg(P->device);
  const int half = P->cfg.time_dim / 2;
  if (!P->time_freqs) CLPK_TRY(P->alloc(&P->time_freqs, half));
  CLPK_CHECK_CUDA(cudaMemcpy(P->time_freqs, freqs_dev, (size_t)half * sizeof(float), cudaMemcpyDeviceToDevice));
  CLPK_CHECK_CUDA(cudaStreamSynchronize(nullptr));  // device-to-device copies do not block the host: the table must be in
                                                    // place before work on any (non-blocking) stream reads it
  P->steps = 0;   // a prepared DDIM loop holds time_proj(temb(t_i)) of the old table: prepare again
  return CLPK_OK;
}

extern "C" int64_t clpk_plan_device_bytes(const clpk_plan* P) { return P ? P->bytes : 0; }
extern "C" double clpk_plan_flops_per_forward(const clpk_plan* P) { return P ? P->flops_fwd : 0.0; }
extern "C" int clpk_plan_launches_per_forward(const clpk_plan* P) { return P ? P->launches_fwd : 0; }

extern "C" int clpk_unet_forward(clpk_plan* P, const float* x, const float* z, const int64_t* t, float* eps,
                                 void* stream) {
  CLPK_REQUIRE(P && x && z && t && eps, "clpk_unet_forward: null argument");
  DevGuard dg(P->device);
  cudaStream_t s = (cudaStream_t)stream;
  const clpk_unet_config& c = P->cfg;
  const int td = c.time_dim;
  CLPK_TRY(launch_timestep_embedding(t, P->temb, P->B, td, 10000.f, s, P->time_freqs));              // unet.py:83
  CLPK_TRY(launch_linear(P->temb, P->tp0_w, P->tp0_b, nullptr, 0, P->h1, P->B, 4 * td, td, 1, s));   // :84 Linear+SiLU
  CLPK_TRY(launch_linear(P->h1, P->tp2_w, P->tp2_b, nullptr, 0, P->ht, P->B, td, 4 * td, 0, s));     // :84 Linear
  CLPK_TRY(launch_linear(z, P->zp_w, P->zp_b, P->ht, P->B, P->hcond, P->B, td, c.z_dim, 1, s));      // :85-86
  CLPK_TRY(film_from_h(P, s));
  CLPK_TRY(forward_body(P, x, s));
  const size_t nb = (size_t)P->B * c.img_ch * P->H * P->W * sizeof(float);
  CLPK_CHECK_CUDA(cudaMemcpyAsync(eps, P->eps_buf, nb, cudaMemcpyDeviceToDevice, s));
  return CLPK_OK;
}

// one DDIM step on plan-owned buffers: conditioning for run->step, eps = UNet(x_buf), x_buf <- update, step += 1
static int ddim_step_body(clpk_plan* P, cudaStream_t s) {
  const int td = P->cfg.time_dim;
  // conditioning on the side stream (env CLPK_SIDE_STREAM=0: in line); per-launch profiling keeps everything in line
  static const bool side_ok = [] { const char* e = getenv("CLPK_SIDE_STREAM"); return !(e && atoi(e) == 0); }();
  const bool fork = side_ok && !P->prof_on && P->side_stream != nullptr;
  cudaStream_t sc = fork ? P->side_stream : s;
  if (fork) {
    CLPK_CHECK_CUDA(cudaEventRecord(P->ev_fork, s));
    CLPK_CHECK_CUDA(cudaStreamWaitEvent(sc, P->ev_fork, 0));
  }
  CLPK_TIMED(P, kProfCond, sc, launch_cond_combine(P->zemb, P->ht_tab, P->run_dev, P->hcond, P->B, td, sc));
  CLPK_TRY(film_from_h(P, sc));
  if (fork) CLPK_CHECK_CUDA(cudaEventRecord(P->ev_join, sc));
  // the stem has consumed x_buf (im2col) long before the head runs, so the head may update it in place
  const bool fused = P->head_fused && P->head_ddim;
  CLPK_TRY(forward_body(P, P->x_buf, s, fork ? P->ev_join : nullptr, fused ? P->x_buf : nullptr));
  if (!fused) {
    const long long n = (long long)P->B * P->cfg.img_ch * P->H * P->W;
    CLPK_TIMED(P, kProfDdim, s, launch_ddim_step(P->x_buf, P->eps_buf, P->coef_tab, P->run_dev, P->x_buf, n, s));
  }
  return CLPK_OK;
}

extern "C" int clpk_plan_prepare_ddim(clpk_plan* P, int steps, const int64_t* ts_host, const float* coef_host,
                                      int use_graph, void* stream) {
  CLPK_REQUIRE(P && steps > 0 && ts_host && coef_host, "clpk_plan_prepare_ddim: bad arguments");
  DevGuard dg(P->device);
  cudaStream_t s = (cudaStream_t)stream;
  const int td = P->cfg.time_dim;
  if (P->graph_exec) { cudaGraphExecDestroy(P->graph_exec); P->graph_exec = nullptr; }
  if (P->graph) { cudaGraphDestroy(P->graph); P->graph = nullptr; }
  P->steps = 0;  // not prepared until everything below (tables, capture, instantiate) has succeeded
  // per-step tables (the time half of the conditioning is batch invariant: ddim.py:32 uses one t for the whole batch).
  // ht_tab / coef_tab live in dedicated members and are re-allocated only when a run needs more steps; the
  // temporaries are freed before returning.
  struct Tmp {
    void* p[3] = {nullptr, nullptr, nullptr};
    ~Tmp() { for (void* q : p) if (q) cudaFree(q); }
  } tmp;
  CLPK_CHECK_CUDA(cudaMalloc(&tmp.p[0], (size_t)steps * sizeof(int64_t)));
  CLPK_CHECK_CUDA(cudaMalloc(&tmp.p[1], (size_t)steps * td * sizeof(float)));
  CLPK_CHECK_CUDA(cudaMalloc(&tmp.p[2], (size_t)steps * 4 * td * sizeof(float)));
  int64_t* ts_dev = reinterpret_cast<int64_t*>(tmp.p[0]);
  float* temb = reinterpret_cast<float*>(tmp.p[1]);
  float* h1 = reinterpret_cast<float*>(tmp.p[2]);
  if (steps > P->tab_cap) {
    if (P->ht_tab) { cudaFree(P->ht_tab); P->ht_tab = nullptr; P->bytes -= (long long)P->tab_cap * td * 4; }
    if (P->coef_tab) { cudaFree(P->coef_tab); P->coef_tab = nullptr; P->bytes -= (long long)P->tab_cap * 5 * 4; }
    P->tab_cap = 0;
    CLPK_CHECK_CUDA(cudaMalloc(reinterpret_cast<void**>(&P->ht_tab), (size_t)steps * td * sizeof(float)));
    CLPK_CHECK_CUDA(cudaMalloc(reinterpret_cast<void**>(&P->coef_tab), (size_t)steps * 5 * sizeof(float)));
    P->tab_cap = steps;
    P->bytes += (long long)steps * (td + 5) * 4;
  }
  CLPK_CHECK_CUDA(cudaMemcpyAsync(ts_dev, ts_host, (size_t)steps * sizeof(int64_t), cudaMemcpyHostToDevice, s));
  CLPK_CHECK_CUDA(cudaMemcpyAsync(P->coef_tab, coef_host, (size_t)steps * 5 * sizeof(float), cudaMemcpyHostToDevice, s));
  CLPK_TRY(launch_timestep_embedding(ts_dev, temb, steps, td, 10000.f, s, P->time_freqs));
  CLPK_TRY(launch_linear(temb, P->tp0_w, P->tp0_b, nullptr, 0, h1, steps, 4 * td, td, 1, s));
  CLPK_TRY(launch_linear(h1, P->tp2_w, P->tp2_b, nullptr, 0, P->ht_tab, steps, td, 4 * td, 0, s));
  CLPK_CHECK_CUDA(cudaStreamSynchronize(s));
  P->any_sigma = false;
  for (int i = 0; i < steps; ++i) P->any_sigma = P->any_sigma || (coef_host[i * 5 + 4] > 0.f);
  if (use_graph) {
    if (!P->cap_stream) CLPK_CHECK_CUDA(cudaStreamCreateWithFlags(&P->cap_stream, cudaStreamNonBlocking));
    cudaStream_t cs = P->cap_stream;
    CLPK_CHECK_CUDA(cudaStreamBeginCapture(cs, cudaStreamCaptureModeThreadLocal));
    int rc = ddim_step_body(P, cs);
    if (rc == CLPK_OK) rc = launch_ddim_advance(P->run_dev, cs);
    cudaGraph_t g = nullptr;
    cudaError_t e = cudaStreamEndCapture(cs, &g);
    if (rc != CLPK_OK) { if (g) cudaGraphDestroy(g); return rc; }
    CLPK_CHECK_CUDA(e);
    P->graph = g;
    CLPK_CHECK_CUDA(cudaGraphInstantiate(&P->graph_exec, P->graph, 0));
  }
  P->steps = steps;
  return CLPK_OK;
}

extern "C" int clpk_ddim_sample(clpk_plan* P, const float* z, float* x, const float* noise, uint64_t seed,
                                float* eps_trace, float* x_trace, void* stream) {
  CLPK_REQUIRE(P && z && x, "clpk_ddim_sample: null argument");
  if (P->steps <= 0) {
    set_error("clpk_ddim_sample: call clpk_plan_prepare_ddim first");
    return CLPK_ERR_STATE;
  }
  DevGuard dg(P->device);
  cudaStream_t s = (cudaStream_t)stream;
  const clpk_unet_config& c = P->cfg;
  const long long n = (long long)P->B * c.img_ch * P->H * P->W;
  const size_t nb = (size_t)n * sizeof(float);
  DdimRun run{};
  run.step = 0;
  run.noise = noise;
  run.noise_step_stride = n;
  run.seed = seed;
  CLPK_CHECK_CUDA(cudaMemcpyAsync(P->run_dev, &run, sizeof(run), cudaMemcpyHostToDevice, s));
  CLPK_CHECK_CUDA(cudaMemcpyAsync(P->x_buf, x, nb, cudaMemcpyDeviceToDevice, s));
  // step-invariant half of the conditioning: zemb = SiLU(z_proj(z))  (unet.py:85)
  CLPK_TRY(launch_linear(z, P->zp_w, P->zp_b, nullptr, 0, P->zemb, P->B, c.time_dim, c.z_dim, 1, s));
  for (int i = 0; i < P->steps; ++i) {
    if (x_trace) CLPK_CHECK_CUDA(cudaMemcpyAsync(x_trace + (size_t)i * n, P->x_buf, nb, cudaMemcpyDeviceToDevice, s));
    if (P->graph_exec) {
      CLPK_CHECK_CUDA(cudaGraphLaunch(P->graph_exec, s));
      count_launch(P->launches_fwd - 5 + 2 + ((P->head_fused && P->head_ddim) ? 1 : 2));  // cond (2) + forward + update / advance
    } else {
      CLPK_TRY(ddim_step_body(P, s));
      CLPK_TRY(launch_ddim_advance(P->run_dev, s));
    }
    if (eps_trace)
      CLPK_CHECK_CUDA(cudaMemcpyAsync(eps_trace + (size_t)i * n, P->eps_buf, nb, cudaMemcpyDeviceToDevice, s));
  }
  CLPK_CHECK_CUDA(cudaMemcpyAsync(x, P->x_buf, nb, cudaMemcpyDeviceToDevice, s));
  // `run` lives on this frame and was copied with cudaMemcpyAsync from pageable memory (staged synchronously by the
  // runtime), so no extra synchronisation is needed here.
  return CLPK_OK;
}

// Runs `iters` eager (non-graph) DDIM steps of the prepared run on the plan's own buffers with CUDA events around every
// launch and returns, per kernel class, the summed device time (ms) and the launch-pair count:
//   0 ResBlock 3x3 convs (tcgen05)  1 other tcgen05 convs (stride 2, transposed, out)  2 GroupNorm (stats+apply)
//   3 stem conv  4 conditioning (combine + FiLM GEMV)  5 DDIM update
// The state advances like a real run (step counter wraps), so the numbers are those of the production kernels on
// production-shaped data.  Used by bench.py for the live roofline figures.
extern "C" int clpk_plan_profile_steps(clpk_plan* P, int iters, float* ms_out6, int* count_out6, void* stream) {
  CLPK_REQUIRE(P && ms_out6 && count_out6 && iters > 0, "clpk_plan_profile_steps: bad arguments");
  if (P->steps <= 0) {
    set_error("clpk_plan_profile_steps: call clpk_plan_prepare_ddim first");
    return CLPK_ERR_STATE;
  }
  DevGuard dg(P->device);
  cudaStream_t s = (cudaStream_t)stream;
  for (int k = 0; k < 6; ++k) { ms_out6[k] = 0.f; count_out6[k] = 0; }
  DdimRun run{};
  for (int it = 0; it < iters; ++it) {
    run.step = it % P->steps;
    run.noise = nullptr;
    run.seed = 0;
    run.noise_step_stride = 0;
    CLPK_CHECK_CUDA(cudaMemcpyAsync(P->run_dev, &run, sizeof(run), cudaMemcpyHostToDevice, s));
    // hold the stream while the host enqueues the step, so that no event pair contains host launch latency
    CLPK_TRY(launch_delay(3000000, s));
    P->prof_on = true;
    P->prof_used = 0;
    P->prof_cat.clear();
    int rc = ddim_step_body(P, s);
    // calibration: 16 EMPTY event pairs in the same stream give the cost of the bracketing itself (event processing
    // between two timestamps), which is subtracted from every measured pair
    constexpr int kCal = 16;
    for (int k = 0; k < kCal; ++k) { P->prof_mark(6, s); P->prof_mark(-1, s); }
    P->prof_on = false;
    if (rc != CLPK_OK) return rc;
    CLPK_CHECK_CUDA(cudaStreamSynchronize(s));
    float overhead = 0.f;
    for (size_t i = 0; i + 1 < P->prof_used; i += 2) {
      if (P->prof_cat[i] != 6) continue;
      float ms = 0.f;
      CLPK_CHECK_CUDA(cudaEventElapsedTime(&ms, P->prof_ev[i], P->prof_ev[i + 1]));
      overhead += ms / kCal;
    }
    for (size_t i = 0; i + 1 < P->prof_used; i += 2) {
      float ms = 0.f;
      CLPK_CHECK_CUDA(cudaEventElapsedTime(&ms, P->prof_ev[i], P->prof_ev[i + 1]));
      const int cat = P->prof_cat[i];
      if (cat >= 0 && cat < 6) { ms_out6[cat] += std::max(ms - overhead, 0.f); count_out6[cat] += 1; }
    }
  }
  return CLPK_OK;
}

// algorithmic FLOPs per forward of the two tcgen05 conv classes (0: ResBlock convs, 1: other) and the number of fp32
// elements GroupNorm reads per forward (all 2*n_resblocks + 1 instances), at the plan's batch.
extern "C" int clpk_plan_work_breakdown(const clpk_plan* P, double* conv_res_flops, double* conv_other_flops,
                                        double* gn_elements) {
  CLPK_REQUIRE(P && conv_res_flops && conv_other_flops && gn_elements, "clpk_plan_work_breakdown: null argument");
  double a = 0, b = 0, g = 0;
  for (const ResBlockPlan& rb : P->rbs) {
    a += rb.conv1.flops + rb.conv2.flops;
    g += ((rb.gn1.in_consumer ? 0.0 : 1.0) + (rb.gn2.in_consumer ? 0.0 : 1.0)) * P->B * (double)rb.h * rb.w * rb.c;
  }
  for (const ConvPlan& c : P->downs) b += c.flops;
  for (const ConvPlan& c : P->ups) b += c.flops;
  b += P->out_conv.flops;
  if (!P->out_gn.in_consumer) g += (double)P->B * P->H * P->W * P->cfg.base;  // out_norm
  *conv_res_flops = a; *conv_other_flops = b; *gn_elements = g;
  return CLPK_OK;
}

// algorithmic HBM bytes of all GroupNorm(+SiLU) applies of one forward at the plan's batch: every instance reads its
// input once (2 B / element from a 16-bit tensor, 4 B from fp32) and writes the 16-bit operand once (2 B)
extern "C" int clpk_plan_groupnorm_bytes(const clpk_plan* P, double* bytes) {
  CLPK_REQUIRE(P && bytes, "clpk_plan_groupnorm_bytes: null argument");
  double tot = 0;
  auto in_bytes_x = [&](const GnPlan& gn) { return (gn.fused && P->x16(gn.level)) ? 2.0 : 4.0; };
  for (const ResBlockPlan& rb : P->rbs) {  // (GroupNorms applied inside their consumer conv have no stand-alone pass)
    const double n = (double)P->B * rb.h * rb.w * rb.c;
    if (!rb.gn1.in_consumer) tot += n * (in_bytes_x(rb.gn1) + 2.0);
    if (!rb.gn2.in_consumer) tot += n * ((rb.y16 ? 2.0 : 4.0) + 2.0);
  }
  if (!P->out_gn.in_consumer) tot += (double)P->B * P->H * P->W * P->cfg.base * ((P->head16 ? 2.0 : in_bytes_x(P->out_gn)) + 2.0);
  *bytes = tot;
  return CLPK_OK;
}
