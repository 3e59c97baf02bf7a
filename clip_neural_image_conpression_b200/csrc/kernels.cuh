// kernels.cuh — internal launchers shared between translation units of libclpk.so.
#pragma once
#include "common.cuh"

#include <algorithm>

namespace clpk {

// Device-resident state of one DDIM run (read by the captured step graph, advanced by ddim_advance).
struct DdimRun {
  int step;
  int pad_;
  const float* noise;            // [steps][n] pre-drawn N(0,1) or nullptr (-> in-kernel Philox)
  long long noise_step_stride;   // n
  unsigned long long seed;
};

// elementwise.cu
int launch_ddim_step(const float* x, const float* eps, const float* coef_tab_dev, const DdimRun* run_dev, float* x_out,
                     long long n, cudaStream_t stream);
int launch_ddim_advance(DdimRun* run_dev, cudaStream_t stream);
int launch_cond_combine(const float* zemb, const float* ht_tab, const DdimRun* run, float* h, int batch, int dim,
                        cudaStream_t stream);
int launch_add_const(float* p, float v, int n, cudaStream_t stream);
int launch_delay(long long ns, cudaStream_t stream);
int launch_timestep_embedding(const int64_t* t, float* out, int batch, int dim, float max_period, cudaStream_t stream,
                              const float* freqs = nullptr);
int launch_linear(const float* x, const float* w, const float* b, const float* add, int add_rows, float* y, int m, int n,
                  int k, int act, cudaStream_t stream);

// groupnorm.cu
struct GnShape {
  int batch, hw, c, groups;
  int chunks;  // partial-sum blocks per image
};
GnShape gn_shape(int batch, int hw, int c, int groups);
long long gn_ws_bytes(const GnShape& s);
int launch_groupnorm(const float* x, const float* gamma, const float* beta, void* y_op, void* ws, const GnShape& s,
                     float eps, int silu, int op_dtype, cudaStream_t stream);
int launch_gn_stats(const float* x, void* ws, const GnShape& s, float eps, const float2** stats_out, cudaStream_t stream);
int launch_gn_finalize(const float2* partial, float2* stats, int batch, int slots, int groups, float eps,
                       cudaStream_t stream);
int launch_gn_affine(const float2* partial, const float* gamma, const float* beta, float* scale, float* shift, int batch,
                     int slots, int pieces, int groups, int c, float eps, cudaStream_t stream);
int launch_gn_apply_ex(const void* x, int x_is_16, const float* gamma, const float* beta, const float2* stats,
                       const float2* partial, int slots, int pieces, double n_per_group, float eps, void* y_op,
                       const GnShape& s, int silu, int op_dtype, cudaStream_t stream);
int launch_gn_apply(const float* x, const float* gamma, const float* beta, const float2* stats, void* y_op,
                    const GnShape& s, int silu, int op_dtype, cudaStream_t stream);

// head_conv.cu
bool head_conv_supported(int h, int w, int c, int cout);
int launch_head_conv(const void* x_op, const float* scale, const float* shift, const void* w_packed, const float* bias,
                     float* out_nchw, int batch, int h, int w, int c, int op_dtype, cudaStream_t stream, int max_stages = 0,
                     float* ddim_x = nullptr, const float* ddim_coef = nullptr, const DdimRun* ddim_run = nullptr);
int launch_pack_head_weight(const float* w, void* out_op, int c, int op_dtype, cudaStream_t stream);

// conv_in.cu
int launch_conv_in(const float* x_nchw, const float* w, const float* b, float* y_nhwc, int batch, int cin, int h, int w_,
                   int cout, cudaStream_t stream);

int launch_stem_im2col(const float* x, void* cols, int batch, int cin, int h, int wd, int op_dtype, cudaStream_t stream);
int launch_pad_rows(const float* src, float* dst, int rows, int k_src, int k_dst, cudaStream_t stream);

}  // namespace clpk
