// conv_in.cu — the UNet stem: Conv2d(img_ch -> base, 3x3, pad 1) (PKG/models/unet.py:55,88).
// K = 9*img_ch = 27 is far too small for the tensor cores (1.8 GFMA at batch 8 vs 268 MB of output), so it runs in fp32
// on the CUDA cores and doubles as the NCHW -> NHWC layout change:
// reads the reference-layout x_t [B,cin,H,W], writes the NHWC fp32 residual stream [B,H,W,cout].
#include "kernels.cuh"

namespace clpk {

constexpr int kCinMax = 4;
constexpr int kPx = 4;  // consecutive pixels (along W) per thread

// Each thread owns 4 output channels (one float4 store per pixel; a warp writes 512 contiguous bytes when cout == 128)
// and walks groups of kPx pixels.  Its CIN*9*4 weights live in registers for the whole kernel; the CIN*3*(kPx+2) inputs
// of a pixel group are warp-broadcast loads.  FMA-bound on the fp32 pipe at ~the HBM write time of the output.
template <int CIN>
__global__ void __launch_bounds__(128)
conv_in_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
               float* __restrict__ y, int batch, int h, int wd, int cout) {
  const int qpp = cout >> 2;                         // channel quads per pixel
  const int gpb = blockDim.x / qpp;                  // pixel groups per block iteration
  const int q = threadIdx.x % qpp, gl = threadIdx.x / qpp;
  if (gl >= gpb) return;
  float4 wr[CIN * 9];                                // wr[k] = weights of tap k for channels 4q..4q+3
#pragma unroll
  for (int k = 0; k < CIN * 9; ++k)
    wr[k] = make_float4(w[(4 * q + 0) * CIN * 9 + k], w[(4 * q + 1) * CIN * 9 + k], w[(4 * q + 2) * CIN * 9 + k],
                        w[(4 * q + 3) * CIN * 9 + k]);
  const float4 b4 = *reinterpret_cast<const float4*>(bias + 4 * q);
  const int gw = (wd + kPx - 1) / kPx;               // pixel groups per image row
  const long long ngroups = (long long)batch * h * gw;
  const long long plane = (long long)h * wd;
  for (long long grp = (long long)blockIdx.x * gpb + gl; grp < ngroups; grp += (long long)gridDim.x * gpb) {
    const int gx = (int)(grp % gw);
    const long long r = grp / gw;
    const int yy = (int)(r % h);
    const int b = (int)(r / h);
    const int x0 = gx * kPx;
    float4 acc[kPx];
#pragma unroll
    for (int i = 0; i < kPx; ++i) acc[i] = b4;
#pragma unroll
    for (int c = 0; c < CIN; ++c) {
      const float* xp = x + ((long long)b * CIN + c) * plane;
#pragma unroll
      for (int rr = 0; rr < 3; ++rr) {
        const int iy = yy + rr - 1;
        float in[kPx + 2];
#pragma unroll
        for (int j = 0; j < kPx + 2; ++j) {
          const int ix = x0 + j - 1;
          in[j] = (iy >= 0 && iy < h && ix >= 0 && ix < wd) ? __ldg(xp + (long long)iy * wd + ix) : 0.f;
        }
#pragma unroll
        for (int ss = 0; ss < 3; ++ss) {
          const float4 wv = wr[(c * 3 + rr) * 3 + ss];
#pragma unroll
          for (int i = 0; i < kPx; ++i) {
            const float v = in[i + ss];
            acc[i].x = fmaf(v, wv.x, acc[i].x); acc[i].y = fmaf(v, wv.y, acc[i].y);
            acc[i].z = fmaf(v, wv.z, acc[i].z); acc[i].w = fmaf(v, wv.w, acc[i].w);
          }
        }
      }
    }
    float4* yp = reinterpret_cast<float4*>(y + (((long long)b * h + yy) * wd + x0) * cout) + q;
#pragma unroll
    for (int i = 0; i < kPx; ++i)
      if (x0 + i < wd) __stcs(yp + (long long)i * qpp, acc[i]);
  }
}

int launch_conv_in(const float* x, const float* w, const float* b, float* y, int batch, int cin, int h, int wd, int cout,
                   cudaStream_t stream) {
  CLPK_REQUIRE(cin >= 1 && cin <= kCinMax, "conv_in supports 1..%d input channels (got %d)", kCinMax, cin);
  CLPK_REQUIRE(cout % 4 == 0 && cout / 4 <= 128, "conv_in needs Cout %% 4 == 0 and Cout <= 512");
  const int gpb = 128 / (cout / 4);
  const long long ngroups = (long long)batch * h * ((wd + kPx - 1) / kPx);
  const int blocks = (int)std::min<long long>((ngroups + gpb - 1) / gpb, (long long)num_sms() * 12);
  switch (cin) {
    case 1: conv_in_kernel<1><<<blocks, 128, 0, stream>>>(x, w, b, y, batch, h, wd, cout); break;
    case 2: conv_in_kernel<2><<<blocks, 128, 0, stream>>>(x, w, b, y, batch, h, wd, cout); break;
    case 3: conv_in_kernel<3><<<blocks, 128, 0, stream>>>(x, w, b, y, batch, h, wd, cout); break;
    case 4: conv_in_kernel<4><<<blocks, 128, 0, stream>>>(x, w, b, y, batch, h, wd, cout); break;
  }
  CLPK_CHECK_LAUNCH();
  return CLPK_OK;
}

// ------------------------------------------------------------------------------------------------ stem im2col
// One thread per pixel: 27 (cin*9) neighbourhood values -> 32 16-bit columns = 64 contiguous bytes (4 x 16-byte stores).
// Reads are coalesced along W within each (channel, row).
__global__ void __launch_bounds__(256)
stem_im2col_kernel(const float* __restrict__ x, uint16_t* __restrict__ cols, int batch, int cin, int h, int wd, int f16) {
  pdl_prologue_done();
  const long long npix = (long long)batch * h * wd;
  const long long plane = (long long)h * wd;
  for (long long pix = (long long)blockIdx.x * blockDim.x + threadIdx.x; pix < npix; pix += (long long)gridDim.x * blockDim.x) {
    const int b = (int)(pix / plane);
    const int rem = (int)(pix - (long long)b * plane);
    const int yy = rem / wd, xx = rem - yy * wd;
    float v[32];
#pragma unroll
    for (int k = 0; k < 32; ++k) v[k] = 0.f;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      if (c < cin) {
        const float* xp = x + ((long long)b * cin + c) * plane;
#pragma unroll
        for (int r = 0; r < 3; ++r) {
          const int iy = yy + r - 1;
#pragma unroll
          for (int s = 0; s < 3; ++s) {
            const int ix = xx + s - 1;
            v[c * 9 + r * 3 + s] = (iy >= 0 && iy < h && ix >= 0 && ix < wd) ? __ldg(xp + (long long)iy * wd + ix) : 0.f;
          }
        }
      }
    }
    uint4* dst = reinterpret_cast<uint4*>(cols + pix * 32);
#pragma unroll
    for (int j = 0; j < 4; ++j)
      dst[j] = make_uint4(pack_op2(v[8 * j + 0], v[8 * j + 1], f16 != 0), pack_op2(v[8 * j + 2], v[8 * j + 3], f16 != 0),
                          pack_op2(v[8 * j + 4], v[8 * j + 5], f16 != 0), pack_op2(v[8 * j + 6], v[8 * j + 7], f16 != 0));
  }
}

int launch_stem_im2col(const float* x, void* cols, int batch, int cin, int h, int wd, int op_dtype, cudaStream_t stream) {
  CLPK_REQUIRE(cin >= 1 && cin * 9 <= 27, "stem im2col supports cin <= 3 (got %d)", cin);
  const long long npix = (long long)batch * h * wd;
  const int blocks = (int)std::min<long long>((npix + 255) / 256, (long long)num_sms() * 16);
  CLPK_CHECK_CUDA(launch_kernel_pdl(stem_im2col_kernel, dim3(blocks), dim3(256), 0, stream, x, reinterpret_cast<uint16_t*>(cols),
                                    batch, cin, h, wd, (int)(op_dtype == CLPK_OP_F16)));
  CLPK_CHECK_LAUNCH();
  return CLPK_OK;
}

// [rows][k_src] fp32 -> [rows][k_dst] fp32, zero padded (stem weight [cout][cin*9] -> [cout][32])
__global__ void pad_rows_kernel(const float* __restrict__ src, float* __restrict__ dst, int rows, int k_src, int k_dst) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * k_dst) return;
  const int r = i / k_dst, k = i - r * k_dst;
  dst[i] = (k < k_src) ? src[r * k_src + k] : 0.f;
}
int launch_pad_rows(const float* src, float* dst, int rows, int k_src, int k_dst, cudaStream_t stream) {
  pad_rows_kernel<<<(rows * k_dst + 255) / 256, 256, 0, stream>>>(src, dst, rows, k_src, k_dst);
  CLPK_CHECK_LAUNCH();
  return CLPK_OK;
}

}  // namespace clpk

extern "C" int clpk_stem_im2col(const float* x, void* cols, int batch, int cin, int h, int wd, int op_dtype,
                                void* stream) {
  CLPK_REQUIRE(x && cols && batch > 0 && h > 0 && wd > 0, "clpk_stem_im2col: bad arguments");
  CLPK_REQUIRE(op_dtype == CLPK_OP_BF16 || op_dtype == CLPK_OP_F16, "clpk_stem_im2col: operand dtype %d unknown", op_dtype);
  return clpk::launch_stem_im2col(x, cols, batch, cin, h, wd, op_dtype, (cudaStream_t)stream);
}

extern "C" int clpk_conv_in(const float* x, const float* w, const float* b, float* y, int batch, int cin, int h, int wd,
                            int cout, void* stream) {
  CLPK_REQUIRE(x && w && b && y && batch > 0 && h > 0 && wd > 0, "clpk_conv_in: bad arguments");
  return clpk::launch_conv_in(x, w, b, y, batch, cin, h, wd, cout, (cudaStream_t)stream);
}
