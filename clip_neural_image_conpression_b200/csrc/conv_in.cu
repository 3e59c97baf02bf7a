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

}  // namespace clpk

extern "C" int clpk_conv_in(const float* x, const float* w, const float* b, float* y, int batch, int cin, int h, int wd,
                            int cout, void* stream) {
  CLPK_REQUIRE(x && w && b && y && batch > 0 && h > 0 && wd > 0, "clpk_conv_in: bad arguments");
  return clpk::launch_conv_in(x, w, b, y, batch, cin, h, wd, cout, (cudaStream_t)stream);
}
