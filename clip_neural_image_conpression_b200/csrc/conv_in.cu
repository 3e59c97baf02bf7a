// conv_in.cu — the UNet stem: Conv2d(img_ch -> base, 3x3, pad 1) (PKG/models/unet.py:55,88).
// K = 9*img_ch = 27 is far too small for the tensor cores and the layer is purely write-bound (base*4 bytes per pixel
// out vs 12 bytes in), so it runs in fp32 on the CUDA cores and doubles as the NCHW -> NHWC layout change:
// reads the reference-layout x_t [B,cin,H,W], writes the NHWC fp32 residual stream [B,H,W,cout].
#include "kernels.cuh"

namespace clpk {

constexpr int kCinMax = 4;

// block = 256 threads = (256 / (cout/4)) pixels x (cout/4) channel quads; each thread produces 4 output channels of
// one pixel (one coalesced float4 store; a warp writes 512 contiguous bytes when cout == 128).
template <int CIN>
__global__ void __launch_bounds__(256)
conv_in_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
               float* __restrict__ y, int batch, int h, int wd, int cout) {
  extern __shared__ float wsm[];  // [CIN*9][cout] transposed weights, then bias[cout]
  const int k = CIN * 9;
  for (int i = threadIdx.x; i < k * cout; i += blockDim.x) {
    const int n = i / k, kk = i - n * k;  // source layout [cout][cin][3][3]
    wsm[kk * cout + n] = w[i];
  }
  for (int i = threadIdx.x; i < cout; i += blockDim.x) wsm[k * cout + i] = bias[i];
  __syncthreads();
  const int qpp = cout >> 2;              // quads per pixel
  const int ppb = blockDim.x / qpp;       // pixels per block iteration
  const int q = threadIdx.x % qpp, pl = threadIdx.x / qpp;
  if (pl >= ppb) return;
  const long long npix = (long long)batch * h * wd;
  const long long plane = (long long)h * wd;
  for (long long pix = (long long)blockIdx.x * ppb + pl; pix < npix; pix += (long long)gridDim.x * ppb) {
    const int b = (int)(pix / plane);
    const int rem = (int)(pix - (long long)b * plane);
    const int yy = rem / wd, xx = rem - yy * wd;
    float4 acc = *reinterpret_cast<const float4*>(&wsm[k * cout + 4 * q]);
#pragma unroll
    for (int c = 0; c < CIN; ++c) {
      const float* xp = x + ((long long)b * CIN + c) * plane;
#pragma unroll
      for (int r = 0; r < 3; ++r) {
        const int iy = yy + r - 1;
#pragma unroll
        for (int s = 0; s < 3; ++s) {
          const int ix = xx + s - 1;
          const float v = (iy >= 0 && iy < h && ix >= 0 && ix < wd) ? __ldg(xp + (long long)iy * wd + ix) : 0.f;
          const float4 wv = *reinterpret_cast<const float4*>(&wsm[((c * 3 + r) * 3 + s) * cout + 4 * q]);
          acc.x = fmaf(v, wv.x, acc.x); acc.y = fmaf(v, wv.y, acc.y);
          acc.z = fmaf(v, wv.z, acc.z); acc.w = fmaf(v, wv.w, acc.w);
        }
      }
    }
    __stcs(reinterpret_cast<float4*>(y + pix * cout) + q, acc);
  }
}

int launch_conv_in(const float* x, const float* w, const float* b, float* y, int batch, int cin, int h, int wd, int cout,
                   cudaStream_t stream) {
  CLPK_REQUIRE(cin >= 1 && cin <= kCinMax, "conv_in supports 1..%d input channels (got %d)", kCinMax, cin);
  CLPK_REQUIRE(cout % 4 == 0 && cout / 4 <= 256, "conv_in needs Cout %% 4 == 0 and Cout <= 1024");
  const int smem = (cin * 9 + 1) * cout * (int)sizeof(float);
  const int ppb = 256 / (cout / 4);
  const long long npix = (long long)batch * h * wd;
  const int blocks = (int)std::min<long long>((npix + ppb - 1) / ppb, (long long)num_sms() * 16);
#define CLPK_CONV_IN_CASE(N)                                                                                        \
  case N:                                                                                                           \
    if (smem > 48 * 1024)                                                                                           \
      CLPK_CHECK_CUDA(cudaFuncSetAttribute(conv_in_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));  \
    conv_in_kernel<N><<<blocks, 256, smem, stream>>>(x, w, b, y, batch, h, wd, cout);                               \
    break;
  switch (cin) {
    CLPK_CONV_IN_CASE(1)
    CLPK_CONV_IN_CASE(2)
    CLPK_CONV_IN_CASE(3)
    CLPK_CONV_IN_CASE(4)
  }
#undef CLPK_CONV_IN_CASE
  CLPK_CHECK_LAUNCH();
  return CLPK_OK;
}

}  // namespace clpk

extern "C" int clpk_conv_in(const float* x, const float* w, const float* b, float* y, int batch, int cin, int h, int wd,
                            int cout, void* stream) {
  CLPK_REQUIRE(x && w && b && y && batch > 0 && h > 0 && wd > 0, "clpk_conv_in: bad arguments");
  return clpk::launch_conv_in(x, w, b, y, batch, cin, h, wd, cout, (cudaStream_t)stream);
}
