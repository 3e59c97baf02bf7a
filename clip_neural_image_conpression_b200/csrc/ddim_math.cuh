// ddim_math.cuh — the DDIM update of one element (PKG/diffusion/ddim.py:36-45) and the in-kernel noise source, shared by
// ddim_step_kernel (elementwise.cu) and the fused head kernel (head_conv.cu), so both evaluate the SAME instruction sequence.
#pragma once
#include "kernels.cuh"

namespace clpk {

// ddim.py:36-45.  Every arithmetic op is individually rounded (no FMA contraction) to stay bit-identical to ATen.
struct DdimCoef { float c_eps, c_den, c_s, c_dir, sigma; };

__device__ __forceinline__ float ddim_update(float x, float e, const DdimCoef& k) {
  float x0 = __fdiv_rn(__fsub_rn(x, __fmul_rn(k.c_eps, e)), k.c_den);
  // torch.clamp propagates NaN; fminf/fmaxf would drop it
  x0 = (x0 != x0) ? x0 : fminf(fmaxf(x0, -1.f), 1.f);
  return __fadd_rn(__fmul_rn(k.c_s, x0), __fmul_rn(k.c_dir, e));
}

// Philox4x32-10 (Salmon et al.), counter = (element quad, step), key = seed; Box-Muller to N(0,1).
__device__ __forceinline__ void philox4x32_10(uint32_t (&ctr)[4], uint32_t k0, uint32_t k1) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, ctr[0]), lo0 = 0xD2511F53u * ctr[0];
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, ctr[2]), lo1 = 0xCD9E8D57u * ctr[2];
    const uint32_t n0 = hi1 ^ ctr[1] ^ k0, n1 = lo1, n2 = hi0 ^ ctr[3] ^ k1, n3 = lo0;
    ctr[0] = n0; ctr[1] = n1; ctr[2] = n2; ctr[3] = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
}
__device__ __forceinline__ void box_muller(uint32_t a, uint32_t b, float& n0, float& n1) {
  const float u = ((float)a + 1.0f) * 2.3283064365386963e-10f;  // (0,1]
  const float v = (float)b * 2.3283064365386963e-10f;
  const float r = sqrtf(-2.0f * __logf(u));
  float s, c;
  __sincosf(6.283185307179586f * v, &s, &c);
  n0 = r * c; n1 = r * s;
}


// coefficients of the run's current step from the device-resident table [steps][5]
__device__ __forceinline__ DdimCoef ddim_load_coef(const float* __restrict__ coef_tab, int step) {
  DdimCoef k;
  k.c_eps = coef_tab[step * 5 + 0]; k.c_den = coef_tab[step * 5 + 1]; k.c_s = coef_tab[step * 5 + 2];
  k.c_dir = coef_tab[step * 5 + 3]; k.sigma = coef_tab[step * 5 + 4];
  return k;
}
// the update of four consecutive elements i4*4 .. i4*4+3 (i4 = the float4 index the Philox counter is keyed with)
__device__ __forceinline__ float4 ddim_update4(float4 xv, float4 ev, const DdimCoef& k, bool stochastic, const float* nz,
                                               long long i4, int step, unsigned long long seed) {
  float4 o;
  o.x = ddim_update(xv.x, ev.x, k); o.y = ddim_update(xv.y, ev.y, k);
  o.z = ddim_update(xv.z, ev.z, k); o.w = ddim_update(xv.w, ev.w, k);
  if (stochastic) {
    float4 z;
    if (nz) {
      z = reinterpret_cast<const float4*>(nz)[i4];
    } else {
      uint32_t ctr[4] = {(uint32_t)i4, (uint32_t)(i4 >> 32), (uint32_t)step, 0x636c706bu};
      philox4x32_10(ctr, (uint32_t)seed, (uint32_t)(seed >> 32));
      box_muller(ctr[0], ctr[1], z.x, z.y);
      box_muller(ctr[2], ctr[3], z.z, z.w);
    }
    o.x = __fadd_rn(o.x, __fmul_rn(k.sigma, z.x)); o.y = __fadd_rn(o.y, __fmul_rn(k.sigma, z.y));
    o.z = __fadd_rn(o.z, __fmul_rn(k.sigma, z.z)); o.w = __fadd_rn(o.w, __fmul_rn(k.sigma, z.w));
  }
  return o;
}

}  // namespace clpk
