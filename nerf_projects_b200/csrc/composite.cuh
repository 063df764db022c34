// Device helpers of the alpha-compositing kernels (composite.cu) shared with the fused composite + resample kernel
// (sample_pdf.cu): ray norm, sigmoid, warp sum and the per-ray transmittance scan of raw2outputs (ipynb:254-349).
#pragma once
#include "common.cuh"

namespace nerfb {

__device__ __forceinline__ float ray_norm(const float* __restrict__ d) {
  const float x = __ldg(d), y = __ldg(d + 1), z = __ldg(d + 2);
  return __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y)), __fmul_rn(z, z)));
}

__device__ __forceinline__ float sigmoidf_(float x) { return __fdiv_rn(1.0f, __fadd_rn(1.0f, expf(-x))); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Computes alpha_i, x_i = 1-alpha_i+1e-10 and the exclusive product T_i for one
// ray; results for sample i are handed to `emit(i, alpha, x, T)` in order of
// 32-sample chunks.
template <typename Emit>
__device__ __forceinline__ void scan_ray(int S, const float* __restrict__ raw, int ch,
                                         const float* __restrict__ z, float nrm,
                                         const float* __restrict__ noise, int lane, Emit emit) {
  double carry = 1.0;
  for (int base = 0; base < S; base += 32) {
    const int i = base + lane;
    float alpha = 0.f, x = 1.f;
    if (i < S) {
      float dist = (i + 1 < S) ? __fsub_rn(z[i + 1], z[i]) : 1e10f;
      dist = __fmul_rn(dist, nrm);
      float s = raw[(size_t)i * ch + 3];
      if (noise) s = __fadd_rn(s, noise[i]);
      const float act = fmaxf(s, 0.f);
      alpha = __fsub_rn(1.0f, expf(__fmul_rn(-act, dist)));
      x = __fadd_rn(__fsub_rn(1.0f, alpha), 1e-10f);
    }
    // inclusive product scan over the chunk (fp64)
    double p = (double)x;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const double q = __shfl_up_sync(0xffffffffu, p, o);
      if (lane >= o) p *= q;
    }
    double excl = __shfl_up_sync(0xffffffffu, p, 1);
    if (lane == 0) excl = 1.0;
    // torch.cumprod keeps a double running product and rounds each output
    const float T = (float)(carry * excl);
    if (i < S) emit(i, alpha, x, T);
    carry *= __shfl_sync(0xffffffffu, p, 31);
  }
}

}  // namespace nerfb
