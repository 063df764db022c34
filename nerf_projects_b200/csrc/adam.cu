// Fused multi-tensor Adam step: torch.optim.Adam(betas=(0.9,0.999)) of
// create_nerf (ipynb:905) applied by train() (ipynb:1275).  All parameter
// tensors of both networks are updated by ONE launch (pointer table passed by
// value), matching torch's single-tensor formula:
//   m = b1 m + (1-b1) g;  v = b2 v + (1-b2) g^2
//   p -= (lr / (1-b1^t)) * m / (sqrt(v)/sqrt(1-b2^t) + eps)
#include "common.cuh"
#include <cmath>

namespace nerfb {

struct AdamTable {
  float* p[NERFB_ADAM_MAX];
  const float* g[NERFB_ADAM_MAX];
  float* m[NERFB_ADAM_MAX];
  float* v[NERFB_ADAM_MAX];
  long long start[NERFB_ADAM_MAX + 1];  // prefix of numel in units of blocks
  long long numel[NERFB_ADAM_MAX];
  int n;
};

constexpr int ADAM_BLOCK = 256, ADAM_PER_THREAD = 4;

__global__ void __launch_bounds__(ADAM_BLOCK)
adam_kernel(const __grid_constant__ AdamTable t, float b1, float b2, float step_size, float inv_bc2_sqrt,
            float eps, float grad_scale) {
  // find tensor of this block (n <= 64: linear scan in registers/const bank)
  const long long blk = blockIdx.x;
  int ti = 0;
  while (ti + 1 < t.n && blk >= t.start[ti + 1]) ++ti;
  const long long base = (blk - t.start[ti]) * (ADAM_BLOCK * ADAM_PER_THREAD);
  float* __restrict__ p = t.p[ti];
  const float* __restrict__ g = t.g[ti];
  float* __restrict__ m = t.m[ti];
  float* __restrict__ v = t.v[ti];
  const long long n = t.numel[ti];
#pragma unroll
  for (int q = 0; q < ADAM_PER_THREAD; ++q) {
    const long long i = base + q * ADAM_BLOCK + threadIdx.x;
    if (i < n) {
      const float gi = g[i] * grad_scale;
      const float mi = b1 * m[i] + (1.f - b1) * gi;
      const float vi = b2 * v[i] + (1.f - b2) * gi * gi;
      m[i] = mi;
      v[i] = vi;
      p[i] -= step_size * mi / (sqrtf(vi) * inv_bc2_sqrt + eps);
    }
  }
}

}  // namespace nerfb

using namespace nerfb;

extern "C" int nerfb_adam_step(int n_tensors, float* const* h_p, const float* const* h_g, float* const* h_m,
                               float* const* h_v, const long long* h_numel, float lr, float beta1, float beta2,
                               float eps, int step, float grad_scale, void* stream) {
  NERFB_REQUIRE(n_tensors >= 0 && n_tensors <= NERFB_ADAM_MAX && step >= 1,
                "nerfb_adam_step: n_tensors must be <= %d and step >= 1", NERFB_ADAM_MAX);
  if (n_tensors == 0) return NERFB_OK;
  AdamTable t{};
  t.n = n_tensors;
  long long blocks = 0;
  for (int i = 0; i < n_tensors; ++i) {
    NERFB_REQUIRE(h_p[i] && h_g[i] && h_m[i] && h_v[i] && h_numel[i] >= 0, "nerfb_adam_step: null tensor %d", i);
    t.p[i] = h_p[i]; t.g[i] = h_g[i]; t.m[i] = h_m[i]; t.v[i] = h_v[i];
    t.numel[i] = h_numel[i];
    t.start[i] = blocks;
    blocks += (h_numel[i] + ADAM_BLOCK * ADAM_PER_THREAD - 1) / (ADAM_BLOCK * ADAM_PER_THREAD);
  }
  t.start[n_tensors] = blocks;
  if (blocks == 0) return NERFB_OK;
  const double bc1 = 1.0 - std::pow((double)beta1, step);
  const double bc2 = 1.0 - std::pow((double)beta2, step);
  const float step_size = (float)(lr / bc1);
  const float inv_bc2_sqrt = (float)(1.0 / std::sqrt(bc2));
  adam_kernel<<<(unsigned)blocks, ADAM_BLOCK, 0, as_stream(stream)>>>(t, beta1, beta2, step_size, inv_bc2_sqrt, eps,
                                                                     grad_scale);
  return check_launch("adam_kernel");
}
