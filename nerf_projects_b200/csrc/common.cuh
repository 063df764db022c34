// Shared helpers for libnerf_b200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdarg>
#include "../../include/nerf_b200.h"

namespace nerfb {

void set_error(const char* fmt, ...);

inline int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("%s: %s", what, cudaGetErrorString(e));
    return NERFB_ERR_CUDA;
  }
  return NERFB_OK;
}

#define NERFB_REQUIRE(cond, ...)            \
  do {                                      \
    if (!(cond)) {                          \
      nerfb::set_error(__VA_ARGS__);        \
      return NERFB_ERR_ARG;                 \
    }                                       \
  } while (0)

#define NERFB_CUDA(expr)                                                        \
  do {                                                                          \
    cudaError_t _e = (expr);                                                    \
    if (_e != cudaSuccess) {                                                    \
      nerfb::set_error("%s failed: %s", #expr, cudaGetErrorString(_e));         \
      return NERFB_ERR_CUDA;                                                    \
    }                                                                           \
  } while (0)

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

inline int cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }

int sm_count();

// torch.linspace(0,1,n)[i] in fp32 (ATen CPU formula):
// step = fl(1/(n-1)); i < n/2 -> fl(step*i); else single-rounded fma(-step, n-1-i, 1).
__device__ __forceinline__ float linspace01(int i, int n) {
  if (n <= 1) return 0.f;
  const float step = __fdiv_rn(1.0f, (float)(n - 1));
  if (i < n / 2) return __fmul_rn(step, (float)i);
  return __fmaf_rn(-step, (float)(n - 1 - i), 1.0f);
}

// coarse sample depth i of S between near and far (ipynb:418-425), the exact operations of coarse_z_kernel
__device__ __forceinline__ float z_at(int i, int S, float near_v, float far_v, int lindisp) {
  const float t = linspace01(i, S);
  const float omt = __fsub_rn(1.0f, t);
  if (!lindisp) return __fadd_rn(__fmul_rn(near_v, omt), __fmul_rn(far_v, t));
  return __fdiv_rn(1.0f, __fadd_rn(__fmul_rn(__fdiv_rn(1.0f, near_v), omt), __fmul_rn(__fdiv_rn(1.0f, far_v), t)));
}

}  // namespace nerfb
