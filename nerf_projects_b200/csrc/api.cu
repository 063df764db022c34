// Error reporting + device queries of the C ABI (include/nerf_b200.h).
#include "common.cuh"

namespace nerfb {
static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int sm_count() {
  static int cached[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (cached[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cached[dev] = n;
  }
  return cached[dev];
}
}  // namespace nerfb

extern "C" int nerfb_abi_version(void) { return NERFB_ABI_VERSION; }
extern "C" const char* nerfb_last_error(void) { return nerfb::g_err; }

extern "C" int nerfb_device_info(int* h_sm_count, int* h_cc_major, int* h_cc_minor) {
  int dev = 0;
  NERFB_CUDA(cudaGetDevice(&dev));
  int sm = 0, maj = 0, min = 0;
  NERFB_CUDA(cudaDeviceGetAttribute(&sm, cudaDevAttrMultiProcessorCount, dev));
  NERFB_CUDA(cudaDeviceGetAttribute(&maj, cudaDevAttrComputeCapabilityMajor, dev));
  NERFB_CUDA(cudaDeviceGetAttribute(&min, cudaDevAttrComputeCapabilityMinor, dev));
  if (h_sm_count) *h_sm_count = sm;
  if (h_cc_major) *h_cc_major = maj;
  if (h_cc_minor) *h_cc_minor = min;
  return NERFB_OK;
}
