// Weight gradients of the NeRF MLP on tcgen05 tensor cores (sm_100a).
//
// Replaces the autograd-generated  dW_l = dY_l^T X_l,  db_l = colsum(dY_l)  of every nn.Linear of
// NeRF (nerf.py:24-55) for loss.backward() (ipynb:1274).  Inputs are the training tile stores
// written by the fused forward (X_l: layer inputs) and by the data-gradient chain (dY_l), both
// as bf16 panel images [128 samples x 64 features] in the UMMA 128B-swizzled layout.  Here the
// contraction runs over SAMPLES, i.e. over the rows of those images, so both operands are read
// as MN-major tiles straight from the stored layout (no transposes anywhere).
//
//   * one CTA per SM (cta_group::1), a classic warp-specialised GEMM: warp 0 streams 64-sample
//     stages (A = dY panels, B = X panels, 8 KB bulk copies, 3-stage mbarrier ring), warp 1 issues
//     UMMA M=128 x N=256|64 x K=16 into fp32 TMEM accumulators that hold a layer's whole dW
//     (2 x 128 x 256 = all 512 columns), warps 2..5 reduce the small things on CUDA cores from
//     the same staged tiles (bias gradients = column sums of dY; alpha/rgb head gradients =
//     d_raw-weighted column sums of h7 / hv) and, at the end of a work item, drain TMEM with
//     red.global.add.f32 into the fp32 gradient tensors (through shared memory, see below);
//   * work = (layer job, range of stages).  Large launches (>= 10 000 stages = 640 k samples): the 12 jobs are laid end to
//     end on a COST timeline (stages x per-stage weight: 65 per plain layer, 50 per gamma(x) block, 73 for the views layer)
//     and every CTA takes an equal slice of it -- one work item for most CTAs, two (tail of one job + head of the next, two
//     TMEM drains) for the ~12 that straddle a job boundary.  With whole CTAs per job the slowest job ran 7 % over the mean
//     (13 vs 13.9 CTAs); the timeline removes that: 1.57 -> 1.48 ms on 786 k samples, 3.54 -> 2.96 ms on 1.57 M.  Small
//     launches keep whole CTAs per job (SMs dealt in proportion to the weights, largest remainder): a drain costs 35-150 us,
//     which a short work item does not amortise (0.54 vs 0.73 ms on 262 k samples).  Either way the kernel is a split-K GEMM
//     over all jobs at once and every byte of X / dY is read from HBM exactly once per job that needs it.  The drain goes
//     through a 1 KB shared-memory tile per warp and adds 8 contiguous floats per row with 8 lanes: whole 32-byte sectors
//     per reduction instead of one float (or float4) per row -- the drain of a launch fell from 0.11-0.15 ms to ~0.02.
// Bound: HBM (10.4 KB per sample against 1.19 MFLOP), see DESIGN.md.
#include "wgrad_body.cuh"

namespace nerfb {

__global__ void __launch_bounds__(kWgThreads, 1) nerf_wgrad_kernel(const __grid_constant__ WArgs a) {
  wgrad_body<false>(a, (int)blockIdx.x);
}

}  // namespace nerfb

using namespace nerfb;

// use_viewdirs != 0: 12 jobs (pts layers, feature + alpha head, views + rgb head); else 10 jobs (pts layers and a
// CUDA-core-only job for output_linear from the staged h7 panels)
static int launch_wgrad(const char* who, long long nrows, int use_viewdirs, int out_ch, const void* acts, const void* gstore,
                        const float* d_raw_padded, const float* d_raw4_padded, float* const* h_grads, void* stream) {
  NERFB_REQUIRE(nrows >= 0, "%s: bad arguments", who);
  if (nrows == 0) return NERFB_OK;
  WArgs a;
  const int grid = sm_count() >= kMaxJobs ? sm_count() : kMaxJobs;
  if (int rc = build_wgrad_args(who, nrows, use_viewdirs, out_ch, acts, gstore, d_raw_padded, d_raw4_padded, h_grads, grid, a))
    return rc;
  static bool attr_set[64] = {false};      // function attributes are per device
  int dev = 0;
  NERFB_CUDA(cudaGetDevice(&dev));
  NERFB_REQUIRE(dev >= 0 && dev < 64, "nerf_wgrad: device ordinal %d out of range", dev);
  if (!attr_set[dev]) {
    NERFB_CUDA(cudaFuncSetAttribute(nerf_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kWgSmemBytes));
    attr_set[dev] = true;
  }
  nerf_wgrad_kernel<<<grid, kWgThreads, kWgSmemBytes, as_stream(stream)>>>(a);
  return check_launch("nerf_wgrad_kernel");
}

extern "C" int nerfb_mlp_wgrad_bf16(long long nrows, const void* acts, const void* gstore, const float* d_raw_padded,
                                    float* const* h_grads, void* stream) {
  return launch_wgrad("nerfb_mlp_wgrad_bf16", nrows, 1, 4, acts, gstore, d_raw_padded, nullptr, h_grads, stream);
}

extern "C" int nerfb_mlp_wgrad_nv_bf16(long long nrows, int out_ch, const void* acts, const void* gstore,
                                       const float* d_raw_padded, const float* d_raw4_padded, float* const* h_grads,
                                       void* stream) {
  NERFB_REQUIRE(out_ch >= 1 && out_ch <= 5, "nerfb_mlp_wgrad_nv_bf16: out_ch=%d unsupported", out_ch);
  NERFB_REQUIRE(out_ch < 5 || d_raw4_padded != nullptr, "nerfb_mlp_wgrad_nv_bf16: out_ch=5 needs d_raw4_padded");
  return launch_wgrad("nerfb_mlp_wgrad_nv_bf16", nrows, 0, out_ch, acts, gstore, d_raw_padded,
                      out_ch == 5 ? d_raw4_padded : nullptr, h_grads, stream);
}
