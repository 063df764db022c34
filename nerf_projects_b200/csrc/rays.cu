// Ray generation / packing / coarse depths.
// Replaces get_rays, ndc_rays (nerf_helpers.py:222-369), the packing block of
// render() (ipynb:598-629) and the z_vals block of render_rays (ipynb:418-444).
// Every fp32 operation is issued with explicit round-to-nearest intrinsics so
// nvcc cannot contract mul+add into FMA: the results are bit-identical to the
// reference's eager fp32 ops.
#include "common.cuh"

namespace nerfb {

__global__ void get_rays_kernel(int H, int W, float fx, float fy, float cx, float cy,
                                const float* __restrict__ c2w, float* __restrict__ ro,
                                float* __restrict__ rd) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= H * W) return;
  const float i = (float)(p % W), j = (float)(p / W);
  const float dx = __fdiv_rn(__fsub_rn(i, cx), fx);
  const float dy = -__fdiv_rn(__fsub_rn(j, cy), fy);
  const float dz = -1.0f;
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const float r0 = __ldg(c2w + 4 * k + 0), r1 = __ldg(c2w + 4 * k + 1), r2 = __ldg(c2w + 4 * k + 2);
    // torch.sum over 3 products: ((0+p0)+p1)+p2
    rd[3 * p + k] = __fadd_rn(__fadd_rn(__fmul_rn(dx, r0), __fmul_rn(dy, r1)), __fmul_rn(dz, r2));
    ro[3 * p + k] = __ldg(c2w + 4 * k + 3);
  }
}


// Ray-batch producer of train() (ipynb:1167-1255): rays of SELECTED pixels generated directly from
// (pose, intrinsics, pixel id) -- same arithmetic as get_rays_kernel, so the values are bit-identical to
// indexing the reference's per-image get_rays() output -- plus the gather of the target colours.
// sel[i] = img * H*W + j*W + i (or pixel id only with a fixed image `img0`); poses: [N,3,4]; images [N,H,W,3].
__global__ void gather_ray_batch_kernel(int n, int H, int W, float fx, float fy, float cx, float cy,
                                        const float* __restrict__ poses, int pose_stride, const float* __restrict__ images,
                                        const long long* __restrict__ sel, int img0, float* __restrict__ ro,
                                        float* __restrict__ rd, float* __restrict__ target) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  const long long hw = (long long)H * W;
  long long s = __ldg(sel + t);
  int img = img0;
  if (img0 < 0) { img = (int)(s / hw); s -= (long long)img * hw; }
  const int p = (int)s;
  const float i = (float)(p % W), j = (float)(p / W);
  const float dx = __fdiv_rn(__fsub_rn(i, cx), fx);
  const float dy = -__fdiv_rn(__fsub_rn(j, cy), fy);
  const float dz = -1.0f;
  const float* c2w = poses + (size_t)img * pose_stride;
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const float r0 = __ldg(c2w + 4 * k + 0), r1 = __ldg(c2w + 4 * k + 1), r2 = __ldg(c2w + 4 * k + 2);
    rd[3 * t + k] = __fadd_rn(__fadd_rn(__fmul_rn(dx, r0), __fmul_rn(dy, r1)), __fmul_rn(dz, r2));
    ro[3 * t + k] = __ldg(c2w + 4 * k + 3);
  }
  if (images != nullptr) {
    const float* px = images + ((size_t)img * hw + p) * 3;
    target[3 * t + 0] = __ldg(px + 0); target[3 * t + 1] = __ldg(px + 1); target[3 * t + 2] = __ldg(px + 2);
  }
}

struct NdcConst {
  float sx, sy, near_f, two_near, neg_two_near;
};

__device__ __forceinline__ void ndc_one(const NdcConst c, float ox, float oy, float oz, float dx,
                                        float dy, float dz, float* o, float* d) {
  // nerf_helpers.py:342-363, same operation order
  const float t = __fdiv_rn(-__fadd_rn(oz, c.near_f), dz);
  ox = __fadd_rn(ox, __fmul_rn(t, dx));
  oy = __fadd_rn(oy, __fmul_rn(t, dy));
  oz = __fadd_rn(oz, __fmul_rn(t, dz));
  o[0] = __fdiv_rn(__fmul_rn(c.sx, ox), oz);
  o[1] = __fdiv_rn(__fmul_rn(c.sy, oy), oz);
  o[2] = __fadd_rn(__fdiv_rn(c.two_near, oz), 1.0f);
  d[0] = __fmul_rn(c.sx, __fsub_rn(__fdiv_rn(dx, dz), __fdiv_rn(ox, oz)));
  d[1] = __fmul_rn(c.sy, __fsub_rn(__fdiv_rn(dy, dz), __fdiv_rn(oy, oz)));
  d[2] = __fdiv_rn(c.neg_two_near, oz);
}

static NdcConst make_ndc(int H, int W, double focal, double near_plane) {
  NdcConst c;
  c.sx = (float)(-1.0 / (W / (2.0 * focal)));
  c.sy = (float)(-1.0 / (H / (2.0 * focal)));
  c.near_f = (float)near_plane;
  c.two_near = (float)(2.0 * near_plane);
  c.neg_two_near = (float)(-2.0 * near_plane);
  return c;
}

__global__ void ndc_rays_kernel(int n, NdcConst c, const float* __restrict__ ro,
                                const float* __restrict__ rd, float* __restrict__ oo,
                                float* __restrict__ od) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n) return;
  float o[3], d[3];
  ndc_one(c, ro[3 * p], ro[3 * p + 1], ro[3 * p + 2], rd[3 * p], rd[3 * p + 1], rd[3 * p + 2], o, d);
#pragma unroll
  for (int k = 0; k < 3; ++k) { oo[3 * p + k] = o[k]; od[3 * p + k] = d[k]; }
}

__global__ void pack_rays_kernel(int n, const float* __restrict__ ro, const float* __restrict__ rd,
                                 const float* __restrict__ vd, int use_viewdirs, int ndc, NdcConst c,
                                 float near_v, float far_v, float* __restrict__ out) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n) return;
  const int w = use_viewdirs ? 11 : 8;
  float o[3] = {ro[3 * p], ro[3 * p + 1], ro[3 * p + 2]};
  float d[3] = {rd[3 * p], rd[3 * p + 1], rd[3 * p + 2]};
  float* row = out + (size_t)p * w;
  if (use_viewdirs) {
    const float vx = vd[3 * p], vy = vd[3 * p + 1], vz = vd[3 * p + 2];
    const float nrm = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(vx, vx), __fmul_rn(vy, vy)), __fmul_rn(vz, vz)));
    row[8] = __fdiv_rn(vx, nrm);
    row[9] = __fdiv_rn(vy, nrm);
    row[10] = __fdiv_rn(vz, nrm);
  }
  if (ndc) ndc_one(c, o[0], o[1], o[2], d[0], d[1], d[2], o, d);
  row[0] = o[0]; row[1] = o[1]; row[2] = o[2];
  row[3] = d[0]; row[4] = d[1]; row[5] = d[2];
  row[6] = near_v; row[7] = far_v;
}

__global__ void coarse_z_kernel(int R, int S, const float* __restrict__ rays, int rw, int lindisp,
                                const float* __restrict__ t_rand, float* __restrict__ z) {
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= (long long)R * S) return;
  const int r = (int)(e / S), i = (int)(e % S);
  const float near_v = __ldg(rays + (size_t)r * rw + 6), far_v = __ldg(rays + (size_t)r * rw + 7);
  const float zi = z_at(i, S, near_v, far_v, lindisp);
  if (t_rand == nullptr) { z[e] = zi; return; }
  // stratified jitter, ipynb:428-444
  const float lo = (i == 0) ? zi : __fmul_rn(0.5f, __fadd_rn(zi, z_at(i - 1, S, near_v, far_v, lindisp)));
  const float hi = (i == S - 1) ? zi : __fmul_rn(0.5f, __fadd_rn(z_at(i + 1, S, near_v, far_v, lindisp), zi));
  z[e] = __fadd_rn(lo, __fmul_rn(__fsub_rn(hi, lo), t_rand[e]));
}

}  // namespace nerfb

using namespace nerfb;

extern "C" int nerfb_get_rays(int H, int W, float fx, float fy, float cx, float cy, const float* c2w,
                              float* rays_o, float* rays_d, void* stream) {
  NERFB_REQUIRE(H > 0 && W > 0 && c2w && rays_o && rays_d, "nerfb_get_rays: bad arguments");
  get_rays_kernel<<<cdiv((long long)H * W, 256), 256, 0, as_stream(stream)>>>(H, W, fx, fy, cx, cy, c2w, rays_o, rays_d);
  return check_launch("get_rays_kernel");
}

extern "C" int nerfb_ndc_rays(int n, int H, int W, double focal, double near_plane, const float* rays_o,
                              const float* rays_d, float* out_o, float* out_d, void* stream) {
  NERFB_REQUIRE(n >= 0 && rays_o && rays_d && out_o && out_d, "nerfb_ndc_rays: bad arguments");
  if (n == 0) return NERFB_OK;
  ndc_rays_kernel<<<cdiv(n, 256), 256, 0, as_stream(stream)>>>(n, make_ndc(H, W, focal, near_plane), rays_o, rays_d, out_o, out_d);
  return check_launch("ndc_rays_kernel");
}

extern "C" int nerfb_pack_rays(int n, const float* rays_o, const float* rays_d, const float* view_d,
                               int use_viewdirs, int ndc, int H, int W, double focal, float near_v,
                               float far_v, float* out, void* stream) {
  NERFB_REQUIRE(n >= 0 && rays_o && rays_d && out, "nerfb_pack_rays: bad arguments");
  NERFB_REQUIRE(!use_viewdirs || view_d, "nerfb_pack_rays: use_viewdirs needs view_d");
  if (n == 0) return NERFB_OK;
  NdcConst c = make_ndc(H > 0 ? H : 1, W > 0 ? W : 1, focal, 1.0);
  pack_rays_kernel<<<cdiv(n, 256), 256, 0, as_stream(stream)>>>(n, rays_o, rays_d, view_d, use_viewdirs, ndc, c, near_v, far_v, out);
  return check_launch("pack_rays_kernel");
}

extern "C" int nerfb_coarse_z(int R, int S, const float* rays, int ray_width, int lindisp,
                              const float* t_rand, float* z, void* stream) {
  NERFB_REQUIRE(R >= 0 && S > 0 && rays && z && ray_width >= 8, "nerfb_coarse_z: bad arguments");
  if (R == 0) return NERFB_OK;
  coarse_z_kernel<<<cdiv((long long)R * S, 256), 256, 0, as_stream(stream)>>>(R, S, rays, ray_width, lindisp, t_rand, z);
  return check_launch("coarse_z_kernel");
}

extern "C" int nerfb_gather_ray_batch(int n, int H, int W, float fx, float fy, float cx, float cy, const float* poses,
                                      int pose_stride, int n_images, const float* images, const long long* sel, int img0,
                                      float* rays_o, float* rays_d, float* target, void* stream) {
  NERFB_REQUIRE(n >= 0 && H > 0 && W > 0 && poses && sel && rays_o && rays_d && pose_stride >= 12 && n_images > 0 &&
                    img0 < n_images && (images == nullptr || target != nullptr),
                "nerfb_gather_ray_batch: bad arguments");
  if (n == 0) return NERFB_OK;
  gather_ray_batch_kernel<<<cdiv(n, 256), 256, 0, as_stream(stream)>>>(n, H, W, fx, fy, cx, cy, poses, pose_stride, images,
                                                                        sel, img0, rays_o, rays_d, target);
  return check_launch("gather_ray_batch_kernel");
}
