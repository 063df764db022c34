// Alpha compositing: raw2outputs (ipynb:254-349) forward and its backward
// w.r.t. raw.  One warp per ray; the exclusive transmittance product is a warp
// scan in fp64 rounded to fp32 per sample, which is what torch.cumprod does on
// CPU (double running product).
#include <cstdlib>
#include "composite.cuh"

namespace nerfb {

constexpr int CW = 4;  // warps per block

__global__ void __launch_bounds__(32 * CW)
raw2outputs_fwd_kernel(int R, int S, const float* __restrict__ raw, int ch, const float* __restrict__ z,
                       const float* __restrict__ rays_d, int ld_d, const float* __restrict__ noise,
                       int white, float* __restrict__ rgb, float* __restrict__ disp, float* __restrict__ acc,
                       float* __restrict__ weights, float* __restrict__ depth) {
  const int lane = threadIdx.x & 31;
  const int r = blockIdx.x * CW + (threadIdx.x >> 5);
  if (r >= R) return;
  const float* rr = raw + (size_t)r * S * ch;
  const float* zr = z + (size_t)r * S;
  const float* nr = noise ? noise + (size_t)r * S : nullptr;
  const float nrm = ray_norm(rays_d + (size_t)r * ld_d);
  float c0 = 0.f, c1 = 0.f, c2 = 0.f, dsum = 0.f, asum = 0.f;
  float* wr = weights ? weights + (size_t)r * S : nullptr;
  scan_ray(S, rr, ch, zr, nrm, nr, lane, [&](int i, float alpha, float x, float T) {
    const float w = __fmul_rn(alpha, T);
    if (wr) wr[i] = w;
    c0 += w * sigmoidf_(rr[(size_t)i * ch + 0]);
    c1 += w * sigmoidf_(rr[(size_t)i * ch + 1]);
    c2 += w * sigmoidf_(rr[(size_t)i * ch + 2]);
    dsum += w * zr[i];
    asum += w;
  });
  c0 = warp_sum(c0); c1 = warp_sum(c1); c2 = warp_sum(c2); dsum = warp_sum(dsum); asum = warp_sum(asum);
  if (lane == 0) {
    if (white) {
      const float bg = __fsub_rn(1.0f, asum);
      c0 += bg; c1 += bg; c2 += bg;
    }
    rgb[3 * r + 0] = c0; rgb[3 * r + 1] = c1; rgb[3 * r + 2] = c2;
    if (depth) depth[r] = dsum;
    if (acc) acc[r] = asum;
    if (disp) disp[r] = __fdiv_rn(1.0f, fmaxf(__fdiv_rn(dsum, fmaxf(1e-10f, asum)), 1e-10f));
  }
}

__global__ void __launch_bounds__(32 * CW)
raw2outputs_bwd_kernel(int R, int S, const float* __restrict__ raw, int ch, const float* __restrict__ z,
                       const float* __restrict__ rays_d, int ld_d, const float* __restrict__ noise,
                       int white, const float* __restrict__ d_rgb, const float* __restrict__ d_disp,
                       const float* __restrict__ d_acc, const float* __restrict__ d_weights,
                       const float* __restrict__ d_depth, float* __restrict__ d_raw) {
  extern __shared__ float smem[];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int r = blockIdx.x * CW + wib;
  if (r >= R) return;
  float* s_alpha = smem + (size_t)wib * 3 * S;
  float* s_x = s_alpha + S;
  float* s_T = s_x + S;
  const float* rr = raw + (size_t)r * S * ch;
  const float* zr = z + (size_t)r * S;
  const float* nr = noise ? noise + (size_t)r * S : nullptr;
  const float nrm = ray_norm(rays_d + (size_t)r * ld_d);

  float dsum = 0.f, asum = 0.f;
  scan_ray(S, rr, ch, zr, nrm, nr, lane, [&](int i, float alpha, float x, float T) {
    s_alpha[i] = alpha; s_x[i] = x; s_T[i] = T;
    const float w = __fmul_rn(alpha, T);
    dsum += w * zr[i];
    asum += w;
  });
  dsum = warp_sum(dsum); asum = warp_sum(asum);
  __syncwarp();

  const float g0 = d_rgb ? d_rgb[3 * r + 0] : 0.f, g1 = d_rgb ? d_rgb[3 * r + 1] : 0.f,
              g2 = d_rgb ? d_rgb[3 * r + 2] : 0.f;
  float g_depth = d_depth ? d_depth[r] : 0.f;
  float g_acc = d_acc ? d_acc[r] : 0.f;
  if (white) g_acc -= (g0 + g1 + g2);
  if (d_disp) {
    const float denom = fmaxf(1e-10f, asum);
    const float q = dsum / denom;
    if (q > 1e-10f) {
      const float gq = -d_disp[r] / (q * q);
      g_depth += gq / denom;
      if (asum > 1e-10f) g_acc += -gq * dsum / (denom * denom);
    }
  }
  const float* gwr = d_weights ? d_weights + (size_t)r * S : nullptr;
  float* out = d_raw + (size_t)r * S * ch;

  // reverse pass: suffix[i] = sum_{k>i} gw_k * w_k
  float carry = 0.f;
  const int nchunk = (S + 31) / 32;
  for (int cidx = nchunk - 1; cidx >= 0; --cidx) {
    const int i = cidx * 32 + lane;
    float gw = 0.f, w = 0.f, alpha = 0.f, x = 1.f, T = 0.f, s0 = 0.f, s1 = 0.f, s2 = 0.f;
    if (i < S) {
      alpha = s_alpha[i]; x = s_x[i]; T = s_T[i];
      w = alpha * T;
      s0 = sigmoidf_(rr[(size_t)i * ch + 0]);
      s1 = sigmoidf_(rr[(size_t)i * ch + 1]);
      s2 = sigmoidf_(rr[(size_t)i * ch + 2]);
      gw = g0 * s0 + g1 * s1 + g2 * s2 + g_depth * zr[i] + g_acc + (gwr ? gwr[i] : 0.f);
    }
    const float v = gw * w;
    // inclusive suffix scan within the chunk
    float sfx = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const float q = __shfl_down_sync(0xffffffffu, sfx, o);
      if (lane + o < 32) sfx += q;
    }
    const float after = sfx - v + carry;  // strictly after i
    if (i < S) {
      const float g_alpha = gw * T - after / x;
      float dist = (i + 1 < S) ? __fsub_rn(zr[i + 1], zr[i]) : 1e10f;
      dist = __fmul_rn(dist, nrm);
      float s = rr[(size_t)i * ch + 3];
      if (nr) s = __fadd_rn(s, nr[i]);
      float g_sigma = 0.f;
      if (s > 0.f) g_sigma = g_alpha * (1.0f - alpha) * dist;  // exp(-s*dist) = 1-alpha
      out[(size_t)i * ch + 0] = g0 * w * s0 * (1.f - s0);
      out[(size_t)i * ch + 1] = g1 * w * s1 * (1.f - s1);
      out[(size_t)i * ch + 2] = g2 * w * s2 * (1.f - s2);
      out[(size_t)i * ch + 3] = g_sigma;
      for (int c = 4; c < ch; ++c) out[(size_t)i * ch + c] = 0.f;
    }
    carry += __shfl_sync(0xffffffffu, sfx, 0);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Register-resident variants for raw_ch == 4 and S <= 32*NCH (NCH = 2, 4, 6, 8: S = 64, 128, 192, 256 ...): every lane
// first issues ALL of its loads of the ray (NCH float4 of raw, NCH depths, NCH noise values) back to back -- 6x the
// memory-level parallelism of the chunk-by-chunk loop above at S = 192, which was latency-bound at 2.2-2.7 TB/s
// (profiles/r01_ncu_full_io_kernels.txt) -- and then runs the SAME arithmetic in the same order (bit-identical maps).
template <int NCH>
struct RayRegs {
  float4 rv[NCH];
  float zv[NCH], nv[NCH];
};

template <int NCH>
__device__ __forceinline__ void load_ray(RayRegs<NCH>& q, int S, const float* __restrict__ rr, const float* __restrict__ zr,
                                         const float* __restrict__ nr, int lane) {
#pragma unroll
  for (int k = 0; k < NCH; ++k) {
    const int i = 32 * k + lane;
    q.rv[k] = make_float4(0.f, 0.f, 0.f, 0.f);
    q.zv[k] = 0.f;
    q.nv[k] = 0.f;
    if (i < S) {
      q.rv[k] = __ldg(reinterpret_cast<const float4*>(rr) + i);
      q.zv[k] = __ldg(zr + i);
      if (nr) q.nv[k] = __ldg(nr + i);
    }
  }
}

// alpha, x = 1 - alpha + 1e-10 and the exclusive transmittance T of chunk k (same operations as scan_ray)
template <int NCH>
__device__ __forceinline__ void chunk_alpha(const RayRegs<NCH>& q, int k, int S, float nrm, bool has_noise, int lane,
                                            double& carry, float& alpha, float& x, float& T, float& dist_out) {
  const int i = 32 * k + lane;
  float zn = __shfl_down_sync(0xffffffffu, q.zv[k], 1);
  const float zfirst = __shfl_sync(0xffffffffu, q.zv[k + 1 < NCH ? k + 1 : k], 0);
  if (lane == 31) zn = zfirst;
  alpha = 0.f; x = 1.f; dist_out = 0.f;
  if (i < S) {
    float dist = (i + 1 < S) ? __fsub_rn(zn, q.zv[k]) : 1e10f;
    dist = __fmul_rn(dist, nrm);
    dist_out = dist;
    float s = q.rv[k].w;
    if (has_noise) s = __fadd_rn(s, q.nv[k]);
    const float act = fmaxf(s, 0.f);
    alpha = __fsub_rn(1.0f, expf(__fmul_rn(-act, dist)));
    x = __fadd_rn(__fsub_rn(1.0f, alpha), 1e-10f);
  }
  double p = (double)x;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const double qq = __shfl_up_sync(0xffffffffu, p, o);
    if (lane >= o) p *= qq;
  }
  double excl = __shfl_up_sync(0xffffffffu, p, 1);
  if (lane == 0) excl = 1.0;
  T = (float)(carry * excl);
  carry *= __shfl_sync(0xffffffffu, p, 31);
}

template <int NCH>
__global__ void __launch_bounds__(32 * CW)
raw2outputs_fwd_regs_kernel(int R, int S, const float* __restrict__ raw, const float* __restrict__ z,
                            const float* __restrict__ rays_d, int ld_d, const float* __restrict__ noise, int white,
                            float* __restrict__ rgb, float* __restrict__ disp, float* __restrict__ acc,
                            float* __restrict__ weights, float* __restrict__ depth) {
  const int lane = threadIdx.x & 31;
  const int r = blockIdx.x * CW + (threadIdx.x >> 5);
  if (r >= R) return;
  RayRegs<NCH> q;
  load_ray<NCH>(q, S, raw + (size_t)r * S * 4, z + (size_t)r * S, noise ? noise + (size_t)r * S : nullptr, lane);
  const float nrm = ray_norm(rays_d + (size_t)r * ld_d);
  float c0 = 0.f, c1 = 0.f, c2 = 0.f, dsum = 0.f, asum = 0.f;
  float* wr = weights ? weights + (size_t)r * S : nullptr;
  double carry = 1.0;
#pragma unroll
  for (int k = 0; k < NCH; ++k) {
    if (32 * k >= S) break;
    float alpha, x, T, dist;
    chunk_alpha<NCH>(q, k, S, nrm, noise != nullptr, lane, carry, alpha, x, T, dist);
    const int i = 32 * k + lane;
    if (i < S) {
      const float w = __fmul_rn(alpha, T);
      if (wr) wr[i] = w;
      c0 += w * sigmoidf_(q.rv[k].x);
      c1 += w * sigmoidf_(q.rv[k].y);
      c2 += w * sigmoidf_(q.rv[k].z);
      dsum += w * q.zv[k];
      asum += w;
    }
  }
  c0 = warp_sum(c0); c1 = warp_sum(c1); c2 = warp_sum(c2); dsum = warp_sum(dsum); asum = warp_sum(asum);
  if (lane == 0) {
    if (white) {
      const float bg = __fsub_rn(1.0f, asum);
      c0 += bg; c1 += bg; c2 += bg;
    }
    rgb[3 * r + 0] = c0; rgb[3 * r + 1] = c1; rgb[3 * r + 2] = c2;
    if (depth) depth[r] = dsum;
    if (acc) acc[r] = asum;
    if (disp) disp[r] = __fdiv_rn(1.0f, fmaxf(__fdiv_rn(dsum, fmaxf(1e-10f, asum)), 1e-10f));
  }
}

template <int NCH>
__global__ void __launch_bounds__(32 * CW)
raw2outputs_bwd_regs_kernel(int R, int S, const float* __restrict__ raw, const float* __restrict__ z,
                            const float* __restrict__ rays_d, int ld_d, const float* __restrict__ noise, int white,
                            const float* __restrict__ d_rgb, const float* __restrict__ d_disp,
                            const float* __restrict__ d_acc, const float* __restrict__ d_weights,
                            const float* __restrict__ d_depth, float* __restrict__ d_raw) {
  const int lane = threadIdx.x & 31;
  const int r = blockIdx.x * CW + (threadIdx.x >> 5);
  if (r >= R) return;
  RayRegs<NCH> q;
  load_ray<NCH>(q, S, raw + (size_t)r * S * 4, z + (size_t)r * S, noise ? noise + (size_t)r * S : nullptr, lane);
  const float* gwr = d_weights ? d_weights + (size_t)r * S : nullptr;
  float gwv[NCH];
#pragma unroll
  for (int k = 0; k < NCH; ++k) gwv[k] = (gwr && 32 * k + lane < S) ? __ldg(gwr + 32 * k + lane) : 0.f;
  const float nrm = ray_norm(rays_d + (size_t)r * ld_d);
  float al[NCH], xs[NCH], Ts[NCH], ds[NCH];
  float dsum = 0.f, asum = 0.f;
  double carry_p = 1.0;
#pragma unroll
  for (int k = 0; k < NCH; ++k) {
    al[k] = 0.f; xs[k] = 1.f; Ts[k] = 0.f; ds[k] = 0.f;
    if (32 * k >= S) continue;
    chunk_alpha<NCH>(q, k, S, nrm, noise != nullptr, lane, carry_p, al[k], xs[k], Ts[k], ds[k]);
    if (32 * k + lane < S) {
      const float w = __fmul_rn(al[k], Ts[k]);
      dsum += w * q.zv[k];
      asum += w;
    }
  }
  dsum = warp_sum(dsum); asum = warp_sum(asum);

  const float g0 = d_rgb ? d_rgb[3 * r + 0] : 0.f, g1 = d_rgb ? d_rgb[3 * r + 1] : 0.f,
              g2 = d_rgb ? d_rgb[3 * r + 2] : 0.f;
  float g_depth = d_depth ? d_depth[r] : 0.f;
  float g_acc = d_acc ? d_acc[r] : 0.f;
  if (white) g_acc -= (g0 + g1 + g2);
  if (d_disp) {
    const float denom = fmaxf(1e-10f, asum);
    const float qd = dsum / denom;
    if (qd > 1e-10f) {
      const float gq = -d_disp[r] / (qd * qd);
      g_depth += gq / denom;
      if (asum > 1e-10f) g_acc += -gq * dsum / (denom * denom);
    }
  }
  float4* out = reinterpret_cast<float4*>(d_raw + (size_t)r * S * 4);
  float carry = 0.f;      // sum of gw_k * w_k over the chunks after this one
#pragma unroll
  for (int k = NCH - 1; k >= 0; --k) {
    if (32 * k >= S) continue;
    const int i = 32 * k + lane;
    float gw = 0.f, w = 0.f, s0 = 0.f, s1 = 0.f, s2 = 0.f;
    if (i < S) {
      w = al[k] * Ts[k];
      s0 = sigmoidf_(q.rv[k].x); s1 = sigmoidf_(q.rv[k].y); s2 = sigmoidf_(q.rv[k].z);
      gw = g0 * s0 + g1 * s1 + g2 * s2 + g_depth * q.zv[k] + g_acc + gwv[k];
    }
    const float v = gw * w;
    float sfx = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const float qq = __shfl_down_sync(0xffffffffu, sfx, o);
      if (lane + o < 32) sfx += qq;
    }
    const float after = sfx - v + carry;
    if (i < S) {
      const float g_alpha = gw * Ts[k] - after / xs[k];
      float s = q.rv[k].w;
      if (noise) s = __fadd_rn(s, q.nv[k]);
      float g_sigma = 0.f;
      if (s > 0.f) g_sigma = g_alpha * (1.0f - al[k]) * ds[k];
      out[i] = make_float4(g0 * w * s0 * (1.f - s0), g1 * w * s1 * (1.f - s1), g2 * w * s2 * (1.f - s2), g_sigma);
    }
    carry += __shfl_sync(0xffffffffu, sfx, 0);
  }
}

// Finish of the rays whose segments nerf_mlp_kernel<FWD_COMP> composited (mlp_tc.cu: composite_tile): one thread per
// ray walks the ray's nseg records {Q, -, sum w', sum w' r, g, b, sum w' z, -} in order, carrying the
// transmittance T = product of the earlier Q in fp64 like torch.cumprod's running product, and writes the maps of
// raw2outputs (ipynb:332-347).  32 bytes per 32 samples are read: 1 byte per sample instead of the 20 of raw + z.
__global__ void __launch_bounds__(128)
composite_finish_kernel(int R, int nseg, const float* __restrict__ rec, int white, float* __restrict__ rgb,
                        float* __restrict__ disp, float* __restrict__ acc, float* __restrict__ depth) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= R) return;
  const float4* q = reinterpret_cast<const float4*>(rec + (size_t)r * nseg * 8);
  double T = 1.0;
  float asum = 0.f, c0 = 0.f, c1 = 0.f, c2 = 0.f, dsum = 0.f;
  for (int s = 0; s < nseg; ++s) {
    const float4 ra = __ldg(q + 2 * s), rb = __ldg(q + 2 * s + 1);
    const float Tf = (float)T;
    asum = fmaf(Tf, ra.z, asum); c0 = fmaf(Tf, ra.w, c0); c1 = fmaf(Tf, rb.x, c1); c2 = fmaf(Tf, rb.y, c2);
    dsum = fmaf(Tf, rb.z, dsum);
    T *= (double)ra.x;
  }
  if (white) {
    const float bg = __fsub_rn(1.0f, asum);
    c0 += bg; c1 += bg; c2 += bg;
  }
  rgb[3 * r + 0] = c0; rgb[3 * r + 1] = c1; rgb[3 * r + 2] = c2;
  if (depth) depth[r] = dsum;
  if (acc) acc[r] = asum;
  if (disp) disp[r] = __fdiv_rn(1.0f, fmaxf(__fdiv_rn(dsum, fmaxf(1e-10f, asum)), 1e-10f));
}

// NCH for the register-resident kernels, or 0 when they do not apply
inline int regs_nch(int S, int raw_ch, const void* raw, const void* out) {
  if (raw_ch != 4 || !aligned16(raw) || !aligned16(out) || S > 256) return 0;
  static const bool force_generic = getenv("NERFB_COMPOSITE_GENERIC") != nullptr;    // A/B measurements only
  if (force_generic) return 0;
  return S <= 64 ? 2 : (S <= 128 ? 4 : (S <= 192 ? 6 : 8));
}

}  // namespace nerfb

using namespace nerfb;

extern "C" int nerfb_raw2outputs_fwd(int R, int S, const float* raw, int raw_ch, const float* z,
                                     const float* rays_d, int ld_d, const float* noise, int white_bkgd,
                                     float* rgb, float* disp, float* acc, float* weights, float* depth,
                                     void* stream) {
  NERFB_REQUIRE(R >= 0 && S > 0 && raw && z && rays_d && rgb && raw_ch >= 4 && ld_d >= 3,
                "nerfb_raw2outputs_fwd: bad arguments");
  if (R == 0) return NERFB_OK;
  const dim3 grid(cdiv(R, CW)), block(32 * CW);
  cudaStream_t st = as_stream(stream);
#define NERFB_FWD_REGS(N) raw2outputs_fwd_regs_kernel<N><<<grid, block, 0, st>>>(R, S, raw, z, rays_d, ld_d, noise, \
                                                                               white_bkgd, rgb, disp, acc, weights, depth)
  switch (regs_nch(S, raw_ch, raw, raw)) {
    case 2: NERFB_FWD_REGS(2); break;
    case 4: NERFB_FWD_REGS(4); break;
    case 6: NERFB_FWD_REGS(6); break;
    case 8: NERFB_FWD_REGS(8); break;
    default:
      raw2outputs_fwd_kernel<<<grid, block, 0, st>>>(R, S, raw, raw_ch, z, rays_d, ld_d, noise, white_bkgd, rgb, disp, acc,
                                                     weights, depth);
  }
#undef NERFB_FWD_REGS
  return check_launch("raw2outputs_fwd_kernel");
}

extern "C" int nerfb_composite_finish(int R, int nseg, const float* rec, int white_bkgd, float* rgb, float* disp, float* acc,
                                      float* depth, void* stream) {
  NERFB_REQUIRE(R >= 0 && nseg >= 1 && nseg <= 32 && rec && rgb && aligned16(rec), "nerfb_composite_finish: bad arguments");
  if (R == 0) return NERFB_OK;
  composite_finish_kernel<<<cdiv(R, 128), 128, 0, as_stream(stream)>>>(R, nseg, rec, white_bkgd, rgb, disp, acc, depth);
  return check_launch("composite_finish_kernel");
}

extern "C" int nerfb_raw2outputs_bwd(int R, int S, const float* raw, int raw_ch, const float* z,
                                     const float* rays_d, int ld_d, const float* noise, int white_bkgd,
                                     const float* d_rgb, const float* d_disp, const float* d_acc,
                                     const float* d_weights, const float* d_depth, float* d_raw, void* stream) {
  NERFB_REQUIRE(R >= 0 && S > 0 && raw && z && rays_d && d_raw && raw_ch >= 4 && ld_d >= 3,
                "nerfb_raw2outputs_bwd: bad arguments");
  NERFB_REQUIRE(S <= 1024, "nerfb_raw2outputs_bwd: at most 1024 samples per ray (got %d)", S);
  if (R == 0) return NERFB_OK;
  const dim3 grid(cdiv(R, CW)), block(32 * CW);
  cudaStream_t st = as_stream(stream);
#define NERFB_BWD_REGS(N) raw2outputs_bwd_regs_kernel<N><<<grid, block, 0, st>>>(R, S, raw, z, rays_d, ld_d, noise, white_bkgd, \
                                                                               d_rgb, d_disp, d_acc, d_weights, d_depth, d_raw)
  switch (regs_nch(S, raw_ch, raw, d_raw)) {
    case 2: NERFB_BWD_REGS(2); break;
    case 4: NERFB_BWD_REGS(4); break;
    case 6: NERFB_BWD_REGS(6); break;
    case 8: NERFB_BWD_REGS(8); break;
    default: {
      const size_t smem = (size_t)CW * 3 * S * sizeof(float);
      raw2outputs_bwd_kernel<<<grid, block, smem, st>>>(R, S, raw, raw_ch, z, rays_d, ld_d, noise, white_bkgd, d_rgb, d_disp,
                                                        d_acc, d_weights, d_depth, d_raw);
    }
  }
#undef NERFB_BWD_REGS
  return check_launch("raw2outputs_bwd_kernel");
}
